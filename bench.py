#!/usr/bin/env python
"""bench.py -- RTFx of the hot path (RNN forward -> Linear -> log-softmax -> CTC beam search) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

Workload = BASELINE.json configs[4] ("cfg5", the configuration the RTFx-at-1/2/4/8-GPUs metric is quoted on): 8192 synthetic
utterances of the cfg2 shape (T=1000 frames x 161 bins, 3-layer tanh RNN H=512, Linear 512->29 + log-softmax, beam 16),
torch-default random weights.  The 8192 utterances are cut into batches of --wave utterances; rank r of N owns a CONTIGUOUS
range of batches (strong scaling, utterances are independent: no data-path collective), runs them through gasr_job_*
(several batches in flight per GPU) and the transcripts + scores are gathered on the host.  A "step" is one pass over all
8192 utterances.  Rank 0 prints ONE JSON line.

  value     RTFx with every batch already resident in HBM; transcripts + scores land in host memory inside the timed
            region.  Device-timed (CUDA events on the library's stream), max over ranks.
  e2e       the same through the host-buffer entry point gasr_job_run_host: every batch is copied from pinned host memory
            each step (h2d_bytes_per_step) and the results are read back (d2h_bytes_per_step).
  roofline  the recurrence kernel (largest algorithmic byte count of the path) against the measured HBM peak, from per-launch
            CUDA-event durations inside the running pipeline; stage_rooflines has every stage.
  parity_checked / gather  oracle check of the benched computation (log-probs 1e-4, transcripts and scores bit-exact) and
            bit-equality of the N-GPU gathered result with the same utterances decoded by rank 0 alone.
  cpu_baseline / --impl reference: the CPU restatement of the reference (oracle/, "port": the reference ships no runnable
            CPU implementation of this path -- CTCBeamSearch.cpp does not compile) on the box's host cores,
            on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "gpu-accelerated-speech-recognition_b200")
sys.path.insert(0, PKG)
sys.path.insert(0, ROOT)

METRIC = "RTFx: audio-sec decoded/sec (RNN fwd + CTC beam) at 1/2/4/8 B200"
UNIT = "audio-seconds per second"
CFG = dict(T=1000, D=161, H=512, L=3, V=29, beam=16)          # the cfg2 shape every utterance of cfg5 has
UTTS, WAVE_MAX, LANES = 8192, 4096, 2                        # BASELINE.json configs[4]; largest batch; batches in flight per GPU
SEED_X, SEED_W, SEED_FC = 1234, 4321, 99
FRAME_SEC = 0.010
WORKLOAD = ("cfg5: 8192 utt x T=1000 x D=161 sharded over the GPUs, 3-layer tanh RNN H=512, Linear 512->29 + log-softmax, "
            "CTC beam 16")


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "measured"
    return 6650.0, 1400.0, "fallback"


def rec_traffic(wave):
    """DRAM bytes (read + write) of one recurrence launch from the committed `ncu --set full` capture (profiles/), if that
    capture was taken at this batch size; None otherwise."""
    path = os.path.join(ROOT, "profiles", "r2_rnn_wide2_traffic.json")
    if not os.path.exists(path):
        return None
    t = json.load(open(path))
    return t["dram_bytes_per_launch"] if int(t.get("utterances_per_launch", 0)) == int(wave) else None


def bind_host_memory_to_gpu_node(local):
    """Pinned input buffers should live on the NUMA node the GPU hangs off: with one process per GPU and no CPU binding every
    rank's pages land on the node it happens to run on, and the H2D copies of the GPUs of the other socket cross the socket
    link.  MPOL_PREFERRED for this process (set_mempolicy) before the buffers are allocated; a no-op on single-node hosts.
    GASR_BENCH_NUMA=0 switches it off.  Returns the node or None."""
    if os.environ.get("GASR_BENCH_NUMA", "1") == "0":
        return None
    try:
        import ctypes
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        if len(nodes) < 2:
            return None
        idx = str(local)
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",")]
            if local < len(ids):
                idx = ids[local]
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", idx],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]                                   # sysfs spells the PCI domain with four digits
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        mask = ctypes.c_ulong(1 << node)
        libc = ctypes.CDLL("libc.so.6", use_errno=True)
        if libc.syscall(238, 1, ctypes.byref(mask), 64) != 0:          # SYS_set_mempolicy (x86-64), MPOL_PREFERRED
            return None
        return node
    except Exception:
        return None


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def weights():
    import synth
    c = CFG
    return synth.rnn_weights(SEED_W, c["D"], c["H"], c["L"]), synth.fc_weights(SEED_FC, c["H"], c["V"])


def cpu_port_rtfx(n_utt, cores, first_utt=0):
    """The oracle port (CPU restatement of the reference path) over n_utt utterances of the workload."""
    import synth
    from oracle import oracle as O   # bench.py's cpu_baseline / reference leg: the one place it may run
    c = CFG
    x = synth.spectrogram_batch(SEED_X, c["T"], n_utt, c["D"], first_utt=first_utt)
    w, (fc_w, fc_b) = weights()
    t0 = time.perf_counter()
    h = O.rnn_forward(x, c["T"], n_utt, *w, nthreads=cores)[-1]
    logp = O.linear(h, fc_w, fc_b, act="logsoftmax")
    O.ctc_decode(logp.reshape(c["T"], n_utt, c["V"]), synth.VOCAB29, 0, c["beam"], domain="log", nthreads=cores)
    dt = time.perf_counter() - t0
    return n_utt * c["T"] * FRAME_SEC / dt, dt


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.path = device, None, f"/tmp/gasr_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.device)], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.proc.wait()
        sm, smax, reasons = [], [], set()
        for line in open(self.path):
            f = [s.strip() for s in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "samples": len(sm)}


def run_reference(args, rank, world):
    """--impl reference: the CPU port on all host threads, bounded sample per step (rank 0 only)."""
    if rank != 0:
        return
    cores = host_cores()
    n_utt = max(cores, 8)
    c = CFG
    for _ in range(args.warmup):
        cpu_port_rtfx(max(2, cores // 4), cores)
    t_total, audio = 0.0, 0.0
    for _ in range(args.steps):
        _, dt = cpu_port_rtfx(n_utt, cores)
        t_total += dt
        audio += n_utt * c["T"] * FRAME_SEC
    value = audio / t_total
    sample = f"{n_utt} utterances x T={c['T']} per step (of the {args.utts}), {cores} host threads, RTFx-normalised"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def oracle_check(gasr, job, x_dev0, res0, n_check):
    """The benched computation against the CPU oracle: utterances 0 .. n_check-1 of batch 0.  Log-probabilities within 1e-4
    of the oracle's forward pass; transcripts and fp32 scores of the GPU decoder bit-exact against the oracle decoder run on
    the GPU's own log-probabilities (the decoder contract, tests/test_gpu_parity.py)."""
    import synth
    from oracle import oracle as O   # checker only
    c = CFG
    paths, lens, scores = job.run_device([x_dev0])            # one batch -> lane 0
    assert (paths == res0[0]).all() and (lens == res0[1]).all() and (scores.view(np.uint32) == res0[2].view(np.uint32)).all(), \
        "batch 0 decoded alone differs from batch 0 inside the job"
    logp = job.lane_logprobs(0).reshape(c["T"], args_wave(job), c["V"])[:, :n_check, :]
    w, (fc_w, fc_b) = weights()
    x = synth.spectrogram_batch(SEED_X, c["T"], n_check, c["D"], first_utt=0)
    ref = O.linear(O.rnn_forward(x, c["T"], n_check, *w, nthreads=host_cores())[-1], fc_w, fc_b, act="logsoftmax")
    err = float(np.abs(logp.reshape(-1, c["V"]) - ref).max())
    op, os_ = O.ctc_decode(np.ascontiguousarray(logp), synth.VOCAB29, 0, c["beam"], domain="log", nthreads=host_cores())
    gp, gs = gasr.unpack_results(paths[:n_check], lens[:n_check], scores[:n_check], job.cfg.max_len)
    same = gp == op and all(np.float32(a).view(np.uint32) == np.float32(b).view(np.uint32) for a, b in zip(gs, os_))
    return {"utterances": n_check, "frames": c["T"], "logprob_max_abs_err": err, "logprob_tol": 1e-4,
            "transcripts_and_scores_bit_exact": bool(same), "ok": bool(err < 1e-4 and same)}


def args_wave(job):
    return job.cfg.N


def other_configs(gasr, device):
    """The other BASELINE.json configs, each bounded to a few seconds: cfg1 (baseline/config.json hot path), cfg2 (ONE batch of
    64 utterances: a latency case), cfg3 (bidirectional GRU stack, bf16 projection), cfg4 (decode-only, T = 4000, beam sweep).
    Parity at these sizes is enforced by tests/test_gpu_sizes.py; here cfg1 and cfg4 are re-checked live (oracle / golden)."""
    import synth
    from oracle import oracle as O   # checker only
    out = {}
    ctx = gasr.Context(device)

    def time_pipe(pipe, xd, reps):
        pipe.run_device(xd)
        best = 1e30
        for _ in range(reps):
            ctx.sync(); ctx.timer_start()
            res = pipe.run_device(xd)
            best = min(best, ctx.timer_stop())
        return best, res

    # cfg1: T=200, N=1, 2048 -> tanh RNN 2048 -> Linear 47 + log-softmax, beam 100 (> V)
    T, N, D, H, L, V, beam = 200, 1, 2048, 2048, 1, 47, 100
    vocab = bytes(range(1, V + 1))
    x = synth.spectrogram_batch(101, T, N, D)
    w = synth.rnn_weights(102, D, H, L)
    fc = synth.fc_weights(103, H, V)
    pipe = gasr.AsrPipeline(ctx, gasr.CELL_TANH, False, T, N, D, H, L, V, beam, 0, vocab)
    pipe.set_weights(*w, *fc)
    xd = ctx.to_device(x)
    ms, (paths, scores) = time_pipe(pipe, xd, 3)
    logp = pipe.logprobs()
    ref = O.linear(O.rnn_forward(x, T, N, *w, nthreads=host_cores())[-1], *fc, act="logsoftmax")
    op, os_ = O.ctc_decode(logp.reshape(T, N, V), vocab, 0, beam, domain="log")
    out["cfg1"] = {"shape": "T=200 N=1 in=2048 H=2048 V=47 beam=100", "ms": ms, "rtfx": N * T * FRAME_SEC / (ms * 1e-3),
                   "stages_ms": pipe.stage_times(),
                   "parity": {"logprob_max_abs_err": float(np.abs(logp - ref).max()),
                              "decode_bit_exact": bool(paths == op and np.float32(scores[0]).view(np.uint32) == np.float32(os_[0]).view(np.uint32))}}
    pipe.close(); ctx.free(xd)

    # cfg2: one batch of 64 utterances (what round 1 benched): latency of a single small batch through the wave engine
    c = CFG
    x = synth.spectrogram_batch(SEED_X, c["T"], 64, c["D"])
    w, fc = weights()
    pipe = gasr.AsrPipeline(ctx, gasr.CELL_TANH, False, c["T"], 64, c["D"], c["H"], c["L"], c["V"], c["beam"], 0, synth.VOCAB29)
    pipe.set_weights(*w, *fc)
    xd = ctx.to_device(x)
    ms, _ = time_pipe(pipe, xd, 5)
    out["cfg2_single_batch"] = {"shape": "T=1000 N=64 D=161 H=512 L=3 V=29 beam=16", "ms": ms, "rtfx": 64 * c["T"] * FRAME_SEC / (ms * 1e-3),
                                "parity": "tests/test_gpu_sizes.py::test_cfg2_full_size_pipeline_vs_oracle (whole batch vs oracle)"}
    res_wave = pipe.run_device(xd)
    logp_wave = pipe.logprobs()
    pipe.close(); ctx.free(xd)
    # the same batch in the opt-in latency mode (GASR_STREAM=1, read when the context is created): three persistent kernels
    # coupled by progress counters -- needs the whole GPU to itself, so it is not a default
    os.environ["GASR_STREAM"] = "1"
    try:
        c2 = gasr.Context(device)
    finally:
        del os.environ["GASR_STREAM"]
    pipe = gasr.AsrPipeline(c2, gasr.CELL_TANH, False, c["T"], 64, c["D"], c["H"], c["L"], c["V"], c["beam"], 0, synth.VOCAB29)
    pipe.set_weights(*w, *fc)
    xd = c2.to_device(x)
    ms, _ = time_pipe(pipe, xd, 5)
    res_stream = pipe.run_device(xd)
    _, mode = pipe.stage_launches()
    out["cfg2_single_batch_latency_mode"] = {
        "shape": "as cfg2_single_batch, GASR_STREAM=1", "ms": ms, "rtfx": 64 * c["T"] * FRAME_SEC / (ms * 1e-3),
        "mode": {-2: "wave", -1: "streaming", 0: "sequential"}.get(mode, f"chunked({mode})"),
        # different recurrence kernels (fp32 summation order): the log-probabilities agree to ~1e-6, and over 1000 frames of
        # near-uniform random-init output a near-tie in some beam flips for some utterances; each mode is bit-exact against the
        # oracle decoder on its OWN log-probabilities (tests)
        "logprob_max_abs_diff_vs_wave_engine": float(np.abs(pipe.logprobs() - logp_wave).max()),
        "transcripts_equal_wave_engine": f"{sum(a == b for a, b in zip(res_stream[0], res_wave[0]))} of 64",
        "parity": "tests/test_gpu_sizes.py::test_cfg2_streaming_latency_mode_vs_oracle"}
    pipe.close(); c2.free(xd); c2.close()

    # cfg3: 5-layer bidirectional GRU H=800, N=256, T=1000, beam 32, bf16 projection
    T, N, D, H, L, V, beam = 1000, 256, 161, 800, 5, 29, 32
    w = synth.rnn_weights(2, D, H, L, cell_gates=3, bidir=True)
    fc = synth.fc_weights(3, 2 * H, V)
    pipe = gasr.AsrPipeline(ctx, gasr.CELL_GRU, True, T, N, D, H, L, V, beam, 0, synth.VOCAB29, precision=gasr.PREC_BF16)
    pipe.set_weights(*w, *fc)
    xd = ctx.malloc(T * N * D * 4)
    ctx.synth_spectrogram(xd, 1, T, N, D)
    ms, _ = time_pipe(pipe, xd, 1)
    out["cfg3"] = {"shape": "5-layer bidirectional GRU H=800, N=256, T=1000, beam 32, bf16 projection", "ms": ms,
                   "rtfx": N * T * FRAME_SEC / (ms * 1e-3), "stages_ms": pipe.stage_times(),
                   "recurrence_hbm_frac": (N * T * 128000 / (max(pipe.stage_times()[1], 1e-9) * 1e-3) / 1e9) / measured_peaks()[0],
                   "parity": "tests/test_gpu_sizes.py::test_cfg3_bf16_mode_persistent_gru_recurrence (this mode's recurrence at N=256: 2e-2 vs the "
                             "fp32 oracle, 2e-3 vs the per-timestep kernel), test_cfg3_gru_at_batch_256 (fp32 mode, 1e-4), "
                             "test_gpu_parity.py::test_pipeline_cfg3_style_gru_bf16_projection (2e-2 + unchanged transcripts)"}
    pipe.close(); ctx.free(xd)

    # cfg4: decode only, T=4000, N=64, beam 8 / 32 / 128 (random-init-like log-probs); one utterance re-checked against the golden
    T, N, V = 4000, 64, 29
    lp = synth.random_logprobs(1, T, N, V)
    d = ctx.to_device(lp)
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "ctc_cfg4.json")))
    g_lp = synth.random_logprobs(77, T, 1, V)
    cfg4 = {}
    for beam in (8, 32, 128):
        best = 1e30
        for _ in range(2):
            ctx.sync(); ctx.timer_start()
            ctx.ctc_decode(d, gasr.DOMAIN_LOG, T, N, V, V, beam, 0, synth.VOCAB29)
            best = min(best, ctx.timer_stop())
        gp, gs = ctx.ctc_decode_host(g_lp, gasr.DOMAIN_LOG, beam, 0, synth.VOCAB29)
        case = [k for k in gold["cases"] if k["kind"] == "random" and k["beam"] == beam][0]
        cfg4[f"beam{beam}"] = {"ms": best, "us_per_frame": 1e3 * best / T, "rtfx": N * T * FRAME_SEC / (best * 1e-3),
                               "hbm_frac": (N * T * 116 / (best * 1e-3) / 1e9) / measured_peaks()[0],
                               "golden_bit_exact": bool(gp[0].hex() == case["path_hex"] and
                                                        int(np.float32(gs[0]).view(np.uint32)) == case["score_bits"])}
    out["cfg4"] = {"shape": "decode only, T=4000 N=64 V=29", **cfg4}
    ctx.free(d)
    ctx.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--utts", type=int, default=UTTS)
    ap.add_argument("--wave", type=int, default=0, help="utterances per batch (0: min(2048, this rank's share))")
    ap.add_argument("--lanes", type=int, default=LANES)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-checks", action="store_true", help="skip the oracle / gather checks (timing only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import gasr   # raises ImportError if libgasr.so is missing: no CPU fallback
    import shard
    import synth
    dist, host_group = None, None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        host_group = dist.new_group(backend="gloo")           # host-side gather of the results (no data-path collective)

    def note(msg):
        if os.environ.get("GASR_BENCH_VERBOSE"):
            print(f"[bench rank {rank}] {msg}", file=sys.stderr, flush=True)

    c = CFG
    # Contiguous shard of utterances per rank, cut into batches of `wave` utterances.  Rows of a batch never interact
    # (RNN.cu:15-27, CTCBeamSearch.cu:416), so every utterance's result is independent of the batch it travels in: ranks may
    # use different batch sizes and the gathered result still equals the single-GPU one bit for bit (checked below).
    assert args.utts % world == 0, "--utts must be a multiple of the number of ranks"
    share = args.utts // world
    if args.wave <= 0:
        args.wave = min(WAVE_MAX, share)
    assert share % args.wave == 0, "this rank's share must be a multiple of --wave"
    n_batches = args.utts // args.wave
    u_lo, u_hi = shard.shard_range(args.utts, world, rank)    # contiguous utterances of this rank
    b_lo, b_hi = u_lo // args.wave, u_hi // args.wave
    job = gasr.Job(local, c["T"], args.wave, c["D"], c["H"], c["L"], c["V"], c["beam"], 0, synth.VOCAB29, lanes=args.lanes)
    w, (fc_w, fc_b) = weights()
    job.set_weights(*w, fc_w, fc_b)
    note("job created, weights set")
    ctx0 = job.lane_context(0)
    batch_bytes = c["T"] * args.wave * c["D"] * 4

    def make_batch(b):
        d = ctx0.malloc(batch_bytes)
        ctx0.synth_spectrogram(d, SEED_X, c["T"], args.wave, c["D"], first_utt=b * args.wave)
        return d

    x_dev = [make_batch(b) for b in range(b_lo, b_hi)]
    numa_node = bind_host_memory_to_gpu_node(local)
    x_pin = []
    for d in x_dev:
        h = ctx0.pinned((c["T"] * args.wave, c["D"]))
        ctx0.d2h_into(h, d)
        x_pin.append(h)
    ctx0.sync()
    note("inputs generated")
    audio_per_step = args.utts * c["T"] * FRAME_SEC

    def barrier():
        ctx0.sync()
        if dist is not None:
            import torch
            torch.cuda.synchronize()
            dist.barrier()

    def timed(fn, steps):
        barrier()
        ctx0.timer_start()
        for _ in range(steps):
            fn()
        ms = ctx0.timer_stop()
        barrier()
        if dist is not None:
            import torch
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    ref = None
    for i in range(args.warmup):
        ref = job.run_device(x_dev)
        note(f"warm-up {i}: run_device done ({job.last_ms():.1f} ms)")
        job.run_host(x_pin)
        note(f"warm-up {i}: run_host done ({job.last_ms():.1f} ms)")

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = job.launch_count()
    ms_dev = timed(lambda: job.run_device(x_dev), args.steps)
    note(f"timed device-resident loop done: {ms_dev / args.steps:.1f} ms per step")
    launches = job.launch_count() - launches0
    ms_e2e = timed(lambda: job.run_host(x_pin), args.steps)
    clocks = sampler.stop() if rank == 0 else None
    last = job.run_device(x_dev)
    assert all((a == b).all() for a, b in zip(last[:2], ref[:2])) and (last[2].view(np.uint32) == ref[2].view(np.uint32)).all(), \
        "results changed between runs"

    # per-stage launch durations inside the running pipeline (profiled pass, outside the timed loops)
    job.profile(True)
    job.run_device(x_dev)
    stage_ms, stage_launches = job.stage_times()
    job.profile(False)

    # ---- gather (host side) and checks ------------------------------------------------------------------------------------
    gathered = None
    if dist is not None:
        bucket = [None] * world if rank == 0 else None
        dist.gather_object((last[0], last[1], last[2]), bucket, dst=0, group=host_group)
        if rank == 0:
            gathered = tuple(np.concatenate([b[i] for b in bucket]) for i in range(3))   # rank order == utterance order
    else:
        gathered = last

    if rank == 0:
        checks = {}
        if not args.no_checks:
            checks["parity_checked"] = oracle_check(gasr, job, x_dev[0], tuple(a[: args.wave] for a in last), 4)
            if world > 1:
                # the same 8192 utterances decoded by this GPU alone must equal the gathered N-GPU result bit for bit
                extra = [make_batch(b) for b in range(b_hi, n_batches)]
                alone = job.run_device(x_dev + extra)
                for d in extra:
                    ctx0.free(d)
                eq = bool((alone[0] == gathered[0]).all() and (alone[1] == gathered[1]).all()
                          and (alone[2].view(np.uint32) == gathered[2].view(np.uint32)).all())
                checks["gather"] = {"utterances": int(gathered[1].shape[0]), "ranks": world,
                                    "equals_single_gpu_result_bit_for_bit": eq}
            else:
                checks["gather"] = {"utterances": int(gathered[1].shape[0]), "ranks": 1, "equals_single_gpu_result_bit_for_bit": True,
                                    "note": "single rank: the result IS the single-GPU result"}
        hbm_peak, tc_peak, peak_kind = measured_peaks()
        local_rows = (b_hi - b_lo) * args.wave * c["T"]        # frame-utterances this rank processes per step
        n_rec = max(stage_launches[1], 1)
        rec_bytes_total = local_rows * c["H"] * 4 * 2 * c["L"]  # per layer: read xproj (fp32) + write h (4 B/element), SURVEY.md 8d
        rec_ms_per_launch = stage_ms[1] / n_rec
        achieved = rec_bytes_total / n_rec / (rec_ms_per_launch * 1e-3) / 1e9
        proj_flop = 2.0 * local_rows * (c["D"] * c["H"] + (c["L"] - 1) * c["H"] * c["H"])
        lin_bytes = local_rows * (c["H"] * 4 + c["V"] * 4)
        dec_bytes = local_rows * c["V"] * 4
        value = audio_per_step * args.steps / (ms_dev * 1e-3)
        e2e = audio_per_step * args.steps / (ms_e2e * 1e-3)
        step_ms = ms_dev / args.steps
        whole = local_rows * 25452 / (step_ms * 1e-3) / 1e9   # 25 452 compulsory bytes per frame and utterance

        def rl(bound, work, ms, peak, unit, scale):
            if ms <= 0:
                return None
            a = work / (ms * 1e-3) / scale
            return {"bound": bound, "achieved": a, "peak": peak, "unit": unit, "frac": a / peak, "ms_sum_of_launches": ms}

        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "utterances": args.utts, "utterances_per_batch": args.wave,
                       "batches_in_flight_per_gpu": args.lanes, "batches_per_gpu": b_hi - b_lo,
                       "pinned_host_memory_numa_node_rank0": numa_node,
                       "l2": "per-batch working set (x, xproj, hidden planes: ~16 MB per utterance) exceeds the 126 MB L2 many times over",
                       "init": "weights U(+-1/sqrt(H)) seed 4321, inputs U[0,1) seed 1234 (splitmix64, generated on the device, "
                               "bit-identical to synth.py)"},
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(args.utts * c["T"] * c["D"] * 4),
                    "d2h_bytes_per_step": int(args.utts * (c["T"] + 1 + 8))},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"kernel": "rnn_wide2_kernel (tcgen05 CTA-pair recurrence, one launch per layer and time chunk of a batch; "
                                   "durations measured inside the running pipeline, where it shares the GPU with the GEMM and decoder kernels)",
                         "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": rec_traffic(args.wave), "peak_source": peak_kind,
                         "algorithmic_bytes_per_launch": rec_bytes_total / n_rec, "ms_per_launch": rec_ms_per_launch,
                         "launches_per_step": n_rec},
            "whole_path_hbm": {"achieved": whole, "peak": hbm_peak, "unit": "GB/s", "frac": whole / hbm_peak,
                               "note": "25 452 compulsory bytes per frame and utterance (SURVEY.md 8d) over the whole step"},
            "pipeline": {"mode": "wave", "chunk_frames": 25 if 256 < args.wave <= 1024 else 50, "launches_per_stage_per_step": stage_launches,
                         "note": "stages overlap on separate streams; stage times are sums of per-launch durations (profiled pass) "
                                 "and exceed ms_per_step"},
            "stages_ms_sum_of_launches": {"projection_gemm": stage_ms[0], "recurrence": stage_ms[1],
                                          "linear_logsoftmax": stage_ms[2], "ctc_decode": stage_ms[3]},
            "stage_rooflines": {
                "projection_gemm": rl("tensor", proj_flop, stage_ms[0], tc_peak, "TFLOP/s", 1e12),
                "linear_logsoftmax": rl("hbm", lin_bytes, stage_ms[2], hbm_peak, "GB/s", 1e9),
                "ctc_decode": rl("hbm", dec_bytes, stage_ms[3], hbm_peak, "GB/s", 1e9),
            },
        }
        line.update(checks)
        if world == 1 and not args.no_checks:
            job.close()                                      # free the job's buffers before the other configs allocate theirs
            job = None
            # every stage ALONE: one batch through a pipeline whose stages all run on one stream (GASR_WAVE_SERIAL, read when the
            # context is created), per-launch CUDA events -- what a stage costs when it does not share the GPU
            os.environ["GASR_WAVE_SERIAL"] = "1"
            try:
                j2 = gasr.Job(local, c["T"], args.wave, c["D"], c["H"], c["L"], c["V"], c["beam"], 0, synth.VOCAB29, lanes=1)
            finally:
                del os.environ["GASR_WAVE_SERIAL"]
            j2.set_weights(*w, fc_w, fc_b)
            c2 = j2.lane_context(0)
            d2 = c2.malloc(batch_bytes)
            c2.synth_spectrogram(d2, SEED_X, c["T"], args.wave, c["D"], first_utt=0)
            j2.run_device([d2])
            j2.profile(True)
            j2.run_device([d2])
            alone_ms, alone_n = j2.stage_times()
            j2.close()
            nb = b_hi - b_lo
            rows1 = args.wave * c["T"]
            rec_alone = rows1 * c["H"] * 4 * 2 * c["L"] / (alone_ms[1] * 1e-3) / 1e9
            line["stages_alone_ms_per_step"] = {
                "projection_gemm": alone_ms[0] * nb, "recurrence": alone_ms[1] * nb, "linear_logsoftmax": alone_ms[2] * nb,
                "ctc_decode": alone_ms[3] * nb,
                "note": f"one batch of {args.wave} utterances with all stages on one stream, sums of per-launch durations x {nb} batches; "
                        "the recurrence launches occupy 64 of the 148 SMs, the other stages the whole GPU"}
            line["roofline"]["alone"] = {"achieved": rec_alone, "frac": rec_alone / hbm_peak, "ms_per_launch": alone_ms[1] / max(alone_n[1], 1),
                                         "note": "the same launches with nothing else on the GPU (64 SMs each)"}
            line["stage_rooflines_alone"] = {
                "projection_gemm": rl("tensor", proj_flop / nb, alone_ms[0], tc_peak, "TFLOP/s", 1e12),
                "linear_logsoftmax": rl("hbm", lin_bytes / nb, alone_ms[2], hbm_peak, "GB/s", 1e9),
                "ctc_decode": rl("hbm", dec_bytes / nb, alone_ms[3], hbm_peak, "GB/s", 1e9),
            }
            line["other_configs"] = other_configs(gasr, local)
        if world == 1 and not args.no_cpu_baseline:
            # bounded sample: batches of 64 utterances on all host threads until ~10 s of CPU work are spent
            cores = host_cores()
            n_utt, total_s, reps = 64, 0.0, 0
            while total_s < 10.0 and reps < 12:
                _, dt = cpu_port_rtfx(n_utt, cores, first_utt=reps * n_utt)
                total_s += dt
                reps += 1
            v = reps * n_utt * c["T"] * FRAME_SEC / total_s
            v1, _ = cpu_port_rtfx(2, 1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "seconds": total_s,
                                    "single_thread_value": v1,
                                    "sample": f"{reps} x 64 utterances x T={c['T']} of the {args.utts}, {cores} host threads, "
                                              f"oracle port (forward + CTC-REF decode); single thread: 2 utterances"}
        print(json.dumps(line), flush=True)

    if job is not None:
        for d in x_dev:
            ctx0.free(d)
        job.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

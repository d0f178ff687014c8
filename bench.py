#!/usr/bin/env python
"""bench.py -- RTFx of the hot path (RNN forward -> Linear -> log-softmax -> CTC beam search) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

A "step" is one pass of the whole path over one batch of BASELINE.json configs[1] ("cfg2": 64 utterances x
1000 frames x 161 bins, 3-layer tanh RNN H=512, Linear 512->29 + log-softmax, beam 16) on synthetic, seeded
inputs and torch-default-initialised random weights.  Every rank owns its own batch (weak scaling, utterances are
independent: no data-path collective); rank 0 prints ONE JSON line.

  value     RTFx with the batch already resident in HBM; transcripts + scores land in host memory inside the
            timed region.  Device-timed (CUDA events on the library's stream), max over ranks.
  e2e       the same through the host-buffer entry point gasr_asr_run_host: the batch is copied from pinned host
            memory every step (h2d_bytes_per_step) and the results are read back (d2h_bytes_per_step).
  roofline  the dominant kernel (the persistent recurrence kernel) against the measured HBM peak.
  cpu_baseline / --impl reference: the CPU restatement of the reference (oracle/, "port": the reference ships no
            runnable CPU implementation of this path -- CTCBeamSearch.cpp does not compile) on the
            box's host cores, on a bounded sample of the same workload.
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
CFG = dict(T=1000, N=64, D=161, H=512, L=3, V=29, beam=16)   # BASELINE.json configs[1]
SEED_X, SEED_W, SEED_FC = 1234, 4321, 99
FRAME_SEC = 0.010


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "measured"
    return 6650.0, 1400.0, "fallback"


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def workload(rank):
    import synth
    c = CFG
    x = synth.spectrogram_batch(SEED_X, c["T"], c["N"], c["D"], first_utt=rank * c["N"])
    w = synth.rnn_weights(SEED_W, c["D"], c["H"], c["L"])
    fc = synth.fc_weights(SEED_FC, c["H"], c["V"])
    return x, w, fc


def cpu_port_rtfx(n_utt, cores):
    """The oracle port (CPU restatement of the reference path) over n_utt utterances of the cfg2 workload."""
    import synth
    from oracle import oracle as O   # bench.py's cpu_baseline / reference leg: the one place it may run
    c = CFG
    x = synth.spectrogram_batch(SEED_X, c["T"], n_utt, c["D"])
    w = synth.rnn_weights(SEED_W, c["D"], c["H"], c["L"])
    fc_w, fc_b = synth.fc_weights(SEED_FC, c["H"], c["V"])
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
    sample = f"{n_utt} utterances x T={c['T']} of cfg2 per step (of 64), {cores} host threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2: 64 utt x T=1000 x D=161, 3-layer tanh RNN H=512, Linear 512->29 + log-softmax, "
                               "CTC beam 16", "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import gasr   # raises ImportError if libgasr.so is missing: no CPU fallback
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    c = CFG
    x, w, (fc_w, fc_b) = workload(rank)
    import synth
    ctx = gasr.Context(local)
    pipe = gasr.AsrPipeline(ctx, gasr.CELL_TANH, False, c["T"], c["N"], c["D"], c["H"], c["L"], c["V"], c["beam"], 0,
                            synth.VOCAB29)
    pipe.set_weights(*w, fc_w, fc_b)
    x_pinned = ctx.pinned(x.shape)
    x_pinned[...] = x
    x_dev = ctx.to_device(x)
    audio_per_step = c["N"] * c["T"] * FRAME_SEC

    def barrier():
        ctx.sync()
        if dist is not None:
            import torch
            torch.cuda.synchronize()
            dist.barrier()

    def timed(fn, steps):
        barrier()
        ctx.timer_start()
        for _ in range(steps):
            fn()
        ms = ctx.timer_stop()
        barrier()
        if dist is not None:
            import torch
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(args.warmup):
        ref = pipe.run_device(x_dev)
        pipe.run_host(x_pinned)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count()
    stage_acc = np.zeros(4)

    def step_dev():
        pipe.run_device(x_dev)
        stage_acc[:] += pipe.stage_times()

    ms_dev = timed(step_dev, args.steps)
    launches = ctx.launch_count() - launches0
    ms_e2e = timed(lambda: pipe.run_host(x_pinned), args.steps)
    clocks = sampler.stop() if rank == 0 else None
    last = pipe.run_device(x_dev)
    assert last == ref, "results changed between runs"

    if rank == 0:
        hbm_peak, tc_peak, peak_kind = measured_peaks()
        stage_ms = stage_acc / args.steps                      # proj, recurrence, linear, decode (per step)
        rows = c["T"] * c["N"]
        rec_bytes = rows * c["H"] * 4 * 2                      # per layer: read xproj (fp32) + write h (4 B/element)
        launches_per_stage, chunk_frames = pipe.stage_launches()
        streaming = chunk_frames < 0
        n_rec = max(launches_per_stage[1], 1)
        rec_bytes = rec_bytes * c["L"] // n_rec                # algorithmic bytes of ONE launch
        rec_ms_per_launch = stage_ms[1] / n_rec
        achieved = rec_bytes / (rec_ms_per_launch * 1e-3) / 1e9
        if streaming:
            rec_kernel = ("rnn_stream_kernel<512> (persistent recurrence of all 3 layers, ONE launch per step; its duration "
                          "includes waiting for the projection GEMM that runs concurrently)")
            mode = {"mode": "streaming", "note": "three persistent kernels (recurrence of all layers / tcgen05 projection + "
                    "output-layer GEMM / decoder) run concurrently for the whole sequence and hand 128-row blocks to each "
                    "other through counters in HBM; stage times are whole-kernel durations and overlap; Linear + "
                    "log-softmax is a target of the GEMM kernel", "launches_per_stage": launches_per_stage}
        else:
            rec_kernel = "rnn_tanh_mma_kernel (recurrence, one launch per layer and time chunk)"
            mode = {"mode": "chunked" if chunk_frames > 0 else "sequential", "chunk_frames": chunk_frames,
                    "launches_per_stage": launches_per_stage,
                    "note": "chunk_frames > 0: stages overlap on separate streams; stage times are sums of per-launch "
                            "durations and may exceed ms_per_step"}
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "rnn_stream_traffic.json")
        if streaming and os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")   # ncu --set full capture of the same kernel and shape
        proj_flop = 2.0 * rows * (c["D"] * c["H"] + (c["L"] - 1) * c["H"] * c["H"])
        if streaming:
            proj_flop += 2.0 * rows * c["H"] * c["V"]          # the output layer is a target of the same GEMM kernel
        lin_bytes = rows * (c["H"] * 4 + c["V"] * 4)
        dec_bytes = rows * c["V"] * 4
        value = world * audio_per_step * args.steps / (ms_dev * 1e-3)
        e2e = world * audio_per_step * args.steps / (ms_e2e * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg2: 64 utt x T=1000 x D=161, 3-layer tanh RNN H=512, Linear 512->29 + "
                                   "log-softmax, CTC beam 16 (per GPU)",
                       "l2": "per-step working set ~0.57 GB (x, xproj, 3 hidden sequences) exceeds the 126 MB L2",
                       "init": "weights U(+-1/sqrt(H)) seed 4321, inputs U[0,1) seed 1234 (splitmix64)"},
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(x.nbytes), "d2h_bytes_per_step": int(c["N"] * (c["T"] + 1 + 8))},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"kernel": rec_kernel, "bound": "hbm",
                         "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": traffic, "peak_source": peak_kind,
                         "algorithmic_bytes_per_launch": rec_bytes, "ms_per_launch": rec_ms_per_launch,
                         "launches_per_step": n_rec},
            "decode_prune": dict(zip(("fallback_utt_frames", "survivors_total"), ctx.ctc_last_stats())),
            "pipeline": mode,
            "stages_ms_per_step": {"projection_gemm": stage_ms[0], "recurrence": stage_ms[1],
                                   "linear_logsoftmax": stage_ms[2], "ctc_decode": stage_ms[3]},
            "stage_rooflines": {
                "projection_gemm": {"bound": "tensor", "achieved": proj_flop / (stage_ms[0] * 1e-3) / 1e12,
                                    "peak": tc_peak, "unit": "TFLOP/s",
                                    "frac": proj_flop / (stage_ms[0] * 1e-3) / 1e12 / tc_peak},
                "linear_logsoftmax": (None if stage_ms[2] <= 0 else
                                      {"bound": "hbm", "achieved": lin_bytes / (stage_ms[2] * 1e-3) / 1e9,
                                       "peak": hbm_peak, "unit": "GB/s",
                                       "frac": lin_bytes / (stage_ms[2] * 1e-3) / 1e9 / hbm_peak}),
                "ctc_decode": {"bound": "hbm", "achieved": dec_bytes / (stage_ms[3] * 1e-3) / 1e9, "peak": hbm_peak,
                               "unit": "GB/s", "frac": dec_bytes / (stage_ms[3] * 1e-3) / 1e9 / hbm_peak},
            },
        }
        if world == 1 and not args.no_cpu_baseline:
            # bounded sample: whole cfg2 batches (64 utterances) on all host threads until ~10 s of CPU work are spent
            cores = host_cores()
            n_utt, total_s, reps = c["N"], 0.0, 0
            while total_s < 10.0 and reps < 12:
                _, dt = cpu_port_rtfx(n_utt, cores)
                total_s += dt
                reps += 1
            v = reps * n_utt * c["T"] * FRAME_SEC / total_s
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "seconds": total_s,
                                    "sample": f"{reps} x the full cfg2 batch ({n_utt} utterances x T={c['T']}), "
                                              f"{cores} host threads, oracle port (forward + CTC-REF decode)"}
        print(json.dumps(line), flush=True)

    pipe.close()
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

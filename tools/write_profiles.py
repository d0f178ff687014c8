"""Turn the captures of tools/capture_profiles.sh (gpurun_out/r1_*) into the tracked summaries under profiles/."""
import csv, io, json, shutil, subprocess

def sh(c):
    return subprocess.run(c, shell=True, capture_output=True, text=True).stdout

for f in ("r1_launches_chunked.csv", "r1_bench_streaming.json", "r1_bench_chunked.json", "r1_config_sweep.txt"):
    shutil.copy(f"gpurun_out/{f}", f"profiles/{f}")
rows = list(csv.reader(io.StringIO(sh("ncu -i gpurun_out/r1_rnn_stream.ncu-rep --page raw --csv"))))
h, u, d = rows[0], rows[1], rows[2]
def val(name):
    i = h.index(name)
    return float(d[i]) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u[i], 1)
rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
json.dump({"kernel": "rnn_stream_kernel<512>",
           "shape": "T=1000 N=64 H=512 L=3 (bench.py workload), all layers in one launch, inputs preset (GASR_STREAM_DEBUG=3 GASR_DEBUG_REC_ALONE=1)",
           "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
           "algorithmic_bytes_per_launch": 786432000, "gpu_time_duration_ms": float(d[h.index("gpu__time_duration.sum")]),
           "source": "ncu --set full --clock-control none, report summarised in profiles/r1_rnn_stream.md"},
          open("profiles/rnn_stream_traffic.json", "w"), indent=1)
SO = "gpu-accelerated-speech-recognition_b200/libgasr.so"
SRC = "gpu-accelerated-speech-recognition_b200/csrc/"
for k, regex, hint, src in (("rnn_stream", "rnn_stream", ["rnn_stream_kernel", "Li512E"], "rnn_stream.cu"),
                            ("xproj_stream", "xproj_stream", ["xproj_stream"], "xproj_stream.cu"),
                            ("ctc_cta2", "ctc_beam_cta2", ["ctc_beam_cta2", "Li1ELi16ELi8E"], "ctc_beam.cu"),
                            ("ctc_general", "ctc_beam_kernel", ["ctc_beam_kernel", "Li1E", "Li1024E"], "ctc_beam.cu"),
                            ("gru_step", "gru_tc_step", ["gru_tc_step_kernel"], "gru_tc.cu")):
    summ = sh(f"python tools/ncu_summary.py gpurun_out/r1_{k}.ncu-rep")
    lines = sh(f"python tools/ncu_lines.py gpurun_out/r1_{k}.ncu-rep {regex} --hint {' '.join(hint)} --so {SO} --src {SRC}{src} --top 25")
    extra = ""
    if k == "ctc_cta2":
        extra = "\n## Samples by algorithm phase (`tools/phase_profile.py`; work vs waiting at the barrier that ends the phase)\n\n```\n" + \
                sh("python tools/phase_profile.py gpurun_out/r1_ctc_cta2.ncu-rep") + "```\n"
    open(f"profiles/r1_{k}.md", "w").write(
        f"# ncu --set full --clock-control none: `{regex}` (round 1)\n\nCaptured with `tools/capture_profiles.sh` on a B200 (kernel run "
        f"ALONE{' with its dependencies preset' if 'stream' in k else ''} -- see profiles/README.md).\n\n{summ}\n{extra}\n## Per-source-line warp-state samples "
        f"(top 25; `tools/ncu_lines.py`)\n\n```\n{lines}```\n")
chunked = json.loads(open("profiles/r1_bench_chunked.json").read().strip().splitlines()[-1])
with open("profiles/r1_launches_chunked.md", "w") as f:
    f.write("# bench.py launch list, chunked execution mode (GASR_STREAM=0)\n\n`GASR_STREAM=0 ncu --metrics gpu__time_duration.sum "
            "--clock-control none -c 600 --csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline`\n\n")
    f.write(sh("python tools/launch_summary.py profiles/r1_launches_chunked.csv"))
    f.write(f"\nbench.py line of the same mode without the profiler (profiles/r1_bench_chunked.json): ms_per_step = {chunked['ms_per_step']:.3f}\n")
s = json.loads(open("profiles/r1_bench_streaming.json").read().strip().splitlines()[-1])
print("streaming:", s["ms_per_step"], s["value"], s["e2e"], s["roofline"]["frac"], s["roofline"]["traffic"], s["cpu_baseline"]["value"])

"""Turn the captures of tools/r2/capture_profiles.sh (gpurun_out/r2_*) into the tracked summaries under profiles/."""
import shutil, subprocess, sys

def sh(c):
    return subprocess.run(c, shell=True, capture_output=True, text=True).stdout

SO = "gpu-accelerated-speech-recognition_b200/libgasr.so"
SRC = "gpu-accelerated-speech-recognition_b200/csrc/"
CMD = "python bench.py --utts 4096 --wave 4096 --lanes 1 --steps 1 --warmup 3 --no-cpu-baseline --no-checks"
shutil.copy("gpurun_out/r2_launches.csv", "profiles/r2_launches.csv")
with open("profiles/r2_launches.md", "w") as f:
    f.write("# bench.py launch list, wave engine, one batch of 4096 utterances (round 2)\n\n`ncu --metrics gpu__time_duration.sum "
            f"--clock-control none -s 1300 -c 400 --csv {CMD}`\n\nncu serialises the launches (cold caches, no overlap), so what must agree "
            "with bench.py is each kernel's SHARE of the step, not the absolute time.\n\n")
    f.write(sh("python tools/launch_summary.py profiles/r2_launches.csv"))
import csv, io, json, os
for k, regex, hint, src in (("rnn_wide2", "rnn_wide2_kernel", ["rnn_wide2_kernel"], "rnn_wide2.cu"),
                            ("gemm_pair", "gemm_pair_kernel", ["gemm_pair_kernel"], "gemm_pair.cu"),
                            ("xproj_stream", "xproj_stream_kernel", ["xproj_stream"], "xproj_stream.cu"),
                            ("ctc_warp", "ctc_beam_warp_kernel", ["ctc_beam_warp", "Li1ELi16E"], "ctc_beam.cu"),
                            ("gru_seq_ncu", "gru_seq_kernel", ["gru_seq_kernel", "Li16E"], "gru_seq.cu")):
    rep = f"gpurun_out/r2_{regex}.ncu-rep"
    if not os.path.exists(rep):
        print("missing", rep); continue
    if k == "rnn_wide2":
        rows = list(csv.reader(io.StringIO(sh(f"ncu -i {rep} --page raw --csv"))))
        h, u, d = rows[0], rows[1], rows[2]
        def val(name):
            i = h.index(name)
            return float(d[i]) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u[i], 1)
        rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
        json.dump({"kernel": "rnn_wide2_kernel", "utterances_per_launch": 4096, "steps_per_launch": 50,
                   "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
                   "algorithmic_bytes_per_launch": 50 * 4096 * 512 * 8, "gpu_time_duration_us": float(d[h.index("gpu__time_duration.sum")]),
                   "source": "ncu --set full --clock-control none inside the running bench command, summarised in profiles/r2_rnn_wide2.md"},
                  open("profiles/r2_rnn_wide2_traffic.json", "w"), indent=1)
    summ = sh(f"python tools/ncu_summary.py {rep}")
    lines = sh(f"python tools/ncu_lines.py {rep} {regex} --hint {' '.join(hint)} --so {SO} --src {SRC}{src} --top 25")
    open(f"profiles/r2_{k}.md", "w").write(
        f"# ncu --set full --clock-control none --import-source on: `{regex}` (round 2)\n\nCaptured with `tools/r2/capture_profiles.sh` "
        f"on a B200 inside `{CMD if 'gru' not in k else 'python tools/r2/cfg3_time.py 1 200'}` (no kernel of these paths waits for another kernel, so they run under ncu as they are).\n\n{summ}\n\n## Per-source-line warp-state samples (top 25; `tools/ncu_lines.py`)\n\n```\n{lines}```\n")
    print(k, len(summ), len(lines))

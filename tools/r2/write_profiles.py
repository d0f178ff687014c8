"""Turn the captures of tools/r2/capture_profiles.sh (gpurun_out/r2_*) into the tracked summaries under profiles/."""
import shutil, subprocess, sys

def sh(c):
    return subprocess.run(c, shell=True, capture_output=True, text=True).stdout

SO = "gpu-accelerated-speech-recognition_b200/libgasr.so"
SRC = "gpu-accelerated-speech-recognition_b200/csrc/"
CMD = "python bench.py --utts 2048 --wave 2048 --lanes 1 --steps 1 --warmup 3 --no-cpu-baseline --no-checks"
shutil.copy("gpurun_out/r2_launches.csv", "profiles/r2_launches.csv")
with open("profiles/r2_launches.md", "w") as f:
    f.write("# bench.py launch list, wave engine, one batch of 2048 utterances (round 2)\n\n`ncu --metrics gpu__time_duration.sum "
            f"--clock-control none -s 1300 -c 400 --csv {CMD}`\n\nncu serialises the launches (cold caches, no overlap), so what must agree "
            "with bench.py is each kernel's SHARE of the step, not the absolute time.\n\n")
    f.write(sh("python tools/launch_summary.py profiles/r2_launches.csv"))
for k, regex, hint, src in (("rnn_wide2", "rnn_wide2_kernel", ["rnn_wide2_kernel"], "rnn_wide2.cu"),
                            ("xproj_stream", "xproj_stream_kernel", ["xproj_stream"], "xproj_stream.cu"),
                            ("ctc_warp", "ctc_beam_warp_kernel", ["ctc_beam_warp", "Li1ELi16E"], "ctc_beam.cu")):
    rep = f"gpurun_out/r2_{regex}.ncu-rep"
    summ = sh(f"python tools/ncu_summary.py {rep}")
    lines = sh(f"python tools/ncu_lines.py {rep} {regex} --hint {' '.join(hint)} --so {SO} --src {SRC}{src} --top 25")
    open(f"profiles/r2_{k}.md", "w").write(
        f"# ncu --set full --clock-control none --import-source on: `{regex}` (round 2)\n\nCaptured with `tools/r2/capture_profiles.sh` "
        f"on a B200 inside `{CMD}` (the wave engine has no kernel that waits for another kernel, so it runs under ncu as it is; "
        f"launches 30-31 of that kernel).\n\n{summ}\n\n## Per-source-line warp-state samples (top 25; `tools/ncu_lines.py`)\n\n```\n{lines}```\n")
    print(k, len(summ), len(lines))

"""Per-source-line stall profile: joins `ncu --page source --csv` (per-SASS samples) with nvdisasm -gi line info.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep <kernel-regex> [--so path/to/lib.so] [--top 40]
"""
import argparse
import csv
import glob
import io
import os
import re
import subprocess
import tempfile


def sass_lines(so, mangled_hint):
    """[(sass_text, line)] in order for the kernel whose mangled name contains every token of mangled_hint."""
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, capture_output=True)
    out = []
    for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
        txt = subprocess.run(["nvdisasm", "-c", "-gi", cubin], capture_output=True, text=True).stdout
        cur, line, active = None, 0, False
        for ln in txt.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
            if m:
                active = all(tok in m.group(1) for tok in mangled_hint)
                if active and out:
                    return out
                continue
            if not active:
                continue
            m = re.search(r'//## File ".*?", line (\d+)', ln)
            if m:
                line = int(m.group(1))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                out.append((m.group(2).strip(), line))
        if out:
            return out
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("kernel")
    ap.add_argument("--so", default="gpu-accelerated-speech-recognition_b200/libgasr.so")
    ap.add_argument("--src", default=None)
    ap.add_argument("--hint", nargs="*", default=None, help="tokens of the mangled name (default: derived)")
    ap.add_argument("--top", type=int, default=40)
    a = ap.parse_args()
    txt = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv", "--kernel-name", f"regex:{a.kernel}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    # first kernel instance only
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    end = next((i for i in range(start + 1, len(rows)) if rows[i] and rows[i][0] == "Kernel Name"), len(rows))
    hdr, data = rows[start], rows[start + 1:end]
    ci = {h: j for j, h in enumerate(hdr)}
    kname = rows[start - 1][1] if start > 0 else a.kernel
    hint = a.hint or [a.kernel.split("|")[0]]
    sl = sass_lines(a.so, hint)
    assert len(sl) >= len(data), (len(sl), len(data))
    per = {}
    S, E = ci["# Samples"], ci["Instructions Executed"]
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = 0
    for (sass, line), r in zip(sl, data):
        s = int(r[S] or 0)
        tot += s
        d = per.setdefault(line, {"s": 0, "e": 0, "n": 0, "st": {}})
        d["s"] += s
        d["e"] += int(r[E] or 0)
        d["n"] += 1
        for h in stall_cols:
            v = int(r[ci[h]] or 0)
            if v:
                d["st"][h] = d["st"].get(h, 0) + v
    src = None
    srcfile = a.src
    if srcfile is None:
        cands = glob.glob("gpu-accelerated-speech-recognition_b200/csrc/*.cu")
        srcfile = next((c for c in cands if a.kernel.split("_")[0] in os.path.basename(c)), None)
    if srcfile and os.path.exists(srcfile):
        src = open(srcfile).read().splitlines()
    print(f"kernel: {kname}\ntotal samples {tot}, sass instrs {len(data)}, warp-instr executed {sum(d['e'] for d in per.values())}")
    for line, d in sorted(per.items(), key=lambda kv: -kv[1]["s"])[: a.top]:
        st = ", ".join(f"{k[6:]}={v}" for k, v in sorted(d["st"].items(), key=lambda kv: -kv[1])[:3])
        code = src[line - 1].strip()[:70] if src and 0 < line <= len(src) else ""
        print(f"L{line:4d} {100 * d['s'] / max(tot, 1):5.1f}%  exec={d['e']:9d} n={d['n']:3d}  [{st}]  {code}")


if __name__ == "__main__":
    main()

"""Aggregate an ncu source-level profile of the cta2 decode kernel by algorithm phase (work vs barrier-wait samples)."""
import csv, io, subprocess, sys
sys.path.insert(0, 'tools')
import ncu_lines
rep = sys.argv[1]
hint = sys.argv[2:] or ['ctc_beam_cta2', 'Li1ELi16ELi8E']
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:ctc_beam_cta2"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
start = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[start]; data = rows[start + 1:]
ci = {h: j for j, h in enumerate(hdr)}
sl = ncu_lines.sass_lines('gpu-accelerated-speech-recognition_b200/libgasr.so', hint)
src = open('gpu-accelerated-speech-recognition_b200/csrc/ctc_beam.cu').read().splitlines()
pats = (('fetch', 'fetch warp: log-probabilities'), ('trie', 'trie warp: one frame behind'), ('result', '---- result (CTCBeamSearch.cu:290-298): the aux'),
        ('main-init', '=============================== main warps'), ('B', 'phase B: merged candidates (+ probe'),
        ('C', 'phase C: lower bound of the beam-th largest merged key ===='), ('D', 'phase D: survivors ===='),
        ('E', 'phase E: exact order of the survivors, straight'), ('F', 'phase F: next beam: scores'))
k0 = next(i for i, l in enumerate(src, 1) if 'struct Cta2Beam' in l)
marks = sorted((i, name) for i, l in enumerate(src, 1) if i > k0 for name, pat in pats if pat in l)
def phase(line):
    p = 'pre'
    for ln, name in marks:
        if line >= ln: p = name
    return p
agg = {}
S = ci['# Samples']
for (sass, line), r in zip(sl, data):
    n = int(r[S] or 0)
    if not n: continue
    bar = int(r[ci['stall_barrier']] or 0) if 'stall_barrier' in ci else 0
    ph = phase(line) if line > k0 else 'helpers'
    d = agg.setdefault(ph, [0, 0]); d[0] += n - bar; d[1] += bar
tot = sum(v[0] + v[1] for v in agg.values())
for ph, v in agg.items():
    print(f"{ph:10s} work {v[0]:6d} ({100 * v[0] / tot:5.1f}%)  barrier-wait {v[1]:6d} ({100 * v[1] / tot:5.1f}%)")

"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv`) per kernel.

    python tools/launch_summary.py gpurun_out/launches.csv > profiles/rNN_launches.md
"""
import collections
import csv
import sys


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        k = r[4].split("(")[0].replace("void ", "")
        d = agg.setdefault(k, [0, 0.0, r[7], r[8]])
        d[0] += 1
        d[1] += float(r[14])
    tot = sum(v[1] for v in agg.values())
    print(f"launches captured: {len(rows)} (ncu serialises launches and runs them cold-cache; shares, not absolutes)\n")
    print("| kernel | launches | total us | avg us | share | block | grid |")
    print("|---|---|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {v[0]} | {v[1] / 1e3:.1f} | {v[1] / v[0] / 1e3:.2f} | {v[1] / tot * 100:.1f}% | {v[2]} | {v[3]} |")


if __name__ == "__main__":
    main(sys.argv[1])

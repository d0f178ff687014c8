"""SASS mnemonic counts per kernel of libgasr.so: python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "gpu-accelerated-speech-recognition_b200", "libgasr.so")
text = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pat = ["UTCHMMA", "LDTM", "UTMALDG", "UBLKCP", "UTCBAR", "UTCATOMSWS", "HMMA", "LDSM", "SYNCS", "UCGABAR", "MUFU", "SHFL", "BAR.SYNC"]
cnt, cur = collections.OrderedDict(), None
for line in text.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        cnt[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for p in pat:
        if re.search(r"\b" + re.escape(p), line):
            cnt[cur][p] += 1
names = subprocess.run(["c++filt"], input="\n".join(cnt.keys()), capture_output=True, text=True).stdout.splitlines()
print("# SASS mnemonic counts per kernel of libgasr.so (cuobjdump -sass), sm_100a")
print("# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UBLKCP = TMA, UTCBAR = tcgen05.commit, HMMA = mma.sync (legacy path),")
print("# SYNCS = mbarrier, UCGABAR = cluster barrier")
print("kernel | " + " | ".join(pat))
for (k, c), name in zip(cnt.items(), names):
    name = re.sub(r"\(.*", "", name).replace("void gasr::", "")
    if sum(c[p] for p in pat[:10]) == 0:
        continue
    print(name + " | " + " | ".join(str(c.get(p, 0)) for p in pat))

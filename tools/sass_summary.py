"""Counts of Blackwell-native SASS mnemonics per kernel of the in-tree library (evidence for profiles/): python tools/sass_summary.py"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "jamun_b200", "csrc", "libjamun_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
names = {"UTCHMMA": "UTCHMMA", "LDTM": "LDTM", "STTM": "STTM", "UBLKCP": "UBLKCP", "UBLKPF": "UBLKPF", "UTCBAR": "UTCBAR", "HMMA": "HMMA", "FFMA2": "FFMA2"}
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur:
        m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            if op in names:
                counts[cur][op] += 1
print("# SASS evidence for jamun_b200/csrc/libjamun_b200.so (built in-tree by __graft_entry__.build(), nvcc 12.9, -gencode arch=compute_100a,code=sm_100a)")
print("# counts of Blackwell-native mnemonics per kernel (cuobjdump -sass): UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk,")
print("# UBLKPF = cp.async.bulk.prefetch.L2, UTCBAR = tcgen05.commit, FFMA2 = packed fp32 FMA, HMMA = legacy mma.sync (must be 0)")
tot_h = 0
for k, c in counts.items():
    tot_h += c["HMMA"]
    if c["UTCHMMA"] or c["UBLKCP"] or c["UBLKPF"] or c["FFMA2"]:
        dem = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
        dem = re.sub(r"\(anonymous namespace\)::", "", dem)
        dem = dem.split("(")[0]
        print(f"{dem:70s} " + " ".join(f"{n}={c[n]}" for n in names if c[n] or n in ("UTCHMMA", "HMMA")))
print(f"total legacy HMMA instructions in the library: {tot_h}")

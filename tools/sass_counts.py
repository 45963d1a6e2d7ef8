"""Per-kernel counts of the SASS mnemonics that prove the Blackwell paths (tcgen05 = UTC*MMA, TMEM = LDTM / STTM, TMA bulk =
UBLKCP, tensor-map TMA = UTMALDG) from the built library.   python tools/sass_counts.py > profiles/r02_sass.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "chromoformer_b200", "libchromo_b200.so")
PAT = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "HMMA", "FFMA", "LDGSTS", "ATOMG", "REDG", "RED."]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
counts, name = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace("(anonymous namespace)::", "").replace("chromo::", "")
        name = re.sub(r"\(.*", "", name)
        counts[name] = collections.Counter()
        continue
    if name is None or not re.search(r"/\*[0-9a-f]{4,}\*/.*;", line):
        continue
    counts[name]["instructions"] += 1
    for p in PAT:
        if re.search(r"\b" + re.escape(p), line):
            counts[name][p] += 1
cols = ["instructions"] + [p for p in PAT if any(c[p] for c in counts.values())]
print("SASS mnemonic counts per kernel of chromoformer_b200/libchromo_b200.so (cuobjdump -sass, sm_100a)")
print("UTCHMMA = tcgen05.mma kind::f16 (BF16), LDTM / STTM = tcgen05.ld / .st (TMEM), UBLKCP = cp.async.bulk (1-D TMA),")
print("UTMALDG = tensor-map TMA (not used: every operand this library copies is a pre-packed contiguous tile),")
print("UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, REDG/ATOMG = global atomics.\n")
w = max(len(n) for n in counts) + 2
print("kernel".ljust(w) + "".join(c.rjust(14) for c in cols))
for n, c in sorted(counts.items(), key=lambda kv: -kv[1]["UTCHMMA"] * 10**6 - kv[1]["instructions"]):
    print(n.ljust(w) + "".join(str(c[k]).rjust(14) for k in cols))

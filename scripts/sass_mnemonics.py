"""Static SASS mnemonic counts per kernel of the shipped library (what proves which kernels are Blackwell-native).
usage: python scripts/sass_mnemonics.py > profiles/r02_sass_mnemonics.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gcgcn_b200", "libgcgcn_b200.so")
WATCH = [("UTC", r"\bUTC[A-Z]*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UBLKCP", r"\bUBLKCP"), ("UTMALDG", r"\bUTMALDG"),
         ("UTMASTG", r"\bUTMASTG"), ("LDGSTS", r"\bLDGSTS"), ("HMMA", r"\bHMMA"), ("LDSM", r"\bLDSM"), ("SYNCS", r"\bSYNCS"),
         ("MUFU", r"\bMUFU")]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
name = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.splitlines()
counts, cur, i = collections.OrderedDict(), None, -1
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        i += 1
        cur = re.sub(r"\(.*", "", name[i])
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for key, pat in WATCH:
        if re.search(pat, line):
            counts[cur][key] += 1
print("# SASS mnemonics per kernel of gcgcn_b200/libgcgcn_b200.so (cuobjdump -sass, sm_100a) -- counts of static instructions")
print("# UTC = tcgen05.mma (UTC*MMA), LDTM / STTM = tcgen05.ld / tcgen05.st, UBLKCP = cp.async.bulk (bulk copy engine), UTMALDG/UTMASTG =")
print("# cp.async.bulk.tensor (none: activation operands are split hi/lo by producer warps, weights travel as pre-split blobs by")
print("# UBLKCP), LDGSTS = cp.async, HMMA = mma.sync, LDSM = ldmatrix, SYNCS = mbarrier ops, MUFU = special-function unit\n")
for k, c in counts.items():
    if c:
        print(f"{k:96s} " + "  ".join(f"{a}:{c[a]}" for a, _ in WATCH if c[a]))

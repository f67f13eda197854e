"""SASS mnemonics that prove the Blackwell-specific paths, per kernel of rlmd_b200/librlmd_b200.so
(cuobjdump -sass; runs on the CPU box):  python tools/sass_evidence.py > profiles/rNN_sass_evidence.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "rlmd_b200", "librlmd_b200.so")
WATCH = {
    "UTMALDG": "TMA tile load (cp.async.bulk.tensor)", "UBLKCP": "TMA bulk copy", "SYNCS": "mbarrier arrive / try_wait",
    "UCGABAR": "thread-block-cluster barrier", "FMNMX3": "three-input fp32 min/max (sm_100)", "REDUX": "warp reduce (REDUX)",
    "ATOMS": "shared-memory atomic", "ATOM": "global / distributed-shared atomic", "RED": "global reduction (no return)",
    "MUFU": "XU pipe (lg2 / sin / cos / sqrt / ex2)", "POPC": "population count", "IMAD.WIDE": "32x32->64 multiply (Philox)",
    "FFMA2": "packed fp32 FMA", "FMUL2": "packed fp32 multiply", "LDGSTS": "cp.async", "MATCH": "warp match",
    "HMMA": "legacy tensor path (must be 0)", "UTCHMMA": "tcgen05.mma (none: no dense contraction on this path)",
}
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
print("library:", os.path.relpath(so, ROOT), " cubin architectures:", arch)
kern, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        for w in sorted(WATCH, key=len, reverse=True):      # longest prefix first (ATOMS before ATOM, REDUX before RED)
            if op.startswith(w):
                counts[kern][w] += 1
                break
demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print("\nmnemonic legend:")
for w, what in WATCH.items():
    print(f"  {w:10s} {what}")
print("\nper kernel (only kernels that show at least one watched mnemonic; counts of static SASS instructions):")
tot = collections.Counter()
for (k, c), name in zip(counts.items(), demangle):
    tot.update(c)
    if c:
        short = re.sub(r"\(.*", "", name)
        print(f"  {short[:86]:86s} " + " ".join(f"{w}={n}" for w, n in sorted(c.items())))
print("\ntotals:", " ".join(f"{w}={tot[w]}" for w in WATCH))

"""per-kernel SASS mnemonic counts of libaxisym_b200.so (cuobjdump -sass), the evidence for which hardware paths a
kernel uses:  python tools/sass_summary.py > profiles/rNN_sass_summary.txt
UTMALDG / UTMASTG = TMA tensor loads / stores, UBLKCP = bulk async copy, SYNCS = mbarrier ops, LDGSTS = cp.async,
DMMA = FP64 tensor-core MMA, DFMA/DADD/DMUL = FP64 pipe, SHFL = warp shuffles, BAR = CTA barriers, ATOM/RED = atomics."""
import collections
import os
import re
import subprocess
import sys

so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pyaxisymflow_b200", "libaxisym_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
raw = re.findall(r"Function : (\S+)", out)
# internal-linkage kernels are emitted as __nv_static_NN__<hash>_<file>__<mangled>: demangle the tail
raw = [re.sub(r"^__nv_static_\d+__\w+?_(_Z)", r"\1", n) for n in raw]
names = subprocess.run(["c++filt"], input="\n".join(raw), capture_output=True, text=True).stdout.splitlines()
MN = ["UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDGSTS", "DMMA", "DFMA", "DADD", "DMUL", "MUFU", "SHFL", "LDS", "STS",
      "LDG", "STG", "BAR", "ATOM", "RED", "UTCMMA", "LDTM"]
counts, cur, i = collections.OrderedDict(), None, 0
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        n = names[i]
        i += 1
        n = re.sub(r"\(anonymous namespace\)::", "", n)
        n = re.sub(r"^void ", "", n)
        cur = re.sub(r"\(GridD.*|\(CUtensorMap.*|\(int.*|\(double.*|\(unsigned.*|\(short.*|\(const.*", "", n)[:56]
        counts.setdefault(cur, collections.Counter())
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["_n"] += 1
        for k in MN:
            if op.startswith(k):
                counts[cur][k] += 1
                break
want = sys.argv[1:]
print(f"{'kernel':56s} {'instr':>6s} " + " ".join(f"{k[:6]:>6s}" for k in MN))
for k, c in counts.items():
    if want and not any(w in k for w in want):
        continue
    print(f"{k:56s} {c['_n']:6d} " + " ".join(f"{c[m]:6d}" if c[m] else "     ." for m in MN))
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print(f"{'TOTAL (' + str(len(counts)) + ' kernels)':56s} {tot['_n']:6d} " + " ".join(f"{tot[m]:6d}" for m in MN))

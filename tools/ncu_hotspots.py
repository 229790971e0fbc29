"""Per-source-line hot spots of one kernel launch in an `ncu --set full --import-source on` report:
     python tools/ncu_hotspots.py report.ncu-rep <kernel regex> [launch index] [top N]
ncu's CSV source page is per SASS instruction; the line numbers come from nvdisasm --print-line-info of the SAME
libaxisym_b200.so (built with -lineinfo), joined by instruction offset.  Columns: warp-stall samples, shared-memory
wavefronts (excess over the conflict-free count), instructions executed."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, pat = sys.argv[1], sys.argv[2]
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
so = os.environ.get("AXB_SO", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pyaxisymflow_b200",
                                           "libaxisym_b200.so"))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat, "--launch-skip",
                      str(skip), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kname = rows[0][1]
h = rows[1]
ix = {n: i for i, n in enumerate(h)}
sass = []
for r in rows[2:]:
    try:
        sass.append((int(r[ix["Address"]], 16), r[ix["Source"]].strip(), int(r[ix["Warp Stall Sampling (All Samples)"]]),
                     int(r[ix["L1 Wavefronts Shared"]]), int(r[ix["L1 Wavefronts Shared Excessive"]]),
                     int(r[ix["Instructions Executed"]])))
    except (ValueError, IndexError):
        pass
base = sass[0][0]
# ---- line info of the same kernel from the library
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.splitlines()
short = re.sub(r"^void |<unnamed>::|\(anonymous namespace\)::", "", kname).split("(")[0]       # k_dct_rows_w<(bool)1>
m = re.match(r"(\w+)(?:<(.*)>)?", short)
fn = m.group(1)
targs = re.findall(r"\)(-?\d+)", m.group(2) or "")                                               # template values
want = re.compile(r"\.text\..*\d+" + fn + "I" + "".join(r"L\w" + (a.replace("-", "n")) + "E" for a in targs)) if targs \
    else re.compile(r"\.text\..*\d+" + fn + r"(E|I)")
line_of, cur, on = {}, None, False
for ln in dis:
    if ln.lstrip().startswith("//--------------------- .text."):
        on = bool(want.search(ln))
        cur = None
        continue
    if not on:
        continue
    mm = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if mm and "inlined at" not in ln:
        cur = (os.path.basename(mm.group(1)), int(mm.group(2)))
    mm = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+\S", ln)
    if mm and cur:
        line_of[int(mm.group(1), 16)] = cur
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
for addr, src, st, wf, exc, ex in sass:
    a = agg[line_of.get(addr - base, ("?", 0))]
    a[0] += st; a[1] += wf; a[2] += exc; a[3] += ex
tot = [sum(v[i] for v in agg.values()) for i in range(4)]
print(f"{kname[:90]}\n{len(sass)} SASS instructions, {len(line_of)} with line info; totals: stall samples {tot[0]}, shared wavefronts "
      f"{tot[1]} (excess {tot[2]}), warp instructions {tot[3]}")
src_cache = {}
def text(f, n):
    if f not in src_cache:
        p = os.path.join(os.path.dirname(so), "csrc", f)
        src_cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
    L = src_cache[f]
    return L[n - 1].strip()[:90] if 0 < n <= len(L) else ""
print(f"{'samples':>8s} {'%':>5s} {'wavefronts':>11s} {'excess':>9s} {'instr':>9s}  line")
for (f, n), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{v[0]:8d} {100.0 * v[0] / max(tot[0], 1):5.1f} {v[1]:11d} {v[2]:9d} {v[3]:9d}  {f}:{n}  {text(f, n)}")

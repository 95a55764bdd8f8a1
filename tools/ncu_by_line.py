"""Join an .ncu-rep's per-SASS-instruction counters with nvdisasm's line table of the same build and aggregate by
source line (development tool).

    python tools/ncu_by_line.py REPORT.ncu-rep CUBIN_NAME KERNEL_MANGLED_SUBSTR [top]

CUBIN_NAME: e.g. ne_flux_tab2.sm_100a.cubin (extracted from libne_b200.so with cuobjdump -xelf).
The report and the library must come from the same build: instruction offsets are matched one to one."""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

rep, cubin, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
prefer = sys.argv[5] if len(sys.argv) > 5 else None   # attribute to the innermost inline frame in this file (basename)
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "numericalearth.jl_b200", "libne_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", cubin, so], cwd=tmp, check=True, capture_output=True)
sass = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()

# offset -> (file, line) for the chosen kernel
loc = {}
inside = False
cur = ("?", 0)
order = []
chain, fresh = [], True
for ln in sass:
    if ln.startswith(".text."):
        inside = kern in ln
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        frame = (os.path.basename(m.group(1)), int(m.group(2)))
        if fresh:
            chain = []
            fresh = False
        chain.append(frame)
        cur = next((f for f in chain if f[0] == prefer), chain[0]) if prefer else chain[0]
        continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", ln)
    if m:
        off = int(m.group(1), 16)
        loc[off] = (cur, m.group(2).strip())
        order.append(off)
        fresh = True

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ci = {n: hdr.index(n) for n in ("Address", "Source", "Instructions Executed", "Warp Stall Sampling (All Samples)",
                                "L1 Wavefronts Shared", "stall_long_sb", "stall_wait", "stall_short_sb", "stall_math")}
body = [r for r in rows[h + 1:] if len(r) > ci["stall_math"] and r[0].startswith("0x")]
base = int(body[0][0], 16)
by_line = defaultdict(lambda: [0, 0, 0, 0, 0, 0, 0])
tot = [0, 0, 0]
mism = 0
for r in body:
    off = int(r[0], 16) - base
    n = int(r[ci["Instructions Executed"]] or 0)
    s = int(r[ci["Warp Stall Sampling (All Samples)"]] or 0)
    w = int(r[ci["L1 Wavefronts Shared"]] or 0)
    key, text = loc.get(off, (("?", 0), ""))
    if text and text.split()[0].lstrip("@!UP0123456789 ") and r[ci["Source"]].split()[0:1] != text.split()[0:1]:
        mism += 1
    a = by_line[key]
    a[0] += n; a[1] += s; a[2] += w
    a[3] += int(r[ci["stall_long_sb"]] or 0); a[4] += int(r[ci["stall_wait"]] or 0)
    a[5] += int(r[ci["stall_short_sb"]] or 0); a[6] += int(r[ci["stall_math"]] or 0)
    tot[0] += n; tot[1] += s; tot[2] += w
print(f"instructions {tot[0]}, samples {tot[1]}, shared wavefronts {tot[2]}; opcode mismatches vs this build: {mism} of {len(body)}")
print(f"{'file:line':34s} {'inst %':>7s} {'samp %':>7s} {'lds wf %':>8s} {'long':>6s} {'wait':>6s} {'short':>6s} {'math':>6s}")
for key, a in sorted(by_line.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{key[0] + ':' + str(key[1]):34s} {100 * a[0] / tot[0]:7.2f} {100 * a[1] / max(tot[1], 1):7.2f} {100 * a[2] / max(tot[2], 1):8.2f} "
          f"{a[3]:6d} {a[4]:6d} {a[5]:6d} {a[6]:6d}")

"""Groups the SASS instructions of an ncu `--page source --csv` export by how often they executed: a kernel's hot loop,
its rarely taken paths and its prologue show up as plateaus.  python tools/ncu_hot.py gpurun_out/prof_<tag>.src.csv [top]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, iex, ist = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ins = [(r[isrc].strip(), int(r[iex]), int(r[ist])) for r in rows[2:] if len(r) > iex and r[iex].isdigit()]
total = sum(e for _, e, _ in ins)
samples = sum(s for _, _, s in ins)
print(f"{len(ins)} SASS instructions, {total} warp-instructions executed, {samples} stall samples")
# plateaus: bucket by executed count rounded to 2 significant digits
buckets = defaultdict(lambda: [0, 0, 0, defaultdict(int)])
for src, e, s in ins:
    if e == 0:
        continue
    key = float(f"{e:.2g}")
    b = buckets[key]
    b[0] += 1
    b[1] += e
    b[2] += s
    b[3][src.split()[0] if not src.startswith("@") else src.split()[1]] += 1
for key in sorted(buckets, key=lambda k: -buckets[k][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 12]:
    n, e, s, ops = buckets[key]
    top = ", ".join(f"{k} {v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:8])
    print(f"  executed ~{key:12.4g} x : {n:4d} instr = {100 * e / total:5.1f} % of executed, {100 * s / max(samples, 1):5.1f} % of samples | {top}")

"""Summarise an `ncu --page source --csv` (SASS) export: executed instructions and stall samples by opcode, and the
hottest address ranges.  usage: ncu_sass.py file.csv[.gz] [warp_coeff_steps]"""
import collections
import csv
import gzip
import io
import sys

path = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else None
f = io.TextIOWrapper(gzip.open(path)) if path.endswith(".gz") else open(path)
rows = list(csv.reader(f))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
h = rows[hi]
ia, isrc, iex, ismp = h.index("Address"), h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
ops, smp = collections.Counter(), collections.Counter()
seq = []
for r in rows[hi + 1:]:
    if len(r) <= max(iex, ismp):
        continue
    try:
        ex, sm = int(r[iex].replace(",", "")), int(r[ismp].replace(",", "") or 0)
    except ValueError:
        continue
    src = r[isrc].strip()
    parts = src.split()
    op = parts[1] if parts and parts[0].startswith("@") and len(parts) > 1 else (parts[0] if parts else "?")
    op = ".".join(op.split(".")[:3])
    ops[op] += ex
    smp[op] += sm
    seq.append((r[ia], src, ex, sm))
tot, tots = sum(ops.values()), sum(smp.values())
print(f"static instructions {len(seq)}, executed warp-instructions {tot}" + (f", per warp-step {tot / steps:.0f}" if steps else ""))
print("top opcodes by executed count (share of executed, share of stall samples):")
for op, c in ops.most_common(22):
    print(f"  {op:28s} {c / tot * 100:5.1f}%  samples {smp[op] / max(tots, 1) * 100:5.1f}%" + (f"  per step {c / steps:7.1f}" if steps else ""))
# hottest 64-instruction windows by samples
W = 48
best = []
for i in range(0, len(seq) - W, W):
    s = sum(x[3] for x in seq[i:i + W])
    e = sum(x[2] for x in seq[i:i + W])
    best.append((s, e, i))
best.sort(reverse=True)
print("hottest windows by stall samples:")
for s, e, i in best[:8]:
    print(f"  @{i:5d} samples {s / max(tots, 1) * 100:5.1f}% executed {e / tot * 100:5.1f}%   e.g. {seq[i][1][:60]} | {seq[i + W // 2][1][:60]}")

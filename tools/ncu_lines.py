"""Summarise an `ncu --page source --print-source cuda,sass --csv` export by CUDA source line: executed warp
instructions, stall samples and the long-scoreboard / no-instruction samples.  usage: ncu_lines.py file.csv[.gz] [top]"""
import collections
import csv
import gzip
import io
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
f = io.TextIOWrapper(gzip.open(path)) if path.endswith(".gz") else open(path)
cur_file, hdr = "?", None
lines = collections.OrderedDict()
for r in csv.reader(f):
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        iex, ismp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        ilsb, inoi = hdr.index("stall_long_sb"), hdr.index("stall_no_inst")
        continue
    if hdr is None or len(r) <= max(iex, ismp, ilsb, inoi) or not r[0]:
        continue
    try:
        vals = [int(r[i].replace(",", "") or 0) for i in (iex, ismp, ilsb, inoi)]
    except ValueError:
        continue
    key = (cur_file, int(r[0]))
    e = lines.setdefault(key, [0, 0, 0, 0, r[1].strip()[:90]])
    for i in range(4):
        e[i] += vals[i]
tot = [sum(e[i] for e in lines.values()) for i in range(4)]
print(f"executed {tot[0]}, samples {tot[1]}, long_sb {tot[2]}, no_inst {tot[3]}")
print("by executed instructions:")
for (fn, ln), e in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"  {fn}:{ln:5d} exec {e[0] / tot[0] * 100:5.1f}% smp {e[1] / max(tot[1], 1) * 100:5.1f}% lsb {e[2] / max(tot[2], 1) * 100:5.1f}% noi {e[3] / max(tot[3], 1) * 100:5.1f}%  {e[4]}")
print("by stall samples:")
for (fn, ln), e in sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"  {fn}:{ln:5d} exec {e[0] / tot[0] * 100:5.1f}% smp {e[1] / max(tot[1], 1) * 100:5.1f}% lsb {e[2] / max(tot[2], 1) * 100:5.1f}% noi {e[3] / max(tot[3], 1) * 100:5.1f}%  {e[4]}")

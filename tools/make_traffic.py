"""profiles/traffic_r02.json from the raw ncu exports of tools/gpu_r02g.sh / gpu_r02x.sh (argument: the file prefix): DRAM bytes read + written per launch of the
dominant kernel of every bench.py config, next to the sizes they were captured at (bench.py's roofline.traffic reads it
and reports null when its own sizes differ)."""
import csv
import json
import sys

CFG = {1: {"batch": 65536, "terms": 2000}, 2: {"batch": 16384, "terms": 256, "calls": 1}, 3: {"terms": 1000, "tile": 131072},
       4: {"terms": 3504, "points": 1000000}}
LAUNCHES_PER_STEP = {1: 1, 2: 1, 3: 8, 4: 1}  # launches of that kernel in one bench step (config 3: 8 tiles on one GPU)


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


PREFIX = sys.argv[1] if len(sys.argv) > 1 else "r02g"
out = {}
for cfg in (1, 2, 3, 4):
    path = f"gpurun_out/{PREFIX}_ncu_cfg{cfg}_raw.csv"
    try:
        rows = list(csv.reader(open(path)))
    except OSError:
        continue
    hdr, units, vals = rows[0], rows[1], rows[-1]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    rd, wr = to_bytes(*d["dram__bytes_read.sum"]), to_bytes(*d["dram__bytes_write.sum"])
    dur = d["gpu__time_duration.sum"]
    out[str(cfg)] = {"kernel": d["Kernel Name"][0], "config": CFG[cfg], "dram_bytes_read_per_launch": rd,
                     "dram_bytes_write_per_launch": wr, "launches_per_step": LAUNCHES_PER_STEP[cfg],
                     "dram_bytes_per_step": (rd + wr) * LAUNCHES_PER_STEP[cfg],
                     "duration_under_ncu": f"{dur[0]} {dur[1]}", "source": path.replace("gpurun_out/", "profiles/")}
json.dump(out, open("profiles/traffic_r02.json", "w"), indent=1)
print(json.dumps(out, indent=1))

"""One step out of an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch
list: the window between two consecutive launches of a marker kernel, aggregated by kernel name with device time and
DRAM bytes.
usage: python scripts/summarize_step.py launches.csv <marker substring> [which occurrence, default last-but-one] [title]
prints markdown; with --json also writes {kernel: {us, dram_bytes, launches}} + totals to the path given."""
import collections
import csv
import json
import re
import sys

args = [a for a in sys.argv[1:] if not a.startswith("--json")]
jpath = next((a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--json=")), None)
path, marker = args[0], args[1]
occ = int(args[2]) if len(args) > 2 else -2
title = args[3] if len(args) > 3 else path
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
launch = collections.OrderedDict()
for r in csv.DictReader(lines):
    i = int(r["ID"])
    d = launch.setdefault(i, {"name": r["Kernel Name"], "grid": r["Grid Size"], "block": r["Block Size"], "us": 0.0, "rd": 0.0, "wr": 0.0})
    v = float(r["Metric Value"].replace(",", ""))
    unit, m = r["Metric Unit"], r["Metric Name"]
    if m == "gpu__time_duration.sum":
        d["us"] = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    else:
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        d["rd" if "read" in m else "wr"] = v * scale
ids = list(launch)
marks = [i for i in ids if marker in launch[i]["name"]]
lo, hi = marks[occ], marks[occ + 1]
rows = [launch[i] for i in ids if lo <= i < hi]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*", "", name)
    return name if len(name) < 84 else name[:81] + "..."


agg = collections.OrderedDict()
for r in rows:
    a = agg.setdefault(short(r["name"]), {"n": 0, "us": 0.0, "bytes": 0.0, "grid": r["grid"], "block": r["block"]})
    a["n"] += 1
    a["us"] += r["us"]
    a["bytes"] += r["rd"] + r["wr"]
tot_us = sum(a["us"] for a in agg.values())
tot_b = sum(a["bytes"] for a in agg.values())
mine = sum(a["us"] for k, a in agg.items() if k.startswith("gcgcn::"))
print(f"# {title}\n")
print(f"launches {lo}..{hi - 1} of `{path}` (one step, between two launches of `{marker}`): {len(rows)} launches, "
      f"serialised device time {tot_us / 1e3:.3f} ms (cold-cache, one kernel at a time: compare shares, not absolutes), "
      f"DRAM traffic {tot_b / 1e9:.3f} GB (dram__bytes_read.sum + dram__bytes_write.sum), gcgcn_b200 kernels "
      f"{100 * mine / max(tot_us, 1e-9):.1f} % of the device time\n")
print("| kernel | launches | total us | share | DRAM MB | GB/s | grid | block |")
print("|---|---:|---:|---:|---:|---:|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    print(f"| `{k}` | {a['n']} | {a['us']:.1f} | {100 * a['us'] / tot_us:.1f} % | {a['bytes'] / 1e6:.1f} | "
          f"{a['bytes'] / max(a['us'], 1e-9) / 1e3:.0f} | {a['grid']} | {a['block']} |")
if jpath:
    with open(jpath, "w") as f:
        json.dump({"source": path, "window": [lo, hi - 1], "total_us": tot_us, "dram_bytes": tot_b,
                   "kernels": {k: {"launches": a["n"], "us": a["us"], "dram_bytes": a["bytes"]} for k, a in agg.items()}}, f, indent=1)

"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python scripts/summarize_launches.py gpurun_out/launches.csv [skip_first_n] > profiles/<name>.md"""
import csv, sys, re, collections
path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^void ", "", name)
    rows.append((int(r["ID"]), name, us, r["Grid Size"], r["Block Size"]))
rows = rows[skip:]
agg = collections.OrderedDict()
for _, name, us, grid, block in rows:
    a = agg.setdefault(name, [0, 0.0, grid, block])
    a[0] += 1; a[1] += us
total = sum(a[1] for a in agg.values())
mine = sum(a[1] for k, a in agg.items() if k.startswith("gcgcn::"))
print(f"# ncu launch list summary: {path}\n")
print(f"launches: {len(rows)}   total device time: {total/1e3:.3f} ms   gcgcn_b200 kernels: {mine/1e3:.3f} ms ({100*mine/max(total,1e-9):.1f} %)\n")
print("| kernel | launches | total us | share | avg us | grid | block |")
print("|---|---:|---:|---:|---:|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    short = k if len(k) < 90 else k[:87] + "..."
    print(f"| `{short}` | {a[0]} | {a[1]:.1f} | {100*a[1]/total:.1f} % | {a[1]/a[0]:.1f} | {a[2]} | {a[3]} |")

"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: the last N launches (the measured step of
tools/one_step.py) grouped by kernel.  usage: launch_summary.py launches.csv N [--list]; N = 0 takes the launches between the last two occurrences of
the step's first kernel (the stem forward), i.e. exactly one steady-state step including the ATen kernels of the loss."""
import csv, sys, re, collections

path, n = sys.argv[1], int(sys.argv[2])
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    rows.append((name, us, r["Grid Size"], r["Block Size"]))
if n == 0:
    marks = [i for i, r in enumerate(rows) if "stem_fwd" in r[0] or "smallcin_fwd" in r[0]]
    rows = rows[marks[-2]:marks[-1]]
else:
    rows = rows[-n:]
tot = sum(r[1] for r in rows)
print(f"# launches in the measured step: {len(rows)}   sum of durations: {tot / 1000:.3f} ms")
if "--list" in sys.argv:
    for name, us, g, b in rows:
        print(f"{us:9.1f} us  {name[:70]:70s} grid {g} block {b}")
    sys.exit(0)
agg = collections.OrderedDict()
for name, us, _, _ in rows:
    a = agg.setdefault(name, [0.0, 0])
    a[0] += us
    a[1] += 1
for name, (us, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{us:10.1f} us {100 * us / tot:5.1f}% x{c:3d}  {name[:100]}")

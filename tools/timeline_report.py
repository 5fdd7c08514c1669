"""Exposed-time table of a kernel timeline CSV (tools/timeline.py): the step is cut at the stem forward kernels; every instant of
the middle step is attributed to the kernels running at that instant (split evenly when several overlap), idle time separately.
usage: timeline_report.py timeline.csv [--list]"""
import sys, csv, re, collections

rows = []
with open(sys.argv[1]) as f:
    for r in csv.DictReader(f):
        rows.append((float(r["ts_us"]), float(r["dur_us"]), r["stream"], re.sub(r"\(.*", "", r["name"]).replace("void ", "").replace("rehr::", "")))
rows.sort()
marks = [i for i, r in enumerate(rows) if "refresh" in r[3]]
stems = [i for i, r in enumerate(rows) if r[3].startswith("stem_fwd")]
if len(stems) < 2:
    sys.exit("need >= 2 steps in the timeline")
# a step starts with the first kernel after the previous step's last kernel: cut at the largest gap before each stem forward
def step_start(i_stem):
    j = i_stem
    while j > 0 and rows[j][0] - (rows[j - 1][0] + rows[j - 1][1]) < 30.0 and i_stem - j < 80:
        j -= 1
    return j
s0, s1 = step_start(stems[-2]), step_start(stems[-1])
step = rows[s0:s1]
t0 = step[0][0]
t1 = max(r[0] + r[1] for r in step)
print(f"# kernels in the step: {len(step)}   span {1e-3 * (t1 - t0):.3f} ms   sum of durations {1e-3 * sum(r[1] for r in step):.3f} ms   streams {sorted(set(r[2] for r in step))}")
ev = []
for i, (ts, du, st, nm) in enumerate(step):
    ev.append((ts, 1, i))
    ev.append((ts + du, 0, i))
ev.sort()
active = set()
exposed = collections.defaultdict(float)
alone = collections.defaultdict(float)
idle = 0.0
prev = t0
for t, kind, i in ev:
    dt = t - prev
    if dt > 0:
        if not active:
            idle += dt
        else:
            for k in active:
                exposed[k] += dt / len(active)
            if len(active) == 1:
                alone[next(iter(active))] += dt
    prev = t
    if kind:
        active.add(i)
    else:
        active.discard(i)
if "--list" in sys.argv:
    for i, (ts, du, st, nm) in enumerate(step):
        print(f"{ts - t0:9.1f} +{du:7.1f} us  s{st:>3s}  exposed {exposed[i]:7.1f}  {nm[:60]}")
    sys.exit(0)
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0])
for i, (ts, du, st, nm) in enumerate(step):
    a = agg[nm]
    a[0] += exposed[i]; a[1] += du; a[2] += alone[i]; a[3] += 1
print(f"idle (no kernel running): {idle:.1f} us")
print("  exposed    total    alone   n  kernel")
for nm, (ex, du, al, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{ex:9.1f} {du:8.1f} {al:8.1f} {n:3d}  {nm[:80]}")

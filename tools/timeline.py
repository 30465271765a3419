"""Summarise an RTL_TRACE_FILE timeline (one JSON line per POA launch group): GPU busy fraction, concurrency, phases."""
import json
import sys

ev = [json.loads(l) for l in open(sys.argv[1])]
if len(sys.argv) > 2:  # keep chains from this epoch on (2 chains per correct_reads call and unit)
    ev = [e for e in ev if e["epoch"] > int(sys.argv[2])]
t0 = min(e["stage0"] for e in ev)
t1 = max(e["fold"] for e in ev)
print("groups %d  span %.1f ms" % (len(ev), t1 - t0))
# union of kernel intervals and average concurrency
pts = []
for e in ev:
    pts.append((e["k0"], 1))
    pts.append((e["k1"], -1))
pts.sort()
busy = 0.0
conc_area = 0.0
cur = 0
last = pts[0][0]
for t, d in pts:
    if cur > 0:
        busy += t - last
        conc_area += cur * (t - last)
    cur += d
    last = t
print("GPU busy (union of kernel intervals) %.1f ms = %.1f%% of span; mean concurrency while busy %.2f" % (
    busy, 100 * busy / (t1 - t0), conc_area / max(busy, 1e-9)))
for k, a, b in (("stage+submit", "stage0", "stage1"), ("queue (submit -> kernel start)", "stage1", "k0"),
                ("kernels", "k0", "k1"), ("kernel end -> host sync return", "k1", "sync"), ("fold", "sync", "fold")):
    v = [e[b] - e[a] for e in ev]
    print("%-34s mean %8.3f ms  max %8.3f  sum/units %9.1f" % (k, sum(v) / len(v), max(v), sum(v) / (1 + max(e["unit"] for e in ev))))

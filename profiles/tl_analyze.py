"""Occupancy of a launch timeline (profiles/c4_timeline.py output): how much of the wall time has only region-growing chains
running, how much has 1 / 2 / 3+ other kernels in flight, and each kernel's summed duration."""
import sys
rows = []
for ln in open(sys.argv[1]):
    if ln.startswith("#") or not ln.strip(): continue
    a = ln.split()
    rows.append((float(a[0]), float(a[1]), int(a[2]), " ".join(a[3:])))
ev = []
for t0, t1, c, n in rows:
    chain = n.startswith("k_lsd_grow_warp") or n.startswith("k_lsd_grow_cta")
    ev.append((t0, 1, chain)); ev.append((t1, -1, chain))
ev.sort()
tot = {}
nch = nbw = 0
last = ev[0][0]
for t, d, chain in ev:
    key = ("chains only" if nch and not nbw else "idle" if not nbw else "%d other kernel(s)%s" % (min(nbw, 3), "+" if nbw >= 3 else ""))
    tot[key] = tot.get(key, 0.0) + (t - last)
    last = t
    if chain: nch += d
    else: nbw += d
span = ev[-1][0] - ev[0][0]
print("span %.1f ms" % span)
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]): print("  %-22s %8.1f ms  %5.1f %%" % (k, v, 100 * v / span))
per = {}
for t0, t1, c, n in rows: per[n] = per.get(n, 0.0) + t1 - t0
print("summed durations:")
for k, v in sorted(per.items(), key=lambda kv: -kv[1])[:24]: print("  %-22s %8.1f ms" % (k, v))

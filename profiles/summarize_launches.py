"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel share table (markdown)."""
import csv, sys, collections

src, out, cmd = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
rows = []
with open(src) as f:
    lines = [l for l in f if not l.startswith("==")]
r = list(csv.reader(lines))
h = r[0]
ik, im, iv = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
iu = h.index("Metric Unit")
tot = collections.OrderedDict()
cnt = collections.Counter()
for row in r[1:]:
    if len(row) <= iv or row[im] != "gpu__time_duration.sum":
        continue
    v = float(row[iv].replace(",", ""))
    u = row[iu]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
    name = row[ik].split("(")[0][:60]
    tot[name] = tot.get(name, 0.0) + ms
    cnt[name] += 1
total = sum(tot.values())
with open(out, "w") as f:
    f.write("# ncu launch list summary\n\nCommand: `%s`\n" % cmd)
    f.write("(per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes; raw list: `%s`)\n\n" % src.split("/")[-1])
    f.write("| kernel | launches | total ms | share |\n|---|---|---|---|\n")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        f.write("| `%s` | %d | %.3f | %.1f%% |\n" % (k, cnt[k], v, 100 * v / total))
print("kernels", len(tot), "total ms", total)

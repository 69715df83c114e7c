"""Per-kernel warp-stall breakdown from an `ncu --page source --csv` export (SASS view): share of the samples by stall
reason and the hottest instructions.  Usage: summarize_source.py source.csv out.md "<command>" """
import csv, sys, collections

src, out = sys.argv[1], sys.argv[2]
cmd = sys.argv[3] if len(sys.argv) > 3 else ""
rows = list(csv.reader(open(src)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1].split("(")[0], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]):
        cur["rows"].append(r)
seen = set()
with open(out, "w") as f:
    f.write("# Warp-stall breakdown per kernel (ncu source page, round 1 final kernels)\n\nCommand: `%s`\n\n" % cmd)
    for b in blocks:
        if b["name"] in seen or not b["rows"]:
            continue
        seen.add(b["name"])
        h = b["hdr"]
        stall_cols = [i for i, k in enumerate(h) if k.startswith("stall_") and "Not Issued" not in k]
        isamp = h.index("# Samples")
        iexe = h.index("Instructions Executed")
        tot = collections.Counter()
        for r in b["rows"]:
            for i in stall_cols:
                try:
                    tot[h[i]] += float(r[i] or 0)
                except ValueError:
                    pass
        total = sum(tot.values()) or 1.0
        ninst = sum(float(r[iexe] or 0) for r in b["rows"])
        f.write("## `%s`\n\nwarp instructions executed: %.3g; samples: %d\n\n" % (b["name"], ninst, int(total)))
        f.write("stalls: " + ", ".join("%s %.0f%%" % (k.replace("stall_", ""), 100 * v / total) for k, v in tot.most_common(6)) + "\n\n")
        f.write("| samples | share | SASS |\n|---|---|---|\n")
        for r in sorted(b["rows"], key=lambda r: -float(r[isamp] or 0))[:8]:
            f.write("| %s | %.1f%% | `%s` |\n" % (r[isamp], 100 * float(r[isamp] or 0) / total, r[1].strip()[:90]))
        f.write("\n")
print(len(seen), "kernels")

"""Aggregate an `ncu --page raw --csv` export (one row per launch) into a per-kernel table: launches, time, DRAM bytes
(read + write), DRAM / SM throughput, achieved occupancy.  Usage: extract_ncu_all.py raw.csv out.json out.md "<command>" """
import csv, json, sys, collections

src, outj, outm = sys.argv[1], sys.argv[2], sys.argv[3]
cmd = sys.argv[4] if len(sys.argv) > 4 else ""
r = list(csv.reader(l for l in open(src) if not l.startswith("==")))
h, units = r[0], r[1]
col = {k: i for i, k in enumerate(h)}


def val(row, k):
    if k not in col or row[col[k]] in ("", "n/a"):
        return None
    x = float(row[col[k]].replace(",", ""))
    u = units[col[k]].lower()
    if u in ("tbyte/s",): x *= 1e12
    elif u in ("gbyte/s",): x *= 1e9
    elif u in ("mbyte/s",): x *= 1e6
    elif u in ("kbyte/s",): x *= 1e3
    elif u in ("gbyte", "gb"): x *= 1e9
    elif u in ("mbyte", "mb"): x *= 1e6
    elif u in ("kbyte", "kb"): x *= 1e3
    elif u in ("usecond", "us"): x *= 1e-3      # -> ms
    elif u in ("nsecond", "ns"): x *= 1e-6
    elif u in ("second", "s"): x *= 1e3
    return x


agg = collections.OrderedDict()
for row in r[2:]:
    name = row[col["Kernel Name"]].split("(")[0][:48]
    a = agg.setdefault(name, {"launches": 0, "ms": 0.0, "dram_bytes": 0.0, "dram_pct_t": 0.0, "sm_pct_t": 0.0, "occ_t": 0.0, "issue_t": 0.0,
                              "regs": None, "block": None})
    ms = val(row, "gpu__time_duration.sum") or 0.0
    a["launches"] += 1
    a["ms"] += ms
    if "dram__bytes_read.sum" in col:
        a["dram_bytes"] += (val(row, "dram__bytes_read.sum") or 0.0) + (val(row, "dram__bytes_write.sum") or 0.0)
    else:       # section captures hold the rate only: bytes = rate x duration
        a["dram_bytes"] += (val(row, "dram__bytes.sum.per_second") or 0.0) * ms * 1e-3
    a["dram_pct_t"] += ms * (val(row, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed") or 0.0)
    a["sm_pct_t"] += ms * (val(row, "sm__throughput.avg.pct_of_peak_sustained_elapsed") or 0.0)
    a["occ_t"] += ms * (val(row, "sm__warps_active.avg.pct_of_peak_sustained_active") or 0.0)
    a["issue_t"] += ms * (val(row, "smsp__issue_active.avg.pct") or val(row, "smsp__issue_active.avg.pct_of_peak_sustained_active") or 0.0)
    a["regs"] = val(row, "launch__registers_per_thread")
    a["block"] = val(row, "launch__block_size")
    a["l2hit_t"] = a.get("l2hit_t", 0.0) + ms * (val(row, "lts__t_sector_hit_rate.pct") or 0.0)
out = []
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
    t = a["ms"] or 1e-9
    out.append({"kernel": k, "launches": a["launches"], "ms": round(a["ms"], 4), "dram_bytes": a["dram_bytes"],
                "dram_gbs": round(a["dram_bytes"] / t / 1e6, 1), "dram_pct": round(a["dram_pct_t"] / t, 1), "sm_pct": round(a["sm_pct_t"] / t, 1),
                "warps_active_pct": round(a["occ_t"] / t, 1), "issue_active_pct": round(a["issue_t"] / t, 1), "l2_hit_pct": round(a["l2hit_t"] / t, 1), "regs": a["regs"], "block": a["block"]})
json.dump({"command": cmd, "kernels": out}, open(outj, "w"), indent=1)
with open(outm, "w") as f:
    f.write("# ncu per-kernel counters (round 1, final kernels)\n\nCommand: `%s`\n\n" % cmd)
    f.write("Times are cold-cache and serialised (first call of a fresh process): compare bytes and percentages, not ms.\n\n")
    f.write("| kernel | launches | ms | DRAM MB | DRAM GB/s | DRAM % | SM % | warps active % | issue active % | L2 hit % | regs | block |\n|---|---|---|---|---|---|---|---|---|---|---|---|\n")
    for o in out:
        f.write("| `%s` | %d | %.3f | %.1f | %.0f | %.1f | %.1f | %.1f | %.1f | %.1f | %s | %s |\n" % (
            o["kernel"], o["launches"], o["ms"], o["dram_bytes"] / 1e6, o["dram_gbs"], o["dram_pct"], o["sm_pct"], o["warps_active_pct"],
            o["issue_active_pct"], o["l2_hit_pct"], int(o["regs"]) if o["regs"] else "", int(o["block"]) if o["block"] else ""))
print(len(out), "kernels")

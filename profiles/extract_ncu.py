"""Extract the headline counters of one kernel from an ncu --set full report (csv of --page raw) into JSON."""
import csv, json, sys
src, kernel, out = sys.argv[1], sys.argv[2], sys.argv[3]
r = list(csv.reader(open(src)))
h, units = r[0], r[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
for v in r[2:]:
    if kernel in v[h.index("Kernel Name")]:
        d = {k: v[h.index(k)] for k in keys if k in h}
        d["_units"] = {k: units[h.index(k)] for k in keys if k in h}
        def mb(k):
            x = float(d[k]); u = d["_units"][k].lower()
            return x * (1e9 if u.startswith("g") else 1e6 if u.startswith("m") else 1e3 if u.startswith("k") else 1)
        d["dram_bytes_per_launch"] = mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum")
        d["note"] = sys.argv[4] if len(sys.argv) > 4 else ""
        json.dump(d, open(out, "w"), indent=1)
        print(json.dumps(d)[:400])
        break

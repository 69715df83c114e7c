"""Per-source-line sample shares of one kernel from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`:
python profiles/source_lines.py src.csv KERNEL [N]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
kern = sys.argv[2]; N = int(sys.argv[3]) if len(sys.argv) > 3 else 40
def num(x):
    try: return int(float(x))
    except Exception: return 0
segs = []; cur = None; i = 0
while i < len(rows):
    r = rows[i]
    if r and r[0] == "File Path":
        cur = {"file": r[1], "func": rows[i + 1][1].split("(")[0], "hdr": rows[i + 2], "rows": []}
        segs.append(cur); i += 3; continue
    if cur is not None and len(r) == len(cur["hdr"]): cur["rows"].append(r)
    i += 1
tot = sum(num(r[s["hdr"].index("# Samples")]) for s in segs if s["func"] == kern for r in s["rows"])
for s in segs:
    if s["func"] != kern: continue
    h = s["hdr"]; isamp = h.index("# Samples"); iexe = h.index("Instructions Executed")
    per = collections.Counter(); src = {}; exe = collections.Counter()
    for r in s["rows"]:
        ln = r[0]; per[ln] += num(r[isamp]); exe[ln] += num(r[iexe]); src[ln] = r[1]
    print("== %s in %s: %d of %d samples" % (kern, s["file"].split("/")[-1], sum(per.values()), tot))
    for ln, c in per.most_common(N):
        if c * 200 < tot: break
        print("%5s %6.2f%% exe %11d  %s" % (ln, 100 * c / tot, exe[ln], src[ln][:120]))

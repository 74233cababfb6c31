"""Summarise the per-instruction stall sampling of one kernel from `ncu --page source --csv`.
    ncu -i rep.ncu-rep --page source --csv --kernel-name regex:NAME --launch-count 1 > src.csv
    python scripts/ncu_hot.py src.csv [top]"""
import csv, sys
allrows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
block = int(sys.argv[3]) if len(sys.argv) > 3 else 0
starts = [i for i, r in enumerate(allrows) if r and r[0] == "Kernel Name"] + [len(allrows)]
rows = allrows[starts[block]:starts[block + 1]]
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[col["# Samples"]]) for r in data)
print("kernel:", rows[0][1][:80], " samples", tot, " instrs", len(data))
agg = {s: sum(int(r[col[s]]) for r in data) for s in stalls}
print("stall totals:", {k[6:]: "%.1f%%" % (100.0 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.01 * tot})
order = sorted(range(len(data)), key=lambda i: -int(data[i][col["# Samples"]]))[:top]
for i in sorted(order):
    r = data[i]
    n = int(r[col["# Samples"]])
    why = sorted(((int(r[col[s]]), s[6:]) for s in stalls), reverse=True)[:2]
    print("%5d %5.1f%%  %-58s %s" % (i, 100.0 * n / max(tot, 1), r[col["Source"]].strip()[:58], " ".join("%s=%d" % (b, a) for a, b in why if a)))

"""Summarise the LAST step of an `ncu --metrics gpu__time_duration.sum --csv` launch list (ids after the last
weight_norm_fwd launch... the last `per_step` launches).   python scripts/launch_list.py launches.csv per_step"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
per = int(sys.argv[2])
rows = rows[-per:]
# columns: ID, Process ID, Process Name, Host Name, Kernel Name, Context, Stream, Block Size, Grid Size, Device, CC, Section, Metric Name, Unit, Value
agg = collections.OrderedDict()
tot = 0.0
for r in rows:
    name = r[4].split("(")[0].replace("void ", "")
    name = name.split("<")[0]
    v = float(r[-1].replace(",", ""))
    unit = r[-2]
    us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += us; tot += us
print("# last %d launches, %.1f us of GPU time (cold-cache, serialised under ncu: compare SHARES)" % (len(rows), tot))
print("kernel,launches,total_us,share")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('"%s",%d,%.1f,%.4f' % (k, n, us, us / tot))

"""Turn the raw outputs of scripts/gpu_round.sh (gpurun_out/) into the tracked summaries under profiles/.
    python scripts/make_profiles.py r01_v5"""
import collections, csv, os, re, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01_v5"

raw = os.path.join(G, "step_raw.csv")
with open(raw, "w") as f:
    subprocess.run(["ncu", "-i", os.path.join(G, "step_full.ncu-rep"), "--page", "raw", "--csv"], stdout=f, stderr=subprocess.DEVNULL)
rows = list(csv.reader(open(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
keys = [k for k in ["gpu__time_duration.sum", "launch__grid_size", "dram__bytes_read.sum", "dram__bytes_write.sum",
                    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
                    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
                    "smsp__issue_active.avg.pct_of_peak_sustained_active",
                    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
                    "smsp__inst_executed.sum"] if k in col]
out = ["# ncu --set full --clock-control none --import-source on -k regex:sdf_query|sdf_render|sdf_bwd|head_fwd|head_bwd|wgrad --launch-skip 45 --launch-count 15",
       "#   python scripts/profile_step.py 1024 4      (the 4th training step: 1024 rays x 98 samples, beta 0.1 -> sampler k = 2), one B200",
       "# per-launch values, cold L2 under replay",
       "# units: " + ", ".join("%s [%s]" % (k, units[col[k]]) for k in keys),
       "kernel," + ",".join(keys)]
for d in data:
    out.append(d[col["Kernel Name"]].split("(")[0].replace("void ", "") + "," + ",".join(d[col[k]] for k in keys))
    if "wgrad" in d[col["Kernel Name"]]:
        print("wgrad DRAM bytes:", float(d[col["dram__bytes_read.sum"]]) + float(d[col["dram__bytes_write.sum"]]), "GB")
open(os.path.join(P, tag + "_ncu_full_summary.csv"), "w").write("\n".join(out) + "\n")

lines = [l for l in open(os.path.join(G, "launches.csv")) if not l.startswith("==")]
r = list(csv.DictReader(lines))
names = [x["Kernel Name"] for x in r]
vals = [float(x["Metric Value"].replace(",", "")) for x in r]
idx = [i for i, n in enumerate(names) if "weight_norm_fwd" in n]
s, e = idx[-2], idx[-1]
agg = collections.OrderedDict()


def short(n):
    n = re.sub(r"\bvoid |\bat::native::|\bat::|<unnamed>::|\(anonymous namespace\)::", "", n).replace("native::", "")
    return n.split("(")[0][:100]


for n, v in zip(names[s:e], vals[s:e]):
    a = agg.setdefault(short(n), [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v for _, v in agg.values())
with open(os.path.join(P, tag + "_launches_summary.csv"), "w") as f:
    f.write("# ncu launch list of ONE train step (3rd of 4; 1024 rays x 98 samples, beta=0.1 -> k=2), B200\n")
    f.write("# cmd: ncu --metrics gpu__time_duration.sum --clock-control none --csv python scripts/profile_step.py 1024 4\n")
    f.write("# %d launches, %.1f us of GPU time (cold-cache, serialised: compare SHARES).  neat:: kernels = hand-written sm_100a.\n" % (e - s, tot / 1e3))
    f.write("kernel,launches,total_us,share\n")
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write('"%s",%d,%.1f,%.4f\n' % (k, c, v / 1e3, v / tot))
neat = sum(v for k, (c, v) in agg.items() if "neat::" in k or "pack_" in k)
print("launches", e - s, "gpu us", round(tot / 1e3, 1), "neat share", round(neat / tot, 4))
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:8]:
    print("   %-40s %2d %8.1f us %5.1f%%" % (k[:40], c, v / 1e3, 100 * v / tot))

for k in ("sdf_query", "head_fwd", "sdf_render", "sdf_bwd", "wgrad", "head_bwd"):
    src = os.path.join(G, k + "_src.csv")
    with open(src, "w") as f:
        subprocess.run(["ncu", "-i", os.path.join(G, "step_full.ncu-rep"), "--page", "source", "--csv", "--kernel-name", "regex:" + k],
                       stdout=f, stderr=subprocess.DEVNULL)
    with open(os.path.join(P, "%s_%s_stalls.txt" % (tag, k)), "w") as f:
        subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_hot.py"), src, "10"], stdout=f)
for a, b in (("bench_default.json", "bench_default_1024"), ("bench_1024_20.json", "bench_1024_20steps"), ("bench_8192.json", "bench_8192"),
             ("bench_beta001.json", "bench_beta0.01"), ("bench_reference.json", "bench_reference_arm"),
             ("bench_eval_65536.json", "bench_eval_65536")):
    if os.path.exists(os.path.join(G, a)):
        shutil.copy(os.path.join(G, a), os.path.join(P, "%s_%s.json" % (tag, b)))

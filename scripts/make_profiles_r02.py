"""Turn the raw outputs of scripts/gpu_round2_evidence.sh (gpurun_out/) into the tracked summaries under profiles/.
    python scripts/make_profiles_r02.py [tag]"""
import collections, csv, json, os, re, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
PEAK_HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def short(n):
    n = re.sub(r"\bvoid |\bat::native::|\bat::|<unnamed>::|\(anonymous namespace\)::", "", n).replace("native::", "")
    return n.split("(")[0].split("<")[0][:100]


# ---- 1. ncu --set full of the six tile-MLP kernels
rep = os.path.join(G, "step_full.ncu-rep")
if os.path.exists(rep):
    raw = os.path.join(G, "step_raw.csv")
    with open(raw, "w") as f:
        subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=f, stderr=subprocess.DEVNULL)
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    keys = [k for k in ["gpu__time_duration.sum", "launch__grid_size", "dram__bytes_read.sum", "dram__bytes_write.sum",
                        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
                        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
                        "smsp__issue_active.avg.pct_of_peak_sustained_active",
                        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
                        "smsp__inst_executed.sum", "launch__registers_per_thread"] if k in col]
    out = ["# ncu --set full --clock-control none --import-source on -k regex:sdf_query|sdf_render|sdf_bwd|head_fwd|head_bwd|wgrad --launch-skip 30 --launch-count 15",
           "#   python scripts/profile_step.py 1024 4   (plugin path, eager launches; 1024 rays x 98 samples, beta 0.1 -> sampler k = 2), one B200",
           "# per-launch values, cold L2 under replay",
           "# units: " + ", ".join("%s [%s]" % (k, units[col[k]]) for k in keys),
           "kernel," + ",".join(keys)]
    dram = {}
    for d in data:
        name = short(d[col["Kernel Name"]]).replace("neat::", "")
        out.append(name + "," + ",".join(d[col[k]] for k in keys))
        b = float(d[col["dram__bytes_read.sum"]].replace(",", "")) + float(d[col["dram__bytes_write.sum"]].replace(",", ""))
        u = units[col["dram__bytes_read.sum"]].lower()
        b *= {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)
        grid = int(d[col["launch__grid_size"]].replace(",", ""))
        key = name.replace("_kernel", "")
        if grid >= 100:       # the full-size launch of the step (not the eikonal / surface-point ones)
            dram[key] = max(dram.get(key, 0.0), b)
    open(os.path.join(P, tag + "_ncu_full_summary.csv"), "w").write("\n".join(out) + "\n")
    dram["source"] = "profiles/%s_ncu_full_summary.csv (dram__bytes_read.sum + dram__bytes_write.sum of the largest launch)" % tag
    json.dump(dram, open(os.path.join(P, "ncu_dram_bytes_1024.json"), "w"), indent=1)
    print("DRAM bytes per launch:", {k: round(v / 1e9, 3) for k, v in dram.items() if k != "source"})
    for k in ("sdf_query", "head_fwd", "sdf_render", "sdf_bwd", "wgrad", "head_bwd"):
        src = os.path.join(G, k + "_src.csv")
        with open(src, "w") as f:
            subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + k], stdout=f, stderr=subprocess.DEVNULL)
        with open(os.path.join(P, "%s_%s_stalls.txt" % (tag, k)), "w") as f:
            subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_hot.py"), src, "10"], stdout=f)

# ---- 2. launch list of one step
lp = os.path.join(G, "launches.csv")
if os.path.exists(lp):
    lines = [l for l in open(lp) if not l.startswith("==")]
    r = list(csv.DictReader(lines))
    names = [x["Kernel Name"] for x in r]
    vals = [float(x["Metric Value"].replace(",", "")) for x in r]
    idx = [i for i, n in enumerate(names) if "weight_norm_fwd" in n]
    s, e = idx[-2], idx[-1]
    agg = collections.OrderedDict()
    for n, v in zip(names[s:e], vals[s:e]):
        a = agg.setdefault(short(n), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v for _, v in agg.values())
    with open(os.path.join(P, tag + "_launches_summary.csv"), "w") as f:
        f.write("# ncu launch list of ONE train step (3rd of 4; plugin path; 1024 rays x 98 samples, beta=0.1 -> k=2), B200\n")
        f.write("# cmd: ncu --metrics gpu__time_duration.sum --clock-control none --csv python scripts/profile_step.py 1024 4\n")
        f.write("# %d launches, %.1f us of GPU time (cold-cache, serialised: compare SHARES).  neat:: kernels = hand-written sm_100a.\n" % (e - s, tot / 1e3))
        f.write("kernel,launches,total_us,share\n")
        for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('"%s",%d,%.1f,%.4f\n' % (k, c, v / 1e3, v / tot))
    neat = sum(v for k, (c, v) in agg.items() if "neat::" in k or "pack_" in k)
    lib = [k for k in agg if "cutlass" in k or "cublas" in k or "gemm" in k.lower() and "neat::" not in k]
    print("launches/step", e - s, "gpu us", round(tot / 1e3, 1), "neat share", round(neat / tot, 4), "library GEMM kernels:", lib)

# ---- 3. HBM metrics of the per-ray / dataset kernels
def hbm_table(path, out_name, title):
    if not os.path.exists(path):
        return
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for x in csv.DictReader(lines):
        k = short(x["Kernel Name"]).replace("neat::", "")
        v = float(x["Metric Value"].replace(",", ""))
        u = x["Metric Unit"].lower()
        m = x["Metric Name"]
        if m.startswith("dram__bytes"):
            v *= {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)
        if m == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(u, 1.0)   # -> us
        a = agg.setdefault(k, collections.defaultdict(list))
        a[m].append(v)
    with open(os.path.join(P, out_name), "w") as f:
        f.write("# %s\n" % title)
        f.write("# per launch (mean over the launches captured): time, DRAM bytes read + written, achieved DRAM GB/s, and ncu's own\n")
        f.write("# gpu__dram_throughput (%% of peak) / sm__throughput; measured copy bandwidth of this pool: %.0f GB/s.  Launches of a few\n" % PEAK_HBM)
        f.write("# microseconds on KB..MB of data are latency bound: their GB/s is what it is, the time column is what the step pays.\n")
        f.write("kernel,launches,grid,mean_us,dram_read_MB,dram_write_MB,achieved_GBps,frac_of_measured_copy_bw,ncu_dram_pct,ncu_sm_pct\n")
        for k, a in agg.items():
            n = len(a["gpu__time_duration.sum"])
            us = sum(a["gpu__time_duration.sum"]) / n
            rd, wr = sum(a["dram__bytes_read.sum"]) / n, sum(a["dram__bytes_write.sum"]) / n
            gbs = (rd + wr) / (us * 1e-6) / 1e9 if us > 0 else 0.0
            grid = max(a.get("launch__grid_size", [0]))
            f.write("%s,%d,%d,%.2f,%.3f,%.3f,%.1f,%.4f,%.1f,%.1f\n" % (k, n, grid, us, rd / 1e6, wr / 1e6, gbs, gbs / PEAK_HBM,
                    sum(a["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]) / n,
                    sum(a["sm__throughput.avg.pct_of_peak_sustained_elapsed"]) / n))
    print("wrote", out_name, len(agg), "kernels")


hbm_table(os.path.join(G, "perray_metrics.csv"), tag + "_perray_kernels_hbm.csv",
          "per-ray / loss / optimizer kernels of the training step: ncu --metrics gpu__time_duration.sum,dram__bytes_*.sum,... python scripts/profile_step.py 1024 3 0.01 (beta 0.01: 5 sampler iterations)")
hbm_table(os.path.join(G, "aux_metrics.csv"), tag + "_dataset_kernels_hbm.csv",
          "dataset-side and finalisation kernels at DTU image size (1200 x 1600, 300 lines; 65536 predictions): ncu ... python scripts/profile_aux.py")
for t in ("racecheck", "synccheck", "memcheck"):
    src = os.path.join(G, "sanitizer_%s.log" % t)
    if os.path.exists(src):
        shutil.copy(src, os.path.join(P, "%s_sanitizer_%s.log" % (tag, t)))

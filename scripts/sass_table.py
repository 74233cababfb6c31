"""Per-kernel SASS evidence table (no GPU needed): cuobjdump -sass of the built library, instruction counts that prove
the Blackwell-native paths (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA bulk copies ->
UBLKCP / UTMALDG, tcgen05.commit -> UTCBAR, mbarrier -> SYNCS, vector reductions -> RED, setmaxnreg -> USETMAXREG).
    python scripts/sass_table.py [tag]      -> profiles/<tag>_sass_table.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
so = os.path.join(ROOT, "neat_b200", "libneat_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
MN = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UBLKPF", "UTMALDG", "SYNCS", "HMMA", "MUFU.EX2", "MUFU.LG2", "RED.E", "REDG",
      "LDG.E", "STG.E", "LDS", "STS", "ELECT", "USETMAXREG"]
cnt = collections.OrderedDict()
cur = None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", name).replace("void ", "")
        cnt[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        cnt[cur]["_total"] += 1
        for k in MN:
            if op.startswith(k):
                cnt[cur][k] += 1
rows = ["# cuobjdump -sass neat_b200/libneat_b200.so (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a): instruction counts per kernel",
        "# UTCHMMA = tcgen05.mma kind::f16, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk (1-D TMA), SYNCS = mbarrier",
        "kernel,total," + ",".join(MN)]
for k, c in cnt.items():
    rows.append(k + "," + str(c["_total"]) + "," + ",".join(str(c[m]) for m in MN))
tot = collections.Counter()
for c in cnt.values():
    tot.update(c)
rows.append("ALL," + str(tot["_total"]) + "," + ",".join(str(tot[m]) for m in MN))
out = os.path.join(ROOT, "profiles", tag + "_sass_table.txt")
open(out, "w").write("\n".join(rows) + "\n")
print(out)
for r in rows:
    if any(x in r for x in ("sdf_", "head_", "wgrad", "ALL", "kernel,")):
        print(r)

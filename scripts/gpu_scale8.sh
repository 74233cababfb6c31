#!/bin/bash
# the driver's scaling command at N = 8 (and N = 4), full line with extras
mkdir -p gpurun_out
for n in 8 4; do
  T0=$(date +%s)
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  echo "N=$n rc=$? secs=$(( $(date +%s) - T0 ))"
done
python - <<'PY'
import json
for n in (8, 4):
    f = "gpurun_out/scale_n%d.json" % n
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(n, "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "dp_check", d.get("dp_check"))
        for r in d.get("per_rank") or []:
            print("   ", r)
        for c in d.get("configs", []):
            print("   cfg", c.get("name", "")[:60], round(c.get("value", 0)), c.get("failed"))
    except Exception as e:
        print(n, "failed", e)
        print(open(f.replace(".json", ".err")).read()[-2000:])
PY

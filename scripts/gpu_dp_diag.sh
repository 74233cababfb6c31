#!/bin/bash
# N=1 and N=2 on the SAME box, headline config only, with the per-rank graph A / graph B timings
mkdir -p gpurun_out
F="--steps 40 --warmup 5 --no-cpu-baseline --no-eager-baseline --no-extra-configs"
python bench.py --gpus 1 $F > gpurun_out/dp_n1.json 2> gpurun_out/dp_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 $F > gpurun_out/dp_n2.json 2> gpurun_out/dp_n2.err
python - <<'PY'
import json
for f in ("gpurun_out/dp_n1.json", "gpurun_out/dp_n2.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "graph", d.get("graph_ms"), "k_mean", d.get("sampler_iters_k_mean"))
        for r in d.get("per_rank") or []:
            print("   ", r)
    except Exception as e:
        print(f, "failed", e)
        print(open(f.replace(".json", ".err")).read()[-1500:])
PY

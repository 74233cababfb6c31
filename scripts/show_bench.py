import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e); continue
    print(f, round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "k", d.get("sampler_iters_k"), "e2e", round(d["e2e"]["value"]))
    if "kernel_ms_per_step" in d: print("   ", d["kernel_ms_per_step"])

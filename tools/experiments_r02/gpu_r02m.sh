#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02m}
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo rc=$?
timeout 600 python bench.py --config 2 --steps 10 --warmup 3 > gpurun_out/${T}_bench_c2.json 2>> gpurun_out/${T}_bench.err; echo rc=$?
timeout 900 python bench.py --config 3 --steps 5 --warmup 3 > gpurun_out/${T}_bench_c3.json 2>> gpurun_out/${T}_bench.err; echo rc=$?
python - <<PY
import json
for f in ("", "_c2", "_c3"):
    try:
        d = json.load(open("gpurun_out/${T}_bench%s.json" % f))
        print(f or "c1", "value %.0f ms/step %.3f e2e %.0f lat %.2f frac %.4f hrnet_ms %.3f crops/step %d | %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["p50_batch_latency_ms"], d["roofline"]["frac"], d["roofline"]["hrnet_ms"], d["crops_per_step_per_gpu"], d["parity_check"]))
        if d.get("stage_rooflines"): print({k: (round(v["frac"], 3), round(v["ms"], 3)) for k, v in d["stage_rooflines"].items()})
        print(d["stages_ms"])
        print(d["cpu_baseline"])
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/${T}_bench.err

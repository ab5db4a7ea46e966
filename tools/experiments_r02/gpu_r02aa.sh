#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02aa}
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -5 | tee gpurun_out/${T}_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02aa_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'hrnet',d['roofline']['hrnet_ms'])
print(d['stages_ms']); print(d['parity_check'])
print({k:round(v['frac'],3) for k,v in d['stage_rooflines'].items()})
PY

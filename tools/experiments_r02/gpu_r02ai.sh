#!/bin/bash
mkdir -p gpurun_out
T=r02ai
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -3 | tee gpurun_out/${T}_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['parity_check'][:12])
print(d['stages_ms'])
print({k:(round(v['frac'],3),round(v['ms'],4)) for k,v in d['stage_rooflines'].items()})" | tee gpurun_out/${T}_bench.log

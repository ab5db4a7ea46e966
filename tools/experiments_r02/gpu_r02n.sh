#!/bin/bash
mkdir -p gpurun_out
T=r02n
timeout 600 python tools/dual_ctx_probe.py 2>&1 | grep -v "^\[" | tee gpurun_out/${T}_dual_ctx.log
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -5 | tee gpurun_out/${T}_pytest.log
HBP_TIMELINE=1 timeout 300 python bench.py --steps 3 --warmup 3 > /dev/null 2> gpurun_out/${T}_timeline.log; grep -c "^\[tl\]" gpurun_out/${T}_timeline.log

#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02ae}
timeout 120 python tools/nms_prof.py 20 2>&1 | grep -v "^\[" | tee gpurun_out/${T}_nms.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_nms_launches.csv python tools/nms_prof.py 3 > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r02ae_nms_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size')
agg=collections.OrderedDict()
for r in rows[1:]:
    k=(r[ki][:60],r[gi]); agg.setdefault(k,[]).append(float(r[vi].replace(',','')))
for k,v in agg.items(): print(k, 'n=%d avg %.1f us'%(len(v), sum(v)/len(v)/1e3))
PY
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_det_pose.py tests/test_gpu_entrypoints.py -m gpu -x -q --timeout 600 2>&1 | tail -3

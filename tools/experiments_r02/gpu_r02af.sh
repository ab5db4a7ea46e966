#!/bin/bash
mkdir -p gpurun_out
T=r02af
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,smsp__inst_executed.sum --clock-control none --cache-control none --csv --log-file gpurun_out/${T}_nms_launches.csv python tools/nms_prof.py 3 > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r02af_nms_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size'); mi=hdr.index('Metric Name')
agg=collections.OrderedDict()
for r in rows[1:]:
    k=(r[ki][:40],r[gi],r[mi]); agg.setdefault(k,[]).append(float(r[vi].replace(',','')))
for k,v in agg.items():
    if 'conv' in k[0] or 'stem' in k[0] or 'upsample' in k[0] or 'head_k' in k[0]: continue
    print(k, 'n=%d avg %.1f'%(len(v), sum(v)/len(v)))
PY

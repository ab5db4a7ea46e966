#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -2
for i in 1 2 3 4 5 6 7 8 9 10 11 12 13 14; do
  timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/flaky_$i.json 2> gpurun_out/flaky_$i.err
  rc=$?
  echo "run $i rc=$rc bytes=$(wc -c < gpurun_out/flaky_$i.json)"
  if [ ! -s gpurun_out/flaky_$i.json ]; then tail -15 gpurun_out/flaky_$i.err; fi
done

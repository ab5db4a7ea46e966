for v in 0 1 2 3 8; do
  HBP_PG_DBG=$v HBP_TIMELINE=1 timeout 300 python bench.py --steps 3 --warmup 2 > /dev/null 2> gpurun_out/tl_dbg$v.log
  echo "dbg=$v"; awk '/----/{c++} c==2' gpurun_out/tl_dbg$v.log | grep -E "stage4.0.fuse_level|stage3.1.fuse_level|conv2 " | sed -E 's/kind=[0-9] //; s/\[tl\] +//' | awk '{printf "%-28s %8.1f\n",$2,$11-$9}'
done

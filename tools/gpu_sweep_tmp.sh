run() { env "$@" timeout 300 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$*','hrnet_ms',round(d['roofline']['hrnet_ms'],4))"; }
run A=1
run HBP_HALO_DBG=4
run HBP_HALO_DBG=4 HBP_PG_DBG=8
HBP_HALO_DBG=4 HBP_TIMELINE=1 timeout 300 python bench.py --steps 3 --warmup 2 > /dev/null 2> gpurun_out/tl_empty.log

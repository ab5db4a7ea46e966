run() { env "$@" timeout 300 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$*','hrnet_ms',round(d['roofline']['hrnet_ms'],4))"; }
run A=1
run HBP_PG_M2=1
run HBP_PG_M2=1 HBP_BRANCH_SHARE4=0.372,0.203,0.263,0.162
HBP_PG_M2=1 timeout 300 python -m pytest tests -m gpu -q -x -k "hrnet_w32_tcgen05" 2>&1 | tail -1

run() { env "$@" timeout 300 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$*','hrnet_ms',round(d['roofline']['hrnet_ms'],4))"; }
run HBP_HALO_1X1_NMAX=256
run HBP_HALO_1X1_NMAX=128
run HBP_HALO_1X1_NMAX=64

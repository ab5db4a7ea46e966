for b in 1 2 4 8; do
  HBP_UPADD_BPSM=$b timeout 300 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bpsm',$b,'hrnet_ms',d['roofline']['hrnet_ms'])"
done

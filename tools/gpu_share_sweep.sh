#!/bin/bash
# sweep of the per-stage branch SM shares (whole-network HRNet time)
run() { env "$@" python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.4f' % d['roofline']['hrnet_ms'])"; }
echo "default: $(run A=1)"
for s4 in "0.36,0.20,0.26,0.18" "0.32,0.20,0.26,0.22" "0.40,0.18,0.24,0.18" "0.36,0.16,0.30,0.18" "0.36,0.24,0.22,0.18" "0.30,0.18,0.24,0.28" "0.25,0.25,0.25,0.25" "0,0,0,0"; do echo "share4 $s4: $(run HBP_BRANCH_SHARE4=$s4)"; done
for s3 in "0.44,0.24,0.32" "0.50,0.22,0.28" "0.40,0.26,0.34" "0.44,0.20,0.36" "0.44,0.30,0.26" "0,0,0"; do echo "share3 $s3: $(run HBP_BRANCH_SHARE3=$s3)"; done
for s2 in "0.64,0.36" "0.70,0.30" "0.56,0.44" "0,0"; do echo "share2 $s2: $(run HBP_BRANCH_SHARE2=$s2)"; done

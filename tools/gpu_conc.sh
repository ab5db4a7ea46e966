#!/bin/bash
# which of the round-2 conv changes breaks two concurrent 64-crop forwards on one GPU?
try() { lbl=$1; shift
  out=$(env "$@" timeout 120 python tools/dual_ctx_probe.py 2 64 2>&1 | grep -v "^\[plan\|^\[pgroup" | tail -2 | tr '\n' ' ' | cut -c1-200)
  echo "$lbl: $out"
}
try default HBP_X=0
try teams2 HBP_HALO_TEAMS=2
try noreverse HBP_REVERSE=0
try nohints HBP_L2_HINTS=0
try smem200 HBP_HALO_SMEM_KB=200
try stores16 HBP_HALO_DBG=32
try teams2_smem200 HBP_HALO_TEAMS=2 HBP_HALO_SMEM_KB=200
try nopdl HBP_PDL=0
try default_again HBP_X=0

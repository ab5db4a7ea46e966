#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02w}
for sms in 53 30; do
  echo "== default plan, SMS=$sms" | tee -a gpurun_out/${T}_b2.log
  HBP_HALO_SMS=$sms HBP_MB_SHAPES=2 HBP_CONV_TRACE=1 timeout 120 python tools/conv_microbench.py 2>&1 | grep "eng=\|\[plan\]" | cut -c1-200 | tee -a gpurun_out/${T}_b2.log
  echo "== N=64 resident a_stages=1, SMS=$sms" | tee -a gpurun_out/${T}_b2.log
  HBP_HALO_SMS=$sms HBP_HALO_N=64 HBP_HALO_RES_MINA=1 HBP_MB_SHAPES=2 HBP_CONV_TRACE=1 timeout 120 python tools/conv_microbench.py 2>&1 | grep "eng=\|\[plan\]" | cut -c1-200 | tee -a gpurun_out/${T}_b2.log
  echo "== N=64 streaming, SMS=$sms" | tee -a gpurun_out/${T}_b2.log
  HBP_HALO_SMS=$sms HBP_HALO_N=64 HBP_MB_SHAPES=2 HBP_CONV_TRACE=1 timeout 120 python tools/conv_microbench.py 2>&1 | grep "eng=\|\[plan\]" | cut -c1-200 | tee -a gpurun_out/${T}_b2.log
  echo "== M=2, SMS=$sms" | tee -a gpurun_out/${T}_b2.log
  HBP_HALO_SMS=$sms HBP_HALO_M=2 HBP_MB_SHAPES=2 HBP_CONV_TRACE=1 timeout 120 python tools/conv_microbench.py 2>&1 | grep "eng=\|\[plan\]" | cut -c1-200 | tee -a gpurun_out/${T}_b2.log
done
for sms in 65 85 53 30; do
  echo "== b0/b1 SMS=$sms" | tee -a gpurun_out/${T}_b2.log
  HBP_HALO_SMS=$sms HBP_MB_SHAPES=0,1 timeout 120 python tools/conv_microbench.py 2>&1 | grep "eng=" | tee -a gpurun_out/${T}_b2.log
done

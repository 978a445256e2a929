#!/bin/bash
mkdir -p gpurun_out
python tools/gemm_timeline.py bf16 64 > gpurun_out/gemm_timeline_bf16.txt 2>&1
cat gpurun_out/gemm_timeline_bf16.txt | tail -40
CASES=sa1_fwd_l1,sa1_fwd_l2,sa1_dgrad_l2,sa1_dgrad_l1,sa1_wgrad_l2,sa1_wgrad_l1,sa2_fwd_l0,fwd_xf,sa2_fwd_l2,dgrad_epi2,sa2_dgrad_l0,wgrad_xf,sa2_wgrad_l0
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_tn_kernel|wgrad_kernel' -o gpurun_out/r02_gemm python tools/ncu_gemm.py $CASES > gpurun_out/ncu_gemm.log 2>&1
tail -20 gpurun_out/ncu_gemm.log
ls -la gpurun_out

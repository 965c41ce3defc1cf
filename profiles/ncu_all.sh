#!/bin/bash
# Round-2 ncu captures of the kernel variants that ship (run on the GPU box from the repo root: bash profiles/ncu_all.sh).
# Every target first runs without ncu (must exit 0), then one launch of it is captured with --set full.
set -u
mkdir -p gpurun_out
ncu --query-metrics 2>/dev/null | grep -i -E "tensor|pipe_tc|tmem|utc" > gpurun_out/r02_ncu_tensor_metrics.txt
declare -A KERN=( [pre128]=conv_gemm_halo [pre256]=conv_gemm_halo [res1x1]=conv_gemm_pair [c1x1]=conv_gemm_pair [shuffle]=conv_gemm_pair [init]=init_conv [final]=final_conv [gca_pool]=gca_pool_kernel [attn]=attn_mqa_tc [cublas]="gemm|cutlass|nvjet|sm100" )
for t in "$@"; do
  python profiles/ncu_targets.py $t > gpurun_out/r02_${t}_plain.log 2>&1 || { echo "$t: plain run failed"; tail -3 gpurun_out/r02_${t}_plain.log; continue; }
  cat gpurun_out/r02_${t}_plain.log
  ncu --set full --clock-control none --import-source on -k "regex:${KERN[$t]}" -s 3 -c 1 -f -o gpurun_out/r02_$t python profiles/ncu_targets.py $t > gpurun_out/r02_${t}_ncu.log 2>&1
  tail -2 gpurun_out/r02_${t}_ncu.log
done

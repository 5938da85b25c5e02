#!/bin/bash
# round-2 GPU pass B (1 GPU): ncu --set full of the stream-K layer-1 products
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for w in l1 w1; do
  python scripts/profile_gemm.py $w || exit 1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32_2cta -s 8 -c 1 -f -o gpurun_out/r2_ncu_$w python scripts/profile_gemm.py $w > gpurun_out/r2_ncu_$w.log 2>&1
  tail -2 gpurun_out/r2_ncu_$w.log
done

#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for st in 3 4 5; do echo "msub1 stages $st"; WGAN_B200_NO_PAIR_MSUB2=1 WGAN_B200_GEMM_STAGES=$st python scripts/profile_gemm.py l1; done
echo "msub2 stages 3"; WGAN_B200_GEMM_STAGES=3 python scripts/profile_gemm.py l1
for w in l1; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32_2cta -s 8 -c 1 -f -o gpurun_out/r2_ncu_${w}_msub2 python scripts/profile_gemm.py $w > gpurun_out/r2_ncu_${w}_msub2.log 2>&1
  tail -1 gpurun_out/r2_ncu_${w}_msub2.log
done

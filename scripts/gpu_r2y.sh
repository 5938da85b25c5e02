#!/bin/bash
# round-2 GPU pass Y (1 GPU): default bench line with the batch-32 extras, ncu refresh of the final layer-1 kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2y_bench.err
python scripts/bench_value.py gpurun_out/r2y_bench.json
for w in l1 w1; do
  python scripts/profile_gemm.py $w
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32_2cta -s 8 -c 1 -f -o gpurun_out/r2_ncu_$w python scripts/profile_gemm.py $w > gpurun_out/r2_ncu_$w.log 2>&1
done

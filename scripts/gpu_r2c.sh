#!/bin/bash
# round-2 GPU pass C (1 GPU): burst issue in the pair kernel's producer -- GEMM parity, isolated timings A/B, whole-iteration A/B
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -q -x > gpurun_out/r2c_gemm_tests.log 2>&1; echo "gemm tests rc=$?"; tail -3 gpurun_out/r2c_gemm_tests.log
WGAN_B200_PAIR_MSUB2=1 timeout 600 python -m pytest tests/test_gemm_gpu.py -q -x -k streamk > gpurun_out/r2c_gemm_tests_msub2.log 2>&1; echo "gemm tests (512-row tiles) rc=$?"; tail -3 gpurun_out/r2c_gemm_tests_msub2.log
for w in l1 w1; do
  for b in 1 2 3; do echo "burst $b"; WGAN_B200_GEMM_BURST=$b python scripts/profile_gemm.py $w; done
  echo "512-row tiles, burst 2"; WGAN_B200_PAIR_MSUB2=1 python scripts/profile_gemm.py $w
done
for b in 1 2 3; do
WGAN_B200_GEMM_BURST=$b timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2c_b1_burst$b.json 2> gpurun_out/r2c_b1_burst$b.err
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2c_b1.json 2> gpurun_out/r2c_b1.err
python scripts/bench_value.py gpurun_out/r2c_b1*.json

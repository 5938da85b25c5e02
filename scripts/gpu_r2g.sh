#!/bin/bash
# round-2 GPU pass G (8 GPUs): scaling of BASELINE config 2 with the engine's own exchange vs NCCL
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
run() { # name nproc env...
  name=$1; n=$2; shift 2
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline --no-gram --no-e2e > gpurun_out/r2g_$name.json 2> gpurun_out/r2g_$name.err
}
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gram --no-e2e > gpurun_out/r2g_b1.json 2> gpurun_out/r2g_b1.err
run b8_engine148 8 WGAN_B200_DP_BLOCKS=148
run b8_engine32 8 WGAN_B200_DP_BLOCKS=32
run b8_nccl 8 WGAN_B200_DP_BACKEND=nccl
run b4_engine32 4 WGAN_B200_DP_BLOCKS=32
run b2_engine32 2 WGAN_B200_DP_BLOCKS=32
python scripts/bench_value.py gpurun_out/r2g_*.json
timeout 300 python -m pytest tests/test_dp_gpu.py -q -x > gpurun_out/r2g_dp_tests_8gpu.log 2>&1; echo "dp tests (8 GPUs) rc=$?"; tail -3 gpurun_out/r2g_dp_tests_8gpu.log

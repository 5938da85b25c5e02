#!/bin/bash
# round-2 GPU pass F (2 GPUs): SM reserve x exchange block count
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gram --no-e2e > gpurun_out/r2f_b1.json 2> gpurun_out/r2f_b1.err
for cfg in "0 148" "8 148" "8 32" "16 32" "8 16" "0 32" "16 148"; do
  set -- $cfg
  WGAN_B200_PREPARE_SM_RESERVE=$1 WGAN_B200_DP_BLOCKS=$2 timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-gram --no-e2e > gpurun_out/r2f_b2_r$1_b$2.json 2> gpurun_out/r2f_b2_r$1_b$2.err
done
python scripts/bench_value.py gpurun_out/r2f_b*.json

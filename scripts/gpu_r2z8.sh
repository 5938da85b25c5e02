#!/bin/bash
# round-2 GPU pass Z8 (8 GPUs): exchange block shape at 8 GPUs
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
B="bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gram --no-e2e --no-extras"
timeout 200 python $B > gpurun_out/r2z8_n1.json 2>/dev/null
for cfg in "1024 8" "1024 16" "256 32"; do
  set -- $cfg
  WGAN_B200_DP_THREADS=$1 WGAN_B200_DP_BLOCKS=$2 timeout 300 $TR $B --gpus 8 > gpurun_out/r2z8_n8_t$1_b$2.json 2>/dev/null
done
python scripts/bench_value.py gpurun_out/r2z8_*.json

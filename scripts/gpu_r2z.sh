#!/bin/bash
# round-2 GPU pass Z (2 GPUs): exchange block shape (threads x blocks) next to the overlapped kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
B="bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gram --no-e2e --no-extras"
timeout 200 python $B > gpurun_out/r2z_n1.json 2>/dev/null
for cfg in "256 32" "512 16" "1024 8" "512 32" "256 16"; do
  set -- $cfg
  WGAN_B200_DP_THREADS=$1 WGAN_B200_DP_BLOCKS=$2 timeout 300 $TR $B --gpus 2 > gpurun_out/r2z_n2_t$1_b$2.json 2>/dev/null
done
python scripts/bench_value.py gpurun_out/r2z_*.json

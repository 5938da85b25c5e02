#!/bin/bash
# round-2 GPU pass X (8 GPUs): final scaling series of BASELINE config 2 on ONE box (same build for every line)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
run() { # name nproc env...
  name=$1; n=$2; shift 2
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 --no-gram --no-extras > gpurun_out/r2x_$name.json 2> gpurun_out/r2x_$name.err
  echo "$name rc=$?"
}
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gram --no-extras > gpurun_out/r2x_n1.json 2> gpurun_out/r2x_n1.err
run n8 8 A=1
run n4 4 A=1
run n2 2 A=1
run n8_nccl 8 WGAN_B200_DP_BACKEND=nccl
python scripts/bench_value.py gpurun_out/r2x_*.json

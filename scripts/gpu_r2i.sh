#!/bin/bash
# round-2 GPU pass I: BASELINE configs 3 / 4 / 5 (and 2 with e2e) on N GPUs; usage: gpu_r2i.sh "<N list>"
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
run() { # config nproc extra...
  c=$1; n=$2; shift 2
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --config $c --steps 10 --warmup 3 "$@" > gpurun_out/r2i_c${c}_n$n.json 2> gpurun_out/r2i_c${c}_n$n.err
  echo "config $c n $n rc=$?"
}
for n in $1; do
  run 3 $n
  run 5 $n
  if [ "$n" != "4" ]; then run 4 $n; fi
  if [ "$n" = "8" ]; then run 2 $n --no-gram --no-extras; fi
done
python scripts/bench_value.py gpurun_out/r2i_*.json

#!/bin/bash
# round-2 GPU pass H (1 GPU): bench.py --config 2/3/4/5
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r2h_c2.json 2> gpurun_out/r2h_c2.err; echo "c2 rc=$?"; tail -3 gpurun_out/r2h_c2.err
for c in 3 4 5; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/r2h_c$c.json 2> gpurun_out/r2h_c$c.err; echo "c$c rc=$?"; tail -3 gpurun_out/r2h_c$c.err
done
python scripts/bench_value.py gpurun_out/r2h_c*.json

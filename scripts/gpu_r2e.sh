#!/bin/bash
# round-2 GPU pass E (2 GPUs): overlapped generator exchange -- DP parity tests, 2-GPU bench, timeline
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dp_gpu.py -q -x > gpurun_out/r2e_dp_tests.log 2>&1; echo "dp tests rc=$?"; tail -3 gpurun_out/r2e_dp_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-gram > gpurun_out/r2e_b2.json 2> gpurun_out/r2e_b2.err
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gram > gpurun_out/r2e_b1.json 2> gpurun_out/r2e_b1.err
python scripts/bench_value.py gpurun_out/r2e_b*.json
timeout 200 $TR scripts/probes/dp_timeline.py > /dev/null 2>&1

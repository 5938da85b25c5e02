#!/bin/bash
# round-2 GPU pass A (2 GPUs): full GPU suite with log, 2-GPU benches of the engine all-reduce, probe
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_parity_evidence_gpu.py > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2a_tests.log
timeout 600 python -m pytest tests/test_parity_evidence_gpu.py -q > gpurun_out/r2a_evidence.log 2>&1; echo "evidence rc=$?"; tail -15 gpurun_out/r2a_evidence.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for b in 32 148; do
  WGAN_B200_DP_BLOCKS=$b timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_b2_push$b.json 2> gpurun_out/r2a_b2_push$b.err
done
WGAN_B200_DP_BACKEND=nccl timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_b2_nccl.json 2> gpurun_out/r2a_b2_nccl.err
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_b1.json 2> gpurun_out/r2a_b1.err
for b in 32 148; do echo "blocks $b"; WGAN_B200_DP_BLOCKS=$b timeout 120 $TR scripts/probes/allreduce_probe.py 2>/dev/null | grep -v "^\*\|OMP"; done
python scripts/bench_value.py gpurun_out/r2a_b*.json

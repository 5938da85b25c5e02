#!/bin/bash
# round-2 GPU pass L (1 GPU): column sums on the auxiliary stream -- step / DP parity tests, whole-iteration A/B
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_steps_gpu.py tests/test_dp_gpu.py tests/test_parity_evidence_gpu.py -q > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2l_tests.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-extras --no-gram"
timeout 300 $B > gpurun_out/r2l_aux.json 2> gpurun_out/r2l_aux.err
WGAN_B200_NO_AUX_OVERLAP=1 timeout 300 $B > gpurun_out/r2l_noaux.json 2> gpurun_out/r2l_noaux.err
timeout 300 $B > gpurun_out/r2l_aux_b.json 2> gpurun_out/r2l_aux_b.err
WGAN_B200_NO_AUX_OVERLAP=1 timeout 300 $B > gpurun_out/r2l_noaux_b.json 2> gpurun_out/r2l_noaux_b.err
python scripts/bench_value.py gpurun_out/r2l_*.json

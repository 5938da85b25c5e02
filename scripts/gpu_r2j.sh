#!/bin/bash
# round-2 GPU pass J (1 GPU): full GPU suite, launch list of the iteration, ncu of the generator-output products, load-pattern ubench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2j_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2j_tests.log
./scripts/ubench/l1_stream > gpurun_out/r2j_l1_stream.txt 2>&1; tail -3 gpurun_out/r2j_l1_stream.txt
BENCH="python bench.py --steps 2 --warmup 1 --no-e2e --no-gram --no-extras --no-cpu-baseline"
$BENCH > gpurun_out/r2j_b.json 2> gpurun_out/r2j_b.err && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2j_launches.csv $BENCH > gpurun_out/r2j_ncu.log 2>&1
echo "launch list rc=$?"
python scripts/profile_gemm.py genout && timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32_2cta -s 8 -c 1 -f -o gpurun_out/r2_ncu_genout python scripts/profile_gemm.py genout > gpurun_out/r2_ncu_genout.log 2>&1
python scripts/profile_prepare.py && timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32_2cta -s 17 -c 1 -f -o gpurun_out/r2_ncu_genout_blend python scripts/profile_prepare.py > gpurun_out/r2_ncu_genout_blend.log 2>&1
tail -1 gpurun_out/r2_ncu_genout_blend.log

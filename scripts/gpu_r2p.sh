#!/bin/bash
# round-2 GPU pass P (1 GPU): final verification -- full suite, smoke, default bench line, reference arm
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2p_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2p_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2p_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2p_smoke.log
timeout 900 python bench.py > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2p_bench_reference.json 2> gpurun_out/r2p_bench_reference.err; echo "reference rc=$?"
python scripts/bench_value.py gpurun_out/r2p_bench.json gpurun_out/r2p_bench_reference.json

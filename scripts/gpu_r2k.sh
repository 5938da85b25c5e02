#!/bin/bash
# round-2 GPU pass K (1 GPU): full GPU suite (no -x) + A/B of the stream-K policy on the whole iteration
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2k_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2k_tests.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-extras"
timeout 300 $B > gpurun_out/r2k_sk6.json 2> gpurun_out/r2k_sk6.err
WGAN_B200_SK_MAX_CONTRIB=4 timeout 300 $B > gpurun_out/r2k_sk4.json 2> gpurun_out/r2k_sk4.err
timeout 300 $B > gpurun_out/r2k_sk6b.json 2> gpurun_out/r2k_sk6b.err
python scripts/bench_value.py gpurun_out/r2k_sk*.json
python - <<'PY'
import json
for f in ("gpurun_out/r2k_sk6.json","gpurun_out/r2k_sk4.json"):
    d=json.loads([l for l in open(f) if l.startswith("{")][-1])
    print(f, "gram", round(d["gram_form"]["value"],1), "launches", d["gpu_launches"])
PY

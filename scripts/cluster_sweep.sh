#!/bin/bash
# GEMM correctness + timing for each cluster size (WGAN_B200_GEMM_CLUSTER = 1, 2, 4)
for c in 1 2 4; do
  echo "== cluster $c"
  WGAN_B200_GEMM_CLUSTER=$c timeout 600 python -m pytest tests/test_gemm_gpu.py -q -m gpu -x 2>&1 | tail -2
  for w in genout gx l1 dgrad2; do WGAN_B200_GEMM_CLUSTER=$c timeout 120 python scripts/profile_gemm.py $w; done
  WGAN_B200_GEMM_CLUSTER=$c timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench it/s', d['value'], 'ms', d['ms_per_step'], 'sampler', d['sampler']['value']); [print(o['kernel'][:40], o['ms']) for o in d['roofline']['others']]"
done

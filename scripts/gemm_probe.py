"""Diagnostic: run the GEMM test hook on a few shapes per operand layout and print the error
(used when bringing up the tcgen05 descriptors on a real B200)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
from test_gemm_gpu import _trainer, _run

for mode in (1, 0):
    tr = _trainer(mode)
    for layout in ("dgrad", "fwd", "wgrad"):
        for (M, N, K) in [(128, 32, 32), (128, 64, 64), (128, 224, 96), (100, 200, 133), (300, 600, 100), (661, 200, 512)]:
            rng = np.random.default_rng(0)
            try:
                r = _run(tr, layout, M, N, K, rng)
                print(f"mode={mode} {layout:6s} M={M} N={N} K={K} err={r['err']:.3e}", flush=True)
            except Exception as e:  # noqa
                print(f"mode={mode} {layout:6s} M={M} N={N} K={K} EXC {e}", flush=True)
                sys.exit(1)
print("probe done")

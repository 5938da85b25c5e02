"""Back-to-back timing of the small (launch-latency-bound) GEMMs of a critic update:
python scripts/small_gemm_probe.py   (WGAN_B200_GEMM_2CTA=0/1 selects the kernel)."""
import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from arshamg_scrnaseq_wgan_b200 import WGANConfig, WGANTrainer

tr = WGANTrainer(WGANConfig(), max_batch=256, device=0)
dev = "cuda:0"
p = lambda t: C.c_void_p(t.data_ptr())
null = C.c_void_p()
tiles = C.c_int(0)
r4 = lambda v: (v + 3) // 4 * 4
SHAPES = [("fwd", 12288, 200, 200), ("dgrad", 12288, 200, 200), ("fwd", 4096, 200, 200), ("wgrad", 200, 200, 12288),
          ("wgrad", 200, 200, 6605), ("wgrad", 200, 200, 4096), ("fwd", 6605, 200, 200), ("fwd", 4096, 600, 100),
          ("fwd", 4096, 600, 600), ("wgrad", 600, 600, 4096), ("wgrad", 100, 600, 4096)]
REPS = 200
for layout, M, N, K in SHAPES:
    a_mn, b_mn = {"fwd": (0, 1), "dgrad": (0, 0), "wgrad": (1, 1)}[layout]
    A = torch.rand((K if a_mn else M, r4(M if a_mn else K) + 32), device=dev) - 0.5
    B = torch.rand((K if b_mn else N, r4(N if b_mn else K) + 32), device=dev) - 0.5
    D = torch.empty((M, r4(N)), device=dev)
    def fn():
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        rc = tr.lib.wgan_gemm(tr._h, a_mn, b_mn, p(A), A.stride(0), p(B), B.stride(0), p(D), D.stride(0), M, N, K,
                              null, 0, null, 0, null, null, C.byref(tiles), 0, st)
        assert rc == 0, tr.lib.wgan_last_error(tr._h)
    for _ in range(5):
        fn()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(REPS):
            fn()
    g.replay(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); g.replay(); e.record(); torch.cuda.synchronize()
    us = s.elapsed_time(e) * 1e3 / REPS
    print(f"{layout:6s} {M:6d} {N:4d} {K:6d}  2cta={os.environ.get('WGAN_B200_GEMM_2CTA','1')}  {us:7.2f} us/call (incl. finish)  "
          f"{2.0*M*N*K/us/1e6:7.1f} TFLOP/s", flush=True)

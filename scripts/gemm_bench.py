"""GEMM sweep: python scripts/gemm_bench.py "layout,M,N,K,splits[,bias]" ...  (cluster via env)."""
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
r32 = lambda v: (v + 31) // 32 * 32
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)

for spec in sys.argv[1:]:
    f = spec.split(",")
    layout, M, N, K, splits = f[0], int(f[1]), int(f[2]), int(f[3]), int(f[4])
    a_mn, b_mn = {"fwd": (0, 1), "dgrad": (0, 0), "wgrad": (1, 1)}[layout]
    lda = int(f[6]) if len(f) > 6 and f[6] else r32(M if a_mn else K)
    ldb = int(f[7]) if len(f) > 7 and f[7] else r32(N if b_mn else K)
    ldd = int(f[8]) if len(f) > 8 and f[8] else r32(N)
    A = torch.rand((K if a_mn else M, lda), device=dev) - 0.5
    B = torch.rand((K if b_mn else N, ldb), device=dev) - 0.5
    D = torch.empty((M, ldd), device=dev)
    bias = torch.zeros((N,), device=dev) if len(f) > 5 and f[5] else None

    def fn():
        rc = tr.lib.wgan_gemm(tr._h, a_mn, b_mn, p(A), A.stride(0), p(B), B.stride(0), p(D), D.stride(0), M, N, K,
                              p(bias) if bias is not None else null, 1 if bias is not None else 0, null, 0, null, null,
                              C.byref(tiles), splits, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0, tr.lib.wgan_last_error(tr._h)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(8):
        flush.zero_()                      # evict L2 between timed launches
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ms = sorted(ts)[len(ts) // 2]
    flops = 2.0 * M * N * K
    byts = 4.0 * (M * K + N * K + M * N)
    print(f"{spec:36s} cluster={os.environ.get('WGAN_B200_GEMM_CLUSTER','2')} {ms*1e3:8.1f} us  {flops/ms/1e9:7.1f} TFLOP/s  {byts/ms/1e6:7.0f} GB/s(alg)", flush=True)

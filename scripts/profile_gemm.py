"""Runs selected GEMM shapes of the WGAN-GP iteration a few times (target of `ncu --set full`)."""
import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from arshamg_scrnaseq_wgan_b200 import WGANConfig, WGANTrainer

which = sys.argv[1] if len(sys.argv) > 1 else "genout"
B = 4096
tr = WGANTrainer(WGANConfig(), max_batch=B, device=0)
G, Gl, Hd, Hg = 6605, 6608, 200, 600
dev = "cuda:0"
p = lambda t: C.c_void_p(t.data_ptr())
null = C.c_void_p()
tiles = C.c_int(0)

def gemm(a_mn, b_mn, A, Bm, D, M, N, K, bias=None, act=0, aux=None, sumsq=None):
    rc = tr.lib.wgan_gemm(tr._h, a_mn, b_mn, p(A), A.stride(0), p(Bm), Bm.stride(0), p(D), D.stride(0), M, N, K,
                          p(bias) if bias is not None else null, act, p(aux) if aux is not None else null,
                          aux.stride(0) if aux is not None else 0, null, p(sumsq) if sumsq is not None else null,
                          C.byref(tiles), 0, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, tr.lib.wgan_last_error(tr._h)

if which == "genout":
    A = torch.rand((B, Hg), device=dev); W = torch.rand((Hg, Gl), device=dev) - 0.5
    bias = torch.zeros((G,), device=dev); D = torch.empty((B, Gl), device=dev)
    fn = lambda: gemm(0, 1, A, W, D, B, G, Hg, bias, 1)
elif which == "gx":
    A = torch.rand((B, Hd), device=dev); W = torch.rand((G, Hd), device=dev) - 0.5
    D = torch.empty((B, Gl), device=dev); sq = torch.empty((128 * B,), device=dev)
    fn = lambda: gemm(0, 0, A, W, D, B, G, Hd, sumsq=sq)
elif which == "l1":
    A = torch.rand((3 * B, Gl), device=dev); W = torch.rand((G, Hd), device=dev) - 0.5
    bias = torch.zeros((Hd,), device=dev); D = torch.empty((3 * B, Hd), device=dev)
    fn = lambda: gemm(0, 1, A, W, D, 3 * B, Hd, G, bias, 1)
elif which == "w1":
    A = torch.rand((3 * B, Gl), device=dev); Dl = torch.rand((3 * B, Hd), device=dev) - 0.5
    D = torch.empty((G, Hd), device=dev)
    fn = lambda: gemm(1, 1, A, Dl, D, G, Hd, 3 * B)
elif which == "dgrad2":
    A = torch.rand((3 * B, Hd), device=dev); W = torch.rand((Hd, Hd), device=dev) - 0.5
    aux = torch.rand((3 * B, Hd), device=dev) - 0.5; D = torch.empty((3 * B, Hd), device=dev)
    fn = lambda: gemm(0, 0, A, W, D, 3 * B, Hd, Hd, aux=aux)
for _ in range(6):
    fn()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    fn()
e.record(); torch.cuda.synchronize()
print(which, "ms per launch", s.elapsed_time(e) / 10)

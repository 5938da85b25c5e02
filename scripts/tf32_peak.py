"""cuBLAS TF32 dense peak, measured the way MEASURED_PEAKS.json measures bf16: torch.matmul 8192^3,
best of 10 (burst) and back to back for 4 s (sustained).  Also our own GEMM on the same problem."""
import sys, os, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
torch.backends.cuda.matmul.allow_tf32 = True
N = 8192
a = torch.randn(N, N, device="cuda"); b = torch.randn(N, N, device="cuda"); c = torch.empty(N, N, device="cuda")
fl = 2.0 * N ** 3
for _ in range(3): torch.matmul(a, b, out=c)
best = 1e9
for _ in range(10):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); torch.matmul(a, b, out=c); e.record(); torch.cuda.synchronize()
    best = min(best, s.elapsed_time(e))
print(f"cuBLAS TF32 8192^3 burst: {fl/best/1e9:.1f} TFLOP/s")
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.time(); n = 0
s.record()
while time.time() - t0 < 4.0:
    for _ in range(10): torch.matmul(a, b, out=c)
    n += 10
    torch.cuda.synchronize()
e.record(); torch.cuda.synchronize()
print(f"cuBLAS TF32 8192^3 sustained (4 s): {fl*n/s.elapsed_time(e)/1e9:.1f} TFLOP/s")
from arshamg_scrnaseq_wgan_b200 import WGANConfig, WGANTrainer
tr = WGANTrainer(WGANConfig(), max_batch=256, device=0)
p = lambda t: C.c_void_p(t.data_ptr()); null = C.c_void_p(); tiles = C.c_int(0)
for name, (amn, bmn) in (("K,K", (0, 0)), ("K,MN", (0, 1)), ("MN,MN", (1, 1))):
    def fn():
        rc = tr.lib.wgan_gemm(tr._h, amn, bmn, p(a), N, p(b), N, p(c), N, N, N, N, null, 0, null, 0, null, null,
                              C.byref(tiles), 1, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0
    for _ in range(3): fn()
    best = 1e9
    for _ in range(10):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    print(f"gemm_tf32_2cta_kernel ({name}) 8192^3 burst: {fl/best/1e9:.1f} TFLOP/s")

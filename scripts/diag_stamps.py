"""Diagnostic: per-CTA phase stamps of the stream-K layer-1 product (needs libwgan_diag_stamps.so, scripts/diag_build.sh).
WGAN_B200_LIB=libwgan_diag_stamps.so python scripts/diag_stamps.py [l1|w1]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from arshamg_scrnaseq_wgan_b200 import WGANConfig, WGANTrainer
which = sys.argv[1] if len(sys.argv) > 1 else "l1"
B, G, Gl, Hd = 4096, 6605, 6608, 200
tr = WGANTrainer(WGANConfig(), max_batch=B, device=0)
p = lambda t: C.c_void_p(t.data_ptr())
null, tiles = C.c_void_p(), C.c_int(0)
X3 = torch.rand((3 * B + 1, Gl), device="cuda")[:3 * B]
if which == "l1":
    W = torch.rand((G + 1, Hd), device="cuda")[:G] - 0.5; bias = torch.zeros((Hd,), device="cuda"); D = torch.empty((3 * B, Hd), device="cuda")
    args = (0, 1, p(X3), Gl, p(W), Hd, p(D), Hd, 3 * B, Hd, G, p(bias), 1)
else:
    Dl = torch.rand((3 * B + 1, Hd), device="cuda")[:3 * B] - 0.5; D = torch.empty((G, Hd), device="cuda")
    args = (1, 1, p(X3), Gl, p(Dl), Hd, p(D), Hd, G, Hd, 3 * B, null, 0)
for _ in range(6):
    rc = tr.lib.wgan_gemm(tr._h, *args, null, 0, null, null, C.byref(tiles), 0, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
torch.cuda.synchronize()
st = np.zeros((160, 8), np.uint64)
fn = tr.lib.wgan_diag_read_stamps
fn.restype = C.c_int; fn.argtypes = [C.c_void_p]
assert fn(st.ctypes.data_as(C.c_void_p)) == 0
st = st[:148].astype(np.int64)
t0 = st[:, 0].min()
modes = st[:, 7]
names = ["start", "wait acc (seg 0)", "acc ready seg 0", "epilogue done seg 0", "acc ready last seg", "epilogue done last seg", "stores drained"]
print(f"{which}: times in us relative to the first CTA's start (148 CTAs)")
for i, n in enumerate(names):
    col = st[:, i]; ok = col > 0
    v = (col[ok] - t0) / 1e3
    print(f"  {n:26s} n={ok.sum():3d} min {v.min():7.2f} avg {v.mean():7.2f} max {v.max():7.2f}")
# per role: last segment mode (1 = partial, 2 = owner, 0 = complete tile); single-segment CTAs have no 'last seg' stamps
last_mode = np.where(st[:, 4] > 0, (modes >> 8) & 0xFF, modes & 0xFF) - 1
end = np.where(st[:, 5] > 0, st[:, 5], st[:, 3])
ready = np.where(st[:, 4] > 0, st[:, 4], st[:, 2])
for m, nm in ((1, "last segment = partial"), (2, "last segment = owner"), (0, "complete tile")):
    sel = last_mode == m
    if sel.any():
        print(f"  {nm:24s} n={sel.sum():3d}  acc ready avg {(ready[sel] - t0).mean() / 1e3:7.2f} max {(ready[sel] - t0).max() / 1e3:7.2f}   epilogue takes avg {(end[sel] - ready[sel]).mean() / 1e3:6.2f} max {(end[sel] - ready[sel]).max() / 1e3:6.2f}   done max {(end[sel] - t0).max() / 1e3:7.2f}")

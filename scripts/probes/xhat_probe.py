"""Probe: is the interpolate x^ written by the generator-output GEMM (second output) ever wrong?"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from test_steps_gpu import _setup, REF, SMALL

B = int(os.environ.get("PB", 130))
tr, cfg, ocfg, P, (x, z, eps), (xd, zd, ed) = _setup(REF, B, 10.0, 0)
slot = tr.real_slot(B)
slot.copy_(xd)
xcat = tr.buffer("xcat")
z3 = torch.full_like(zd, 3.0)
out = torch.empty((B, tr.Gl), device="cuda")[:, :cfg.gex_size]
bad = 0
ref_hat = None
for it in range(int(os.environ.get("PN", 300))):
    mode = it % 3
    xcat[B:2 * B].fill_(-7.0)                      # poison the x^ rows
    if mode == 0:
        tr.critic_prepare(slot, zd, ed)
    elif mode == 1:
        tr.critic_prepare(slot, zd, ed); tr.generator(z3, out=out); tr.critic_prepare(slot, zd, ed)
    else:
        tr.generator(z3, out=out); tr.critic_prepare(slot, zd, ed)
    torch.cuda.synchronize()
    xt = xcat[:B].clone(); xh = xcat[B:2 * B].clone()
    want = ed[:, None] * slot + (1 - ed[:, None]) * xt
    if ref_hat is None:
        ref_hat, ref_t = xh.clone(), xt.clone()
        print("first: max |xh - torch blend|", (xh - want).abs().max().item())
    d = (xh - ref_hat).abs()
    if d.max().item() != 0 or not torch.equal(xt, ref_t):
        bad += 1
        rows = (d.max(dim=1).values > 0).nonzero().flatten().tolist()
        cols = (d.max(dim=0).values > 0).nonzero().flatten().tolist()
        print(f"it {it} mode {mode}: xt equal {torch.equal(xt, ref_t)}; xh maxdiff {d.max().item():.4g} rows {rows[:10]} n={len(rows)} "
              f"cols {cols[:6]}..{cols[-3:]} n={len(cols)}; poison left {(xh == -7.0).sum().item()}")
print("bad iterations:", bad)

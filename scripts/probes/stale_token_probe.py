"""Probe: exact sequence of test_stacked_real_slot_path_matches_oracle[REF-130-0]; on a mismatch, say which buffers differ."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from test_steps_gpu import _setup, REF, SMALL
from oracle import wgan_oracle as O

def bufs(tr):
    torch.cuda.synchronize()
    return {n: tr.buffer(n).clone() for n in ("norms", "h1", "h2", "d", "gx", "xcat")}

for dims, B, mode in [(SMALL, 300, 1), (REF, 130, 1), (SMALL, 300, 0), (REF, 130, 0)]:
    tr, cfg, ocfg, P, (x, z, eps), (xd, zd, ed) = _setup(dims, B, 10.0, mode)
    slot = tr.real_slot(B)
    slot.copy_(xd)
    s = tr.critic_step(slot, zd, ed, apply=False)
    g1 = {k: tr.get_tensor(k, "grad").copy() for k in O.CRITIC_NAMES}
    slot.copy_(xd)
    tok = tr.critic_prepare(slot, zd, ed)
    s2 = tr.critic_step(slot, zd, ed, apply=False, token=tok)
    for k in O.CRITIC_NAMES:
        assert np.array_equal(g1[k], tr.get_tensor(k, "grad")), k
    tok = tr.critic_prepare(slot, zd, ed)
    tr.generator(torch.full_like(zd, 3.0))
    s3 = tr.critic_step(slot, zd, ed, apply=False, token=tok)
    b3 = bufs(tr)
    g3 = {k: tr.get_tensor(k, "grad").copy() for k in O.CRITIC_NAMES}
    s4 = tr.critic_step(slot, zd, ed, apply=False)
    b4 = bufs(tr)
    print(dims["gex_size"], B, mode, "s==s2", s == s2, "s==s3", s == s3, "s==s4", s == s4)
    if s != s3:
        print(" s ", s); print(" s3", s3)
        for n in b3:
            d = (b3[n] - b4[n]).abs()
            rows = (d.reshape(d.shape[0], -1).max(dim=1).values > 0).nonzero().flatten().tolist() if d.dim() == 2 else []
            print("  ", n, "maxdiff", d.max().item(), "rows differing", len(rows), rows[:12], rows[-3:])
        for k in O.CRITIC_NAMES:
            print("  grad", k, np.abs(g3[k] - g1[k]).max())

"""Diagnostic: per-tensor gradient error of the device step vs the fp64 oracle, both GEMM modes."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from test_steps_gpu import _setup, _relerr, SMALL, REF
from oracle import wgan_oracle as O

def _fro(got, ref):
    ref = np.asarray(ref, np.float64).reshape(got.shape)
    return float(np.linalg.norm(got - ref) / (np.linalg.norm(ref) + 1e-30))

for dims, B in ((SMALL, 48), (REF, 64), (REF, 512), (REF, 4096)):
    for lam in (0.0, 10.0):
        for mode in (1, 0):
            tr, cfg, ocfg, P, (x, z, eps), (xd, zd, ed) = _setup(dims, B, lam, mode)
            s = tr.critic_step(xd, zd, ed, apply=False)
            so, go = O.critic_step_grads(P, x, z, eps, ocfg)
            g = tr.get_grads(O.CRITIC_NAMES)
            errs = {k.split('/')[-2][-6:] + '/' + k[-1]: (_relerr(g[k], go[k]), _fro(g[k], go[k])) for k in O.CRITIC_NAMES if np.abs(go[k]).max() > 0}
            msg = " ".join(f"{k}={v[0]:.1e}|{v[1]:.1e}" for k, v in errs.items())
            extra = ""
            if lam:
                _, _, n = O.gradient_penalty(P, x, O.generator_forward(P, z)[2], eps, lam)
                gn = tr.buffer("norms").cpu().numpy()[0]
                extra = f" norms[min {n.min():.3f} max {n.max():.3f}] normerr={_relerr(gn, n):.1e}"
            print(f"G={dims['gex_size']} B={B} lam={lam} mode={mode}: W={s['wasserstein']:.6f}/{so['wasserstein']:.6f} gp={s['gp']:.5f}/{so['gp']:.5f} {msg}{extra}", flush=True)
            s2 = tr.generator_step(zd, apply=False)
            so2, go2 = O.generator_step_grads(P, z, ocfg)
            g2 = tr.get_grads(O.GEN_NAMES)
            errs = {k.split('/')[-2][-6:] + '/' + k[-1]: (_relerr(g2[k], go2[k]), _fro(g2[k], go2[k])) for k in O.GEN_NAMES}
            print(f"   gen: loss={s2['gen_loss']:.6f}/{so2['gen_loss']:.6f} " + " ".join(f"{k}={v[0]:.1e}|{v[1]:.1e}" for k, v in errs.items()), flush=True)
            tr.close()

"""FP32 multi-threaded CPU restatement of the reference graph for TIMING (TEST INFRASTRUCTURE ONLY).

"CPU restatement of the reference graph (TensorFlow not installable)": the same hand-derived
math as oracle/wgan_oracle.py, evaluated with torch CPU float32 ops (MKL/oneDNN SGEMM, the same
kernel class TensorFlow's CPU MatMul uses) so that it can use all host cores.  Only bench.py's
``cpu_baseline`` leg and ``--impl reference`` arm run it.  It is validated against the numpy
oracle in tests/test_oracle.py.  PARITY UNPINNED BY THE REFERENCE (see wgan_oracle.py header).

Follows: generator [W:61-83], critic [W:103-124], losses [W:137-138], var lists [W:140-142],
update order [W:196,199], TF-Adam [A:140-151,214-225]; gradient penalty = SURVEY A.4.
"""
from __future__ import annotations

import torch

from .wgan_oracle import GEN_NAMES, CRITIC_NAMES, OracleConfig


def _lrelu(a, alpha):
    return torch.maximum(a, alpha * a)


def _slope(h, alpha):
    return torch.where(h > 0, torch.ones((), dtype=h.dtype), torch.full((), alpha, dtype=h.dtype))


class TorchCpuTrainer:
    def __init__(self, cfg: OracleConfig, params, dtype=torch.float32):
        self.cfg = cfg
        self.P = {k: torch.tensor(v, dtype=dtype) for k, v in params.items()}
        self.m = {k: torch.zeros_like(v) for k, v in self.P.items()}
        self.v = {k: torch.zeros_like(v) for k, v in self.P.items()}
        self.pw = {"g": [cfg.beta1, cfg.beta2], "c": [cfg.beta1, cfg.beta2]}   # [A:123-128]

    # ---- forward -----------------------------------------------------------------------
    def gen_fwd(self, z):
        P, a = self.P, self.cfg.alpha
        h1 = _lrelu(z @ P[GEN_NAMES[0]] + P[GEN_NAMES[1]], a)
        h2 = _lrelu(h1 @ P[GEN_NAMES[2]] + P[GEN_NAMES[3]], a)
        return h1, h2, _lrelu(h2 @ P[GEN_NAMES[4]] + P[GEN_NAMES[5]], a)

    def critic_fwd(self, x):
        P, a = self.P, self.cfg.alpha
        h1 = _lrelu(x @ P[CRITIC_NAMES[0]] + P[CRITIC_NAMES[1]], a)
        h2 = _lrelu(h1 @ P[CRITIC_NAMES[2]] + P[CRITIC_NAMES[3]], a)
        return h1, h2, h2 @ P[CRITIC_NAMES[4]] + P[CRITIC_NAMES[5]]

    def _critic_bwd(self, x, h1, h2, dd, need_dx=False):
        P, a = self.P, self.cfg.alpha
        g = {CRITIC_NAMES[4]: h2.T @ dd, CRITIC_NAMES[5]: dd.sum(0)}
        d2 = (dd @ P[CRITIC_NAMES[4]].T) * _slope(h2, a)
        g[CRITIC_NAMES[2]] = h1.T @ d2
        g[CRITIC_NAMES[3]] = d2.sum(0)
        d1 = (d2 @ P[CRITIC_NAMES[2]].T) * _slope(h1, a)
        g[CRITIC_NAMES[0]] = x.T @ d1
        g[CRITIC_NAMES[1]] = d1.sum(0)
        return g, (d1 @ P[CRITIC_NAMES[0]].T if need_dx else None)

    # ---- Adam --------------------------------------------------------------------------
    def _adam(self, names, grads, key):
        c = self.cfg
        b1p, b2p = self.pw[key]
        lr_t = c.learn_rate * (1 - b2p) ** 0.5 / (1 - b1p)
        for k in names:
            g = grads[k].reshape(self.P[k].shape)
            self.m[k] += (g - self.m[k]) * (1 - c.beta1)
            self.v[k] += (g * g - self.v[k]) * (1 - c.beta2)
            self.P[k] -= (self.m[k] * lr_t) / (self.v[k].sqrt() + c.epsilon)
        self.pw[key] = [b1p * c.beta1, b2p * c.beta2]

    # ---- steps -------------------------------------------------------------------------
    def critic_step(self, x, z, eps=None, apply=True):
        P, a, lam = self.P, self.cfg.alpha, self.cfg.lambda_gp
        B = x.shape[0]
        xt = self.gen_fwd(z)[2]
        h1r, h2r, dr = self.critic_fwd(x)
        h1f, h2f, df = self.critic_fwd(xt)
        wass = df.mean() - dr.mean()
        gr, _ = self._critic_bwd(x, h1r, h2r, torch.full_like(dr, -1.0 / B))
        gf, _ = self._critic_bwd(xt, h1f, h2f, torch.full_like(df, 1.0 / B))
        grads = {k: gr[k] + gf[k] for k in CRITIC_NAMES}
        gp = torch.zeros(())
        if lam != 0.0:
            W1, W2, w3 = P[CRITIC_NAMES[0]], P[CRITIC_NAMES[2]], P[CRITIC_NAMES[4]]
            e = eps.reshape(-1, 1)
            xh = e * x + (1 - e) * xt
            h1, h2, _ = self.critic_fwd(xh)
            s1, s2 = _slope(h1, a), _slope(h2, a)
            g2 = w3.T * s2
            g1 = (g2 @ W2.T) * s1
            gx = g1 @ W1.T
            n = gx.square().sum(1).sqrt()
            gp = lam / B * (n - 1).square().sum()
            c = (lam * 2.0 / B) * (n - 1) / n
            u = c[:, None] * gx
            grads[CRITIC_NAMES[0]] = grads[CRITIC_NAMES[0]] + u.T @ g1
            v1 = (u @ W1) * s1
            grads[CRITIC_NAMES[2]] = grads[CRITIC_NAMES[2]] + v1.T @ g2
            v2 = (v1 @ W2) * s2
            grads[CRITIC_NAMES[4]] = grads[CRITIC_NAMES[4]] + v2.sum(0).reshape(-1, 1)
        if apply:
            self._adam(CRITIC_NAMES, grads, "c")
        return {"wasserstein": float(wass), "gp": float(gp), "critic_loss": float(wass + gp)}, grads

    def generator_step(self, z, apply=True):
        P, a = self.P, self.cfg.alpha
        B = z.shape[0]
        gh1, gh2, xt = self.gen_fwd(z)
        h1, h2, d = self.critic_fwd(xt)
        _, dxt = self._critic_bwd(xt, h1, h2, torch.full_like(d, -1.0 / B), need_dx=True)
        d3 = dxt * _slope(xt, a)
        g = {GEN_NAMES[4]: gh2.T @ d3, GEN_NAMES[5]: d3.sum(0)}
        d2 = (d3 @ P[GEN_NAMES[4]].T) * _slope(gh2, a)
        g[GEN_NAMES[2]] = gh1.T @ d2
        g[GEN_NAMES[3]] = d2.sum(0)
        d1 = (d2 @ P[GEN_NAMES[2]].T) * _slope(gh1, a)
        g[GEN_NAMES[0]] = z.T @ d1
        g[GEN_NAMES[1]] = d1.sum(0)
        if apply:
            self._adam(GEN_NAMES, g, "g")
        return {"gen_loss": float(-d.mean())}, g

    def reference_iteration(self, x, z, eps=None):
        """[W:196] generator update, then [W:199] critic update, same batch and noise."""
        sg, _ = self.generator_step(z)
        sc, _ = self.critic_step(x, z, eps)
        return {**sg, **sc}

    def wgan_gp_iteration(self, xs, zs, epss, zg):
        """n_critic critic updates (fresh batch each) then one generator update."""
        out = {}
        for x, z, e in zip(xs, zs, epss):
            out.update(self.critic_step(x, z, e)[0])
        out.update(self.generator_step(zg)[0])
        return out

    def sample(self, z):
        return self.gen_fwd(z)[2]

"""Regenerates tests/golden/wgan_small.npz from the numpy oracle (fp64, fixed seeds).

The reference ships no golden vectors and cannot run here (TensorFlow 1.x, Python 3.12, no
network), so these vectors pin the ORACLE ITSELF against regressions; the oracle in turn is
cross-checked against an independent torch.autograd implementation (tests/test_oracle.py).
Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import wgan_oracle as O  # noqa: E402

DIMS = dict(noise_input_size=6, inflate_to_size=10, gex_size=13, disc_internal_size=7)


def main():
    cfg = O.OracleConfig(lambda_gp=10.0, learn_rate=1e-3, **DIMS)
    P = O.random_params(cfg, seed=42)
    rng = np.random.default_rng(7)
    B = 5
    x = rng.random((B, cfg.gex_size))
    z = np.abs(rng.normal(0, 0.1, (B, cfg.noise_input_size)) + rng.poisson(1.0, (B, cfg.noise_input_size)))
    eps = rng.random(B)
    out = {"x": x, "z": z, "eps": eps}
    for k, v in P.items():
        out["P/" + k] = v
    sc, gc = O.critic_step_grads(P, x, z, eps, cfg)
    sg, gg = O.generator_step_grads(P, z, cfg)
    for k, v in {**gc, **gg}.items():
        out["G/" + k] = v
    out["scalars"] = np.array([sc["wasserstein"], sc["gp"], sc["critic_loss"], sg["gen_loss"]])
    tr = O.OracleTrainer(cfg, P)
    for _ in range(3):
        tr.critic_step(x, z, eps)
        tr.generator_step(z)
    for k, v in tr.P.items():
        out["P3/" + k] = v
    out["sample"] = tr.sample(z)
    out["noise"] = O.rng_noise_prior(123, 4, 10, 6, 5, 0.1)
    out["uniform"] = O.rng_uniform(123, 4, 10, 6)
    out["indices"] = O.rng_indices(123, 4, 10, 6, 1500)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "wgan_small.npz"), **out)
    print("wrote wgan_small.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()

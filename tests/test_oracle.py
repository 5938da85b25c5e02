"""CPU tests of the oracle (no GPU): hand-derived numpy fp64 vs independent torch.autograd fp64
(gradient penalty through autograd-of-autograd), golden fixtures, TF-Adam known answers, the FP32
torch CPU port used as the timed CPU baseline, host data-plane helpers."""
import os

import numpy as np
import pytest

from oracle import wgan_oracle as O
from oracle import wgan_autograd as AG

HERE = os.path.dirname(os.path.abspath(__file__))
DIMS = dict(noise_input_size=6, inflate_to_size=10, gex_size=13, disc_internal_size=7)
MID = dict(noise_input_size=20, inflate_to_size=40, gex_size=133, disc_internal_size=24)


def _inputs(cfg, B, seed):
    rng = np.random.default_rng(seed)
    x = rng.random((B, cfg.gex_size))
    z = np.abs(rng.normal(0, 0.1, (B, cfg.noise_input_size)) + rng.poisson(1.0, (B, cfg.noise_input_size)))
    return x, z, rng.random(B)


@pytest.mark.parametrize("dims,B", [(DIMS, 5), (MID, 33), (MID, 1)])
@pytest.mark.parametrize("lam", [0.0, 10.0])
def test_analytic_grads_match_autograd(dims, B, lam):
    cfg = O.OracleConfig(lambda_gp=lam, **dims)
    P = O.random_params(cfg, seed=3)
    x, z, eps = _inputs(cfg, B, 4)
    s, g = O.critic_step_grads(P, x, z, eps, cfg)
    s2, g2 = AG.critic_step_grads(P, x, z, eps, lam)
    for k in ("wasserstein", "gp", "critic_loss"):
        assert abs(s[k] - s2[k]) < 1e-12
    for k in O.CRITIC_NAMES:
        assert np.abs(g[k].reshape(g2[k].shape) - g2[k]).max() < 1e-12, k
    sg, gg = O.generator_step_grads(P, z, cfg)
    sg2, gg2 = AG.generator_step_grads(P, z)
    assert abs(sg["gen_loss"] - sg2["gen_loss"]) < 1e-12
    for k in O.GEN_NAMES:
        assert np.abs(gg[k].reshape(gg2[k].shape) - gg2[k]).max() < 1e-12, k


def test_reference_literal_losses():
    """[W:137-138]: obj_d = mean(D_fake) - mean(D_real), obj_g = -mean(D_fake); no penalty."""
    cfg = O.OracleConfig(**DIMS)
    assert cfg.lambda_gp == 0.0 and cfg.n_critic == 1
    P = O.random_params(cfg, seed=1)
    x, z, eps = _inputs(cfg, 9, 2)
    xt = O.generator_forward(P, z)[2]
    dr, df = O.critic_forward(P, x)[2], O.critic_forward(P, xt)[2]
    s, _ = O.critic_step_grads(P, x, z, None, cfg)
    assert s["critic_loss"] == pytest.approx(df.mean() - dr.mean(), abs=1e-14)
    sg, _ = O.generator_step_grads(P, z, cfg)
    assert sg["gen_loss"] == pytest.approx(-df.mean(), abs=1e-14)


def test_generator_output_is_leaky_relu():
    """[W:81]: the output layer is LeakyReLU too, so generated values may be slightly negative."""
    cfg = O.OracleConfig(**DIMS)
    P = O.random_params(cfg, seed=1)
    _, z, _ = _inputs(cfg, 64, 2)
    h1, h2, xt = O.generator_forward(P, z)
    pre = h2 @ P[O.GEN_NAMES[4]] + P[O.GEN_NAMES[5]]
    assert (xt < 0).any()
    assert np.allclose(xt[pre < 0], 0.2 * pre[pre < 0])


def test_tf_adam_known_answers():
    """SURVEY hard part: first-step update for g = 1e-9 is 3.15e-8 under TF-Adam (epsilon outside
    the bias correction), not torch.optim.Adam's 9.09e-7 (lr 1e-5... scaled: lr = 1e-3 here)."""
    var, m, v = np.zeros(1), np.zeros(1), np.zeros(1)
    O.tf_adam_update(var, np.array([1e-9]), m, v, 0.9, 0.999, 1e-3, 0.9, 0.999, 1e-8)
    # lr_t = lr*sqrt(1-b2)/(1-b1); m = 1e-10; sqrt(v) = 3.162e-11; update = lr_t*m/(sqrt(v)+eps)
    lr_t = 1e-3 * np.sqrt(1 - 0.999) / (1 - 0.9)
    assert -var[0] == pytest.approx(lr_t * 1e-10 / (np.sqrt(1e-21) + 1e-8), rel=1e-12)
    assert -var[0] == pytest.approx(3.15e-6, rel=2e-3)
    # large gradient: first step moves by ~lr * sign(g)
    var[:] = 0; m[:] = 0; v[:] = 0
    O.tf_adam_update(var, np.array([-2.0]), m, v, 0.9, 0.999, 1e-3, 0.9, 0.999, 1e-8)
    assert var[0] == pytest.approx(1e-3, rel=1e-6)


def test_adam_powers_advance_after_all_variables():
    """[A:214-225]: beta powers start at beta [A:123-128] and are multiplied once per apply."""
    cfg = O.OracleConfig(**DIMS)
    P = O.random_params(cfg, seed=1)
    opt = O.OracleAdam(O.CRITIC_NAMES, P, cfg)
    assert opt.beta1_power == 0.9 and opt.beta2_power == 0.999
    g = {k: np.ones_like(P[k]) for k in O.CRITIC_NAMES}
    opt.apply(P, g)
    assert opt.beta1_power == pytest.approx(0.81) and opt.beta2_power == pytest.approx(0.998001)


def test_golden_fixture():
    gold = np.load(os.path.join(HERE, "golden", "wgan_small.npz"))
    cfg = O.OracleConfig(lambda_gp=10.0, learn_rate=1e-3, **DIMS)
    P = {k[2:]: gold[k] for k in gold.files if k.startswith("P/")}
    x, z, eps = gold["x"], gold["z"], gold["eps"]
    sc, gc = O.critic_step_grads(P, x, z, eps, cfg)
    sg, gg = O.generator_step_grads(P, z, cfg)
    assert np.allclose([sc["wasserstein"], sc["gp"], sc["critic_loss"], sg["gen_loss"]], gold["scalars"],
                       rtol=1e-12, atol=1e-14)
    for k, v in {**gc, **gg}.items():
        assert np.allclose(v, gold["G/" + k], rtol=1e-11, atol=1e-14), k
    tr = O.OracleTrainer(cfg, P)
    for _ in range(3):
        tr.critic_step(x, z, eps)
        tr.generator_step(z)
    for k in O.PARAM_NAMES:
        assert np.allclose(tr.P[k], gold["P3/" + k], rtol=1e-11, atol=1e-14), k
    assert np.allclose(tr.sample(z), gold["sample"], rtol=1e-11, atol=1e-14)
    assert np.allclose(O.rng_noise_prior(123, 4, 10, 6, 5, 0.1), gold["noise"], atol=1e-6)
    assert np.array_equal(O.rng_uniform(123, 4, 10, 6), gold["uniform"])
    assert np.array_equal(O.rng_indices(123, 4, 10, 6, 1500), gold["indices"])


def test_torch_cpu_port_matches_oracle():
    import torch
    from oracle.wgan_torch_cpu import TorchCpuTrainer
    cfg = O.OracleConfig(lambda_gp=10.0, learn_rate=1e-3, **MID)
    P = O.random_params(cfg, seed=5)
    x, z, eps = _inputs(cfg, 17, 6)
    t = TorchCpuTrainer(cfg, P, dtype=torch.float64)
    tx, tz, te = (torch.tensor(a) for a in (x, z, eps))
    s, g = t.critic_step(tx, tz, te, apply=False)
    so, go = O.critic_step_grads(P, x, z, eps, cfg)
    assert s["critic_loss"] == pytest.approx(so["critic_loss"], rel=1e-12)
    for k in O.CRITIC_NAMES:
        assert np.abs(g[k].numpy().reshape(go[k].shape) - go[k]).max() < 1e-12
    s, g = t.generator_step(tz, apply=False)
    so, go = O.generator_step_grads(P, z, cfg)
    assert s["gen_loss"] == pytest.approx(so["gen_loss"], rel=1e-12)
    for k in O.GEN_NAMES:
        assert np.abs(g[k].numpy().reshape(go[k].shape) - go[k]).max() < 1e-12
    # three full reference-literal iterations including Adam
    ot = O.OracleTrainer(cfg, P)
    for _ in range(3):
        t.reference_iteration(tx, tz, te)
        ot.reference_iteration(x, z, eps)
    for k in O.PARAM_NAMES:
        assert np.abs(t.P[k].numpy() - ot.P[k]).max() < 1e-10, k


def test_noise_prior_distribution():
    """[W:91-94]: abs(N(0, 0.1) + Poisson(1)); SURVEY a12: mean ~1.03, std ~0.975."""
    zn = O.rng_noise_prior(1, 0, 0, 4000, 100, 0.1)
    assert zn.dtype == np.float32 and zn.min() >= 0
    assert zn.mean() == pytest.approx(1.03, abs=0.02)
    assert zn.std() == pytest.approx(0.975, abs=0.03)
    u = O.rng_uniform(1, 0, 0, 100000)
    assert 0 <= u.min() and u.max() < 1 and u.mean() == pytest.approx(0.5, abs=0.01)
    # disjoint row ranges of the same call concatenate to the full draw (rank sharding)
    a = O.rng_noise_prior(9, 3, 0, 64, 10, 0.1)
    b = np.concatenate([O.rng_noise_prior(9, 3, 0, 32, 10, 0.1), O.rng_noise_prior(9, 3, 32, 32, 10, 0.1)])
    assert np.array_equal(a, b)


def test_minmax_scale_genes():
    """[W:34-39]: per-gene min-max over all cells; constant genes map to 0."""
    m = np.array([[1.0, 3.0, 2.0], [5.0, 5.0, 5.0], [0.0, 0.0, 4.0]])
    s, lo, rng = O.minmax_scale_genes(m)
    assert np.allclose(s[0], [0, 1, 0.5]) and np.allclose(s[1], 0) and np.allclose(s[2], [0, 0, 1])
    raw = O.synthetic_ltpm(50, 40, 0)
    s, _, _ = O.minmax_scale_genes(raw)
    assert s.min() == 0 and s.max() == 1 and raw.min() >= 0


def test_gram_form_identities_match_the_explicit_chain():
    """The Gram form the CUDA path offers (gp_form=1; SURVEY 8d shortcuts i + ii) restated in numpy fp64, step by
    step as csrc/engine.cu::critic_grads_gram evaluates it, against the explicit chain of the oracle: same penalty,
    same per-sample norms, same parameter gradients (exact in real arithmetic, <= 1e-12 here)."""
    cfg = O.OracleConfig(noise_input_size=9, inflate_to_size=14, gex_size=37, disc_internal_size=11, lambda_gp=10.0)
    P = O.random_params(cfg, seed=3)
    rng = np.random.default_rng(5)
    B = 23
    x = rng.random((B, cfg.gex_size))
    xt = O.generator_forward(P, np.abs(rng.normal(1.0, 0.5, (B, cfg.noise_input_size))))[2]
    eps = rng.random(B)
    gp, g, n = O.gradient_penalty(P, x, xt, eps, cfg.lambda_gp)

    W1, b1, W2, b2, w3 = (P[k] for k in O.CRITIC_NAMES[:5])
    a = 0.2
    e = eps.reshape(-1, 1)
    # (i) layer 1 is linear before its activation: blend the pre-activations, x_hat is never formed
    h1 = O.lrelu(e * (x @ W1) + (1.0 - e) * (xt @ W1) + b1, a)
    h2 = O.lrelu(h1 @ W2 + b2, a)
    s1, s2 = O.slope(h1, a), O.slope(h2, a)
    g2 = np.ones((B, 1)) @ w3.T * s2
    g1 = (g2 @ W2.T) * s1
    # (ii) K = W1^T W1: norms, v1 and the direct W1 term without any product over the gene axis per sample
    K = W1.T @ W1
    T = g1 @ K
    n_gram = np.sqrt((T * g1).sum(axis=1))
    c = cfg.lambda_gp * 2.0 / B * (n_gram - 1.0) / n_gram
    v1 = c[:, None] * T * s1
    S = (c[:, None] * g1).T @ g1
    dW1 = W1 @ S
    dW2 = v1.T @ g2
    dw3 = ((v1 @ W2) * s2).sum(axis=0).reshape(-1, 1)
    gp_gram = cfg.lambda_gp / B * ((n_gram - 1.0) ** 2).sum()

    assert np.allclose(n_gram, n, rtol=1e-12, atol=0)
    assert abs(gp_gram - gp) <= 1e-12 * abs(gp)
    for got, key in ((dW1, O.CRITIC_NAMES[0]), (dW2, O.CRITIC_NAMES[2]), (dw3, O.CRITIC_NAMES[4])):
        assert np.abs(got - g[key]).max() <= 1e-12 * np.abs(g[key]).max(), key

"""CPU oracle for the WGAN(-GP) hot path of scripts/WGAN-GP_minimal.py.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this module; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may use it, and there only as the checker / the timed CPU baseline.

PARITY UNPINNED BY THE REFERENCE: the reference ships no tests, seeds, golden vectors or
recorded outputs, TensorFlow 1.x cannot be installed here (Python 3.12, no network) and the
bundled CSV is a Git-LFS pointer.  This file is therefore a *restatement* of the reference
graph, line-by-line traceable to the citations below, and it is cross-checked against an
independent ``torch.autograd`` fp64 implementation (``oracle/wgan_autograd.py``).

Citations:  [W:n] = /root/reference/scripts/WGAN-GP_minimal.py line n
            [A:n] = /root/reference/scripts/Adam_prediction.py line n

The arithmetic lives in TensorFlow (un-vendored, unpinned, inferred 1.4 <= TF <= 1.6):
  tf.layers.dense      y = x @ kernel + bias, kernel [in, out]        [W:65,71,77,107,113,119]
  tf.nn.leaky_relu     max(a, 0.2 a); gradient 1 if a > 0 else 0.2    [W:69,75,81,111,117]
  tf.reduce_mean                                                       [W:137-138]
  training_ops.apply_adam (ApplyAdam functor)                          [A:140-151]
"""
from __future__ import annotations

import dataclasses
import numpy as np

# TF variable creation order [W:128-130] -> generator first, then discriminator.
PARAM_NAMES = (
    "generator/gen_dense1/kernel", "generator/gen_dense1/bias",
    "generator/gen_dense2/kernel", "generator/gen_dense2/bias",
    "generator/gen_output/kernel", "generator/gen_output/bias",
    "discriminator/disc_dense1/kernel", "discriminator/disc_dense1/bias",
    "discriminator/disc_dense2/kernel", "discriminator/disc_dense2/bias",
    "discriminator/disc_output/kernel", "discriminator/disc_output/bias",
)
GEN_NAMES = PARAM_NAMES[:6]      # g_params  [W:142]
CRITIC_NAMES = PARAM_NAMES[6:]   # d_params  [W:141]


@dataclasses.dataclass
class OracleConfig:
    noise_input_size: int = 100      # [W:13]
    inflate_to_size: int = 600       # [W:14]
    gex_size: int = 6605             # [W:15]
    disc_internal_size: int = 200    # [W:17]
    learn_rate: float = 1e-5         # [W:22]
    beta1: float = 0.9               # [W:154-155]
    beta2: float = 0.999             # [W:154-155]
    epsilon: float = 1e-8            # [A:38]
    alpha: float = 0.2               # tf.nn.leaky_relu default
    lambda_gp: float = 0.0           # 0 => reference-literal (no GP term exists upstream)
    n_critic: int = 1                # [W:196,199] one critic update per iteration

    def shapes(self):
        Z, H, G, D = (self.noise_input_size, self.inflate_to_size, self.gex_size,
                      self.disc_internal_size)
        return {
            PARAM_NAMES[0]: (Z, H), PARAM_NAMES[1]: (H,),
            PARAM_NAMES[2]: (H, H), PARAM_NAMES[3]: (H,),
            PARAM_NAMES[4]: (H, G), PARAM_NAMES[5]: (G,),
            PARAM_NAMES[6]: (G, D), PARAM_NAMES[7]: (D,),
            PARAM_NAMES[8]: (D, D), PARAM_NAMES[9]: (D,),
            PARAM_NAMES[10]: (D, 1), PARAM_NAMES[11]: (1,),
        }


def glorot_init(cfg: OracleConfig, seed: int, dtype=np.float64):
    """tf.layers.dense defaults: glorot_uniform kernel, zeros bias."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shp in cfg.shapes().items():
        if name.endswith("kernel"):
            lim = np.sqrt(6.0 / (shp[0] + shp[1]))
            out[name] = rng.uniform(-lim, lim, size=shp).astype(dtype)
        else:
            out[name] = np.zeros(shp, dtype=dtype)
    return out


def random_params(cfg: OracleConfig, seed: int, dtype=np.float64, bias_scale=0.05):
    """Glorot kernels plus NON-zero biases (parity tests must exercise the bias path)."""
    p = glorot_init(cfg, seed, dtype)
    rng = np.random.default_rng(seed + 7919)
    for name in p:
        if name.endswith("bias"):
            p[name] = (bias_scale * rng.standard_normal(p[name].shape)).astype(dtype)
    return p


# ----------------------------------------------------------------------------------------
# Forward                                                                      [W:61-124]
# ----------------------------------------------------------------------------------------
def lrelu(a, alpha):
    return np.maximum(a, a.dtype.type(alpha) * a)      # TF 1.4: maximum(alpha*x, x)


def slope(h, alpha):
    """phi'(a) recovered from the post-activation (phi is sign preserving); phi'(0)=alpha."""
    return np.where(h > 0, h.dtype.type(1.0), h.dtype.type(alpha))


def generator_forward(P, z, alpha=0.2):
    """[W:61-83]; the output layer is also LeakyReLU [W:81]."""
    h1 = lrelu(z @ P[GEN_NAMES[0]] + P[GEN_NAMES[1]], alpha)
    h2 = lrelu(h1 @ P[GEN_NAMES[2]] + P[GEN_NAMES[3]], alpha)
    xt = lrelu(h2 @ P[GEN_NAMES[4]] + P[GEN_NAMES[5]], alpha)
    return h1, h2, xt


def critic_forward(P, x, alpha=0.2):
    """[W:103-124]; linear output, no sigmoid."""
    h1 = lrelu(x @ P[CRITIC_NAMES[0]] + P[CRITIC_NAMES[1]], alpha)
    h2 = lrelu(h1 @ P[CRITIC_NAMES[2]] + P[CRITIC_NAMES[3]], alpha)
    d = h2 @ P[CRITIC_NAMES[4]] + P[CRITIC_NAMES[5]]
    return h1, h2, d


# ----------------------------------------------------------------------------------------
# First-order backward (what Optimizer.minimize builds, [W:154-155])     SURVEY A.3
# ----------------------------------------------------------------------------------------
def critic_backward(P, x, h1, h2, delta_d, alpha=0.2, need_dx=False):
    W1, W2, w3 = P[CRITIC_NAMES[0]], P[CRITIC_NAMES[2]], P[CRITIC_NAMES[4]]
    g = {}
    g[CRITIC_NAMES[4]] = h2.T @ delta_d
    g[CRITIC_NAMES[5]] = delta_d.sum(axis=0)
    d2 = (delta_d @ w3.T) * slope(h2, alpha)
    g[CRITIC_NAMES[2]] = h1.T @ d2
    g[CRITIC_NAMES[3]] = d2.sum(axis=0)
    d1 = (d2 @ W2.T) * slope(h1, alpha)
    g[CRITIC_NAMES[0]] = x.T @ d1
    g[CRITIC_NAMES[1]] = d1.sum(axis=0)
    dx = d1 @ W1.T if need_dx else None
    return g, dx


def generator_backward(P, z, h1, h2, xt, delta_xt, alpha=0.2, need_dz=False):
    W1, W2, W3 = P[GEN_NAMES[0]], P[GEN_NAMES[2]], P[GEN_NAMES[4]]
    g = {}
    d3 = delta_xt * slope(xt, alpha)
    g[GEN_NAMES[4]] = h2.T @ d3
    g[GEN_NAMES[5]] = d3.sum(axis=0)
    d2 = (d3 @ W3.T) * slope(h2, alpha)
    g[GEN_NAMES[2]] = h1.T @ d2
    g[GEN_NAMES[3]] = d2.sum(axis=0)
    d1 = (d2 @ W2.T) * slope(h1, alpha)
    g[GEN_NAMES[0]] = z.T @ d1
    g[GEN_NAMES[1]] = d1.sum(axis=0)
    dz = d1 @ W1.T if need_dz else None
    return g, dz


# ----------------------------------------------------------------------------------------
# Gradient penalty — ABSENT upstream; canonical Gulrajani et al. form, A.4
# ----------------------------------------------------------------------------------------
def gradient_penalty(P, x, xt, eps, lam, alpha=0.2, global_batch=None):
    """Explicit analytic double-backprop chain.

    x_hat = eps*x + (1-eps)*xt ; gx = d/dx_hat sum_b D(x_hat)_b ; n_b = ||gx_b||_2
    GP = lam/B * sum_b (n_b-1)^2.  Returns (gp, grads wrt W1,W2,w3; bias grads are exactly 0,
    per-sample norms).
    """
    B = x.shape[0] if global_batch is None else global_batch
    dt = x.dtype.type
    W1, W2, w3 = P[CRITIC_NAMES[0]], P[CRITIC_NAMES[2]], P[CRITIC_NAMES[4]]
    e = eps.reshape(-1, 1).astype(x.dtype)
    xh = e * x + (dt(1.0) - e) * xt
    h1, h2, _ = critic_forward(P, xh, alpha)
    s1, s2 = slope(h1, alpha), slope(h2, alpha)
    # input-gradient pass
    g2 = np.ones((x.shape[0], 1), x.dtype) @ w3.T * s2
    g1 = (g2 @ W2.T) * s1
    gx = g1 @ W1.T
    # norm
    n = np.sqrt((gx * gx).sum(axis=1))
    gp = dt(lam) / dt(B) * ((n - dt(1.0)) ** 2).sum()
    # parameter-gradient-of-norm pass
    c = (dt(lam) * dt(2.0) / dt(B)) * (n - dt(1.0)) / n
    u = c[:, None] * gx
    g = {}
    g[CRITIC_NAMES[0]] = u.T @ g1
    v1 = (u @ W1) * s1
    g[CRITIC_NAMES[2]] = v1.T @ g2
    v2 = (v1 @ W2) * s2
    g[CRITIC_NAMES[4]] = v2.sum(axis=0).reshape(-1, 1)
    g[CRITIC_NAMES[1]] = np.zeros_like(P[CRITIC_NAMES[1]])
    g[CRITIC_NAMES[3]] = np.zeros_like(P[CRITIC_NAMES[3]])
    g[CRITIC_NAMES[5]] = np.zeros_like(P[CRITIC_NAMES[5]])
    return gp, g, n


def critic_step_grads(P, x, z, eps, cfg: OracleConfig, global_batch=None):
    """Gradients of obj_d (+GP) w.r.t. d_params; generator is a constant [W:137,141,154].

    Returns (scalars dict, grads dict).  With lambda_gp == 0 this is exactly the reference.
    """
    a = cfg.alpha
    B = x.shape[0] if global_batch is None else global_batch
    dt = x.dtype.type
    _, _, xt = generator_forward(P, z, a)
    h1r, h2r, dr = critic_forward(P, x, a)
    h1f, h2f, df = critic_forward(P, xt, a)
    wass = df.sum() / dt(B) - dr.sum() / dt(B)          # obj_d [W:137]
    inv = dt(1.0) / dt(B)
    gr, _ = critic_backward(P, x, h1r, h2r, np.full_like(dr, -inv), a)
    gf, _ = critic_backward(P, xt, h1f, h2f, np.full_like(df, inv), a)
    grads = {k: gr[k] + gf[k] for k in CRITIC_NAMES}
    gp = dt(0.0)
    if cfg.lambda_gp != 0.0:
        gp, gg, _ = gradient_penalty(P, x, xt, eps, cfg.lambda_gp, a, global_batch)
        for k in CRITIC_NAMES:
            grads[k] = grads[k] + gg[k]
    return {"wasserstein": wass, "gp": gp, "critic_loss": wass + gp}, grads


def generator_step_grads(P, z, cfg: OracleConfig, global_batch=None):
    """Gradients of obj_g = -mean(D(G(z))) w.r.t. g_params through the critic [W:138,142,155]."""
    a = cfg.alpha
    B = z.shape[0] if global_batch is None else global_batch
    dt = z.dtype.type
    gh1, gh2, xt = generator_forward(P, z, a)
    h1, h2, d = critic_forward(P, xt, a)
    gen_loss = -(d.sum() / dt(B))
    _, dxt = critic_backward(P, xt, h1, h2, np.full_like(d, -dt(1.0) / dt(B)), a, need_dx=True)
    grads, _ = generator_backward(P, z, gh1, gh2, xt, dxt, a)
    return {"gen_loss": gen_loss}, grads


# ----------------------------------------------------------------------------------------
# TF Adam                                             [A:110-151] + [A:214-225], SURVEY A.5
# ----------------------------------------------------------------------------------------
def tf_adam_update(var, grad, m, v, beta1_power, beta2_power, lr, beta1, beta2, epsilon):
    """ApplyAdam functor, evaluated in var.dtype, in place.  epsilon sits OUTSIDE the bias
    correction ("epsilon hat"), so this is NOT torch.optim.Adam."""
    T = var.dtype.type
    alpha = T(lr) * np.sqrt(T(1) - T(beta2_power)) / (T(1) - T(beta1_power))
    m += (grad - m) * (T(1) - T(beta1))
    v += (grad * grad - v) * (T(1) - T(beta2))
    var -= (m * alpha) / (np.sqrt(v) + T(epsilon))


class OracleAdam:
    """One Adam_Prediction_Optimizer instance [A:32]: per-optimizer beta powers initialised to
    beta1/beta2 [A:123-128], zero m/v slots [A:130-132]; `prediction` is accepted and ignored
    because all gradients are dense."""

    def __init__(self, names, P, cfg: OracleConfig, prediction=False):
        self.names = names
        self.cfg = cfg
        dt = P[names[0]].dtype.type
        self.m = {k: np.zeros_like(P[k]) for k in names}
        self.v = {k: np.zeros_like(P[k]) for k in names}
        self.beta1_power = dt(cfg.beta1)
        self.beta2_power = dt(cfg.beta2)
        self.prediction = prediction

    def apply(self, P, grads):
        c = self.cfg
        for k in self.names:                                    # _apply_dense [A:140-151]
            tf_adam_update(P[k], grads[k].astype(P[k].dtype), self.m[k], self.v[k],
                           self.beta1_power, self.beta2_power,
                           c.learn_rate, c.beta1, c.beta2, c.epsilon)
        dt = type(self.beta1_power)                              # _finish [A:214-225]
        self.beta1_power = dt(self.beta1_power * dt(c.beta1))
        self.beta2_power = dt(self.beta2_power * dt(c.beta2))


class OracleTrainer:
    def __init__(self, cfg: OracleConfig, params):
        self.cfg = cfg
        self.P = {k: v.copy() for k, v in params.items()}
        self.opt_critic = OracleAdam(CRITIC_NAMES, self.P, cfg, prediction=True)   # [W:154]
        self.opt_gen = OracleAdam(GEN_NAMES, self.P, cfg, prediction=False)        # [W:155]

    def critic_step(self, x, z, eps=None, global_batch=None, apply=True):
        s, g = critic_step_grads(self.P, x, z, eps, self.cfg, global_batch)
        if apply:
            self.opt_critic.apply(self.P, g)
        return s, g

    def generator_step(self, z, global_batch=None, apply=True):
        s, g = generator_step_grads(self.P, z, self.cfg, global_batch)
        if apply:
            self.opt_gen.apply(self.P, g)
        return s, g

    def reference_iteration(self, x, z, eps=None):
        """[W:196] generator update first, then [W:199] critic update, same x and noise; the
        critic step sees the already-updated generator."""
        sg, _ = self.generator_step(z)
        sc, _ = self.critic_step(x, z, eps)
        return {**sg, **sc}

    def sample(self, z):
        return generator_forward(self.P, z, self.cfg.alpha)[2]              # [W:217]


# ----------------------------------------------------------------------------------------
# Latent-vector fit — build-defined workload: min_z sum_g (G(z)-x)^2, G frozen.
# ----------------------------------------------------------------------------------------
def latent_fit_grads(P, z, x, alpha=0.2):
    h1, h2, xt = generator_forward(P, z, alpha)
    r = xt - x
    loss = (r * r).sum(axis=1)
    _, dz = generator_backward(P, z, h1, h2, xt, r.dtype.type(2.0) * r, alpha, need_dz=True)
    return loss, dz


# ----------------------------------------------------------------------------------------
# Counter-based RNG the device path uses (Philox4x32-10), restated with numpy integers.
# Replaces the reference's unseeded global NumPy RNG [W:92-93,188]; distributions match:
#   noise_prior = |Normal(0, data_max/10) + Poisson(1)|                       [W:91-94]
#   idx = randint(N, size=B)                                                   [W:188]
# ----------------------------------------------------------------------------------------
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
STREAM_NOISE, STREAM_EPS, STREAM_IDX, STREAM_SAMPLE, STREAM_LATENT = 1, 2, 3, 4, 5
# P(Poisson(1) <= k), float32, k = 0..11
POISSON1_CDF = np.cumsum([np.exp(-1.0) / np.prod(np.arange(1, k + 1, dtype=np.float64))
                          for k in range(12)]).astype(np.float32)


def philox4x32(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3))
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _M0 * c0.astype(np.uint64)
            p1 = _M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32(k0 + _W0)
            k1 = np.uint32(k1 + _W1)
    return c0, c1, c2, c3


def _philox_elems(seed, stream, call_id, idx):
    idx = np.asarray(idx, dtype=np.uint64)
    return philox4x32((idx & np.uint64(0xFFFFFFFF)).astype(np.uint32),
                      (idx >> np.uint64(32)).astype(np.uint32),
                      np.full(idx.shape, np.uint32(call_id & 0xFFFFFFFF)),
                      np.full(idx.shape, np.uint32(stream)),
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)


def rng_noise_prior(seed, call_id, row0, nrows, dim, sigma, stream=STREAM_NOISE):
    """|N(0,sigma) + Poisson(1)|, float32, element index = global_row*dim + col."""
    idx = (np.arange(row0, row0 + nrows, dtype=np.uint64)[:, None] * np.uint64(dim)
           + np.arange(dim, dtype=np.uint64)[None, :])
    r0, r1, r2, _ = _philox_elems(seed, stream, call_id, idx)
    u1 = ((r0 >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -24)
    u2 = ((r1 >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -24)
    u3 = (r2 >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
    normal = np.sqrt(np.float32(-2.0) * np.log(u1)) * np.cos(np.float32(2.0 * np.pi) * u2)
    pois = (u3[..., None] >= POISSON1_CDF[None, None, :]).sum(axis=-1).astype(np.float32)
    return np.abs(np.float32(sigma) * normal.astype(np.float32) + pois).astype(np.float32)


def rng_uniform(seed, call_id, row0, nrows, stream=STREAM_EPS):
    """U[0,1) float32 (24-bit), one per global row."""
    idx = np.arange(row0, row0 + nrows, dtype=np.uint64)
    r0, _, _, _ = _philox_elems(seed, stream, call_id, idx)
    return (r0 >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)


def rng_indices(seed, call_id, row0, nrows, n_cells, stream=STREAM_IDX):
    """Uniform-with-replacement cell indices in [0, n_cells)  [W:188]."""
    idx = np.arange(row0, row0 + nrows, dtype=np.uint64)
    r0, _, _, _ = _philox_elems(seed, stream, call_id, idx)
    return ((r0.astype(np.uint64) * np.uint64(n_cells)) >> np.uint64(32)).astype(np.int64)


# ----------------------------------------------------------------------------------------
# Host data plane                                                             [W:32-42]
# ----------------------------------------------------------------------------------------
def minmax_scale_genes(genes_by_cells):
    """sklearn MinMaxScaler on the transposed matrix == per-gene min-max over all cells;
    constant genes map to 0 (sklearn replaces a zero range by 1).  [W:34-39]"""
    m = np.asarray(genes_by_cells, dtype=np.float64)
    lo = m.min(axis=1, keepdims=True)
    rng = m.max(axis=1, keepdims=True) - lo
    rng[rng == 0.0] = 1.0
    return (m - lo) / rng, lo[:, 0], rng[:, 0]


def synthetic_ltpm(n_genes, n_cells, seed):
    """log2(TPM+1)-like genes x cells matrix: per-gene dropout + gamma magnitude."""
    rng = np.random.default_rng(seed)
    p_drop = rng.uniform(0.3, 0.95, size=(n_genes, 1))
    scale = rng.uniform(2.0, 12.0, size=(n_genes, 1))
    expressed = rng.random((n_genes, n_cells)) > p_drop
    mag = rng.gamma(2.0, 1.0, size=(n_genes, n_cells)) * (scale / 2.0)
    return np.where(expressed, mag, 0.0)

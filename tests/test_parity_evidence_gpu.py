"""Parity evidence the round-1 review asked for, through the C ABI against the fp64 oracle:

* a MEASURED error table (relative Frobenius and max-normalised error of every gradient tensor, loss errors) for
  B in {32, 200, 4096} x {TF32 tensor-core, FP32 verification} x {explicit, Gram} penalty form, written to
  gpurun_out/r02_parity_report.json (copied to profiles/), with every entry held to the thresholds below;
* the latent-vector fit (BASELINE config 5) on the TENSOR-CORE path at reference dims;
* BASELINE config 3 (critic + penalty at batch 65536) against the oracle on sampled rows;
* padding of every arena stays exactly zero through training (the layout invariant DESIGN.md section 3 relies on);
* host-fed iterations with n_critic = 1 and no read-back (stream race found by the round-1 advisor)."""
import json
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import wgan_oracle as O  # noqa: E402
from test_steps_gpu import _setup, _fro, _relerr, _dscale, REF, SMALL  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# Stated tolerances = twice the worst value measured on B200 over both penalty forms (profiles/r02_parity_report.json
# holds the measured numbers; the runs are deterministic).  A flipped LeakyReLU slope moves a batch-summed gradient
# entry by O(1/B): that is what the TF32 rows (10-bit operand mantissa) show, falling as ~1/sqrt(B), and it is also why
# the FP32 verification mode is not at 1e-7 for B >= 200 (one unit within fp32 rounding of zero flips).
#   (mode, B): worst relative Frobenius error of a gradient tensor, worst max-normalised entry error,
#              |Wasserstein error| / (|mean D_fake| + |mean D_real|), penalty relative error, generator-loss relative error
MEASURED = {
    (1, 32):   dict(fro=1.2e-6, max_rel=1.4e-6, w=1.3e-7, gp=2.0e-6, gen=3e-9),
    (1, 200):  dict(fro=9.3e-5, max_rel=9.8e-4, w=3e-8,   gp=4.4e-6, gen=3e-8),
    (1, 4096): dict(fro=7.3e-5, max_rel=2.7e-4, w=4e-8,   gp=4.5e-5, gen=4e-8),
    (0, 32):   dict(fro=1.45e-2, max_rel=5.95e-2, w=1.2e-4, gp=3.8e-3, gen=1.1e-4),
    (0, 200):  dict(fro=4.6e-3,  max_rel=1.7e-2,  w=1.4e-4, gp=2.9e-3, gen=1.1e-4),
    (0, 4096): dict(fro=1.15e-3, max_rel=1.7e-3,  w=1.6e-4, gp=1.35e-3, gen=6.5e-5),
}
TOL = {k: {q: 2.0 * v for q, v in m.items()} for k, m in MEASURED.items()}


def _measure(B, mode, form):
    tr, cfg, ocfg, P, (x, z, eps), (xd, zd, ed) = _setup(REF, B, 10.0, mode, gp_form=form)
    rec = {"batch": B, "mode": "tf32" if mode == 0 else "fp32", "gp_form": form, "tensors": {}}
    s = tr.critic_step(xd, zd, ed, apply=False)
    so, go = O.critic_step_grads(P, x, z, eps, ocfg)
    rec["wasserstein_err"] = abs(s["wasserstein"] - so["wasserstein"]) / _dscale(P, x, z)
    rec["gp_rel_err"] = abs(s["gp"] - so["gp"]) / abs(so["gp"])
    g = tr.get_grads(O.CRITIC_NAMES)
    for k in O.CRITIC_NAMES:
        if np.abs(go[k]).max() > 0:
            rec["tensors"][k] = {"fro": _fro(g[k], go[k]), "max_rel": _relerr(g[k], go[k])}
        else:
            assert np.abs(g[k]).max() == 0, k
    s2 = tr.generator_step(zd, apply=False)
    so2, go2 = O.generator_step_grads(P, z, ocfg)
    rec["gen_loss_rel_err"] = abs(s2["gen_loss"] - so2["gen_loss"]) / abs(so2["gen_loss"])
    g2 = tr.get_grads(O.GEN_NAMES)
    for k in O.GEN_NAMES:
        rec["tensors"][k] = {"fro": _fro(g2[k], go2[k]), "max_rel": _relerr(g2[k], go2[k])}
    tr.close()
    return rec


def test_measured_error_table():
    report = []
    for B in (32, 200, 4096):
        for mode in (1, 0):
            for form in ("explicit", "gram"):
                report.append(_measure(B, mode, form))
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "r02_parity_report.json"), "w") as f:
        json.dump({"what": "measured error of the CUDA path vs the fp64 oracle, reference dims, lambda 10",
                   "thresholds": "tests/test_parity_evidence_gpu.py", "cases": report}, f, indent=1)
    for rec in report:
        B, mode = rec["batch"], 0 if rec["mode"] == "tf32" else 1
        where, tol = (B, rec["mode"], rec["gp_form"]), TOL[(mode, B)]
        assert rec["wasserstein_err"] <= tol["w"], (where, rec["wasserstein_err"])
        assert rec["gen_loss_rel_err"] <= tol["gen"], (where, rec["gen_loss_rel_err"])
        assert rec["gp_rel_err"] <= tol["gp"], (where, rec["gp_rel_err"])
        for k, e in rec["tensors"].items():
            assert e["fro"] <= tol["fro"], (where, k, e)
            assert e["max_rel"] <= tol["max_rel"], (where, k, e)


def test_latent_fit_tensor_core_path_reference_dims():
    """BASELINE config 5 latent fit through wgan_latent_fit_step in TF32 mode at reference dims: per-cell loss and
    the direction of the first TF-Adam step against O.latent_fit_grads (round 1 only covered the FP32 fallback)."""
    B = 256
    tr, cfg, ocfg, P, (x, z, eps), (xd, zd, ed) = _setup(REF, B, 0.0, 0)
    target = O.generator_forward(P, np.abs(z[::-1]).copy())[2]
    td = torch.tensor(target, dtype=torch.float32, device="cuda:0")
    z1, loss, m1, _ = tr.latent_fit(td, steps=1, learn_rate=1e-2, z0=zd, return_state=True)
    lo, dz = O.latent_fit_grads(P, z, target.astype(np.float32).astype(np.float64))
    assert _relerr(loss.cpu().numpy(), lo) < 2e-3
    # the first TF-Adam step from zero state leaves m = (1 - beta1) dL/dz: the device's latent gradient
    dz_dev = m1.cpu().numpy().astype(np.float64) / (1.0 - cfg.beta1)
    # measured 8.5e-3 (per-row gradients: a flipped LeakyReLU slope is not averaged over the batch here); bound = 2x
    assert _fro(dz_dev, dz) < 1.7e-2, _fro(dz_dev, dz)
    # ... and moves every coordinate whose gradient is not tiny by lr against its sign [A:56-61]
    step = (zd - z1).cpu().numpy()
    big = np.abs(dz_dev) > 1e-3 * np.abs(dz_dev).max()
    assert np.array_equal(np.sign(step[big]), np.sign(dz_dev[big]))
    assert np.allclose(np.abs(step[big]), 1e-2, rtol=1e-3)
    # and the fit converges on the tensor-core path
    z2, loss2 = tr.latent_fit(td, steps=150, learn_rate=1e-2, z0=zd)
    assert float(loss2.mean()) < 0.5 * float(loss.mean())
    tr.close()


def test_config3_stress_batch_rows_match_oracle():
    """BASELINE config 3: one critic + penalty step at batch 65536, reference dims, tensor cores.  The oracle cannot
    run the whole batch in fp64 in test time, but every per-row quantity is row-local: D(G(z_b)), D(x_b) and the
    penalty norm ||grad_x^ D||_2 of 256 sampled rows are compared with the oracle, and the loss scalars with the
    exact means over those quantities computed from the device's own per-row outputs."""
    from arshamg_scrnaseq_wgan_b200 import WGANConfig, WGANTrainer
    B, S = 65536, 256
    cfg = WGANConfig(lambda_gp=10.0)
    tr = WGANTrainer(cfg, max_batch=B, device=0)
    ocfg = O.OracleConfig(lambda_gp=10.0)
    P = O.random_params(ocfg, seed=9, dtype=np.float64)
    P = {k: v.astype(np.float32).astype(np.float64) for k, v in P.items()}
    tr.set_params(P)
    g = torch.Generator(device="cuda:0").manual_seed(1)
    x = torch.rand((B, cfg.gex_size), device="cuda:0", generator=g) * (torch.rand((B, cfg.gex_size), device="cuda:0", generator=g) > 0.5)
    z = tr.noise_prior(B, seed=5, call_id=1, row0=0)
    eps = tr.uniform(B, seed=5, call_id=1, row0=0)
    s = tr.critic_step(x, z, eps, apply=False)
    rows = np.sort(np.random.default_rng(0).choice(B, S, replace=False))
    d = tr.buffer("d").cpu().numpy()[0]
    norms = tr.buffer("norms").cpu().numpy()[0][:B]
    d_fake, d_real = d[:B], d[2 * B:3 * B]
    xs, zs, es = (t[rows].cpu().numpy().astype(np.float64) for t in (x, z, eps))
    xt_o = O.generator_forward(P, zs)[2]
    df_o = O.critic_forward(P, xt_o)[2].ravel()
    dr_o = O.critic_forward(P, xs)[2].ravel()
    _, _, n_o = O.gradient_penalty(P, xs, xt_o, es, 10.0)
    scale = np.abs(df_o).mean() + np.abs(dr_o).mean()
    assert np.abs(d_fake[rows] - df_o).max() <= 2e-3 * scale
    assert np.abs(d_real[rows] - dr_o).max() <= 2e-3 * scale
    rel = np.abs(norms[rows] - n_o) / n_o
    assert np.median(rel) < 1e-3 and rel.max() < 5e-2, (np.median(rel), rel.max())
    # the reduced scalars are the means of the per-row outputs (fp64 on the host)
    assert abs(s["wasserstein"] - (d_fake.astype(np.float64).mean() - d_real.astype(np.float64).mean())) <= 2e-5 * scale
    assert abs(s["gp"] - 10.0 * ((norms.astype(np.float64) - 1.0) ** 2).mean()) <= 1e-4 * s["gp"]
    # the O(Hd^2) gradients need the whole batch: checked through shard linearity in test_steps_gpu.py
    tr.close()


@pytest.mark.parametrize("form", ["explicit", "gram"])
def test_arena_padding_stays_zero(form):
    """Leading dimensions are padded to 4 floats (6605 -> 6608) and tensors to 256 bytes; TMA stores clip to the
    logical extent and Adam runs over the whole arena, so after training every pad float of params / grads / m / v
    must still be exactly zero."""
    from arshamg_scrnaseq_wgan_b200 import WGANConfig, WGANTrainer, PARAM_NAMES
    dims = dict(noise_input_size=10, inflate_to_size=30, gex_size=133, disc_internal_size=24)   # 133 -> ld 136, 10 -> 12
    B = 64
    tr = WGANTrainer(WGANConfig(lambda_gp=10.0, learn_rate=1e-3, gp_form=form, **dims), max_batch=B, device=0)
    tr.init_params(seed=1)
    tr.set_data(torch.rand((300, dims["gex_size"])))
    for it in range(10):
        tr.train_iteration_resident(it, B, seed=3)
    torch.cuda.synchronize()
    npad = 0
    for net in (0, 1):
        n_floats = tr.arena(net, 0).numel()
        live = np.zeros(n_floats, bool)
        for name in PARAM_NAMES:
            i, tnet, rows, cols, ld, off = tr.info[name]
            if tnet != net:
                continue
            idx = off + (np.arange(rows)[:, None] * ld + np.arange(cols)[None, :]).ravel()
            live[idx] = True
        live[-4:] = True                                     # loss scalars in the tail of the gradient arena
        npad += int((~live).sum())
        for kind in range(4):
            a = tr.arena(net, kind).cpu().numpy()
            assert np.all(a[~live] == 0.0), (net, kind, np.flatnonzero(a[~live])[:5])
            assert np.isfinite(a).all()
    assert npad > 0
    tr.close()


def test_host_fed_single_critic_update_without_readback():
    """n_critic = 1 (reference-literal ratio) and read=False: the copy stream must not overwrite the generator's
    latent while the previous iteration's generator step still reads it.  Compared with the synchronous path."""
    from arshamg_scrnaseq_wgan_b200 import WGANConfig, WGANTrainer, PARAM_NAMES
    B, iters = 2048, 6
    cfg = WGANConfig(learn_rate=1e-3, n_critic=1, lambda_gp=10.0)
    g = torch.Generator().manual_seed(0)
    data = torch.rand((256, cfg.gex_size), generator=g)
    feeds = []
    for _ in range(iters):
        idx = torch.randint(256, (B,), generator=g)
        feeds.append((idx.pin_memory(), data[idx].contiguous().pin_memory(),
                      torch.rand((B, cfg.noise_input_size), generator=g).pin_memory(),
                      torch.rand((B,), generator=g).pin_memory(),
                      torch.rand((B, cfg.noise_input_size), generator=g).pin_memory()))
    outs = []
    for mode in ("sync", "x", "idx"):
        tr = WGANTrainer(cfg, max_batch=B, device=0, gemm_mode=1)
        tr.init_params(seed=2)
        tr.set_data(data)
        for idx, xh, zh, eh, zg in feeds:
            if mode == "sync":
                tr.critic_step(xh.cuda(), zh.cuda(), eh.cuda(), read=False)
                tr.generator_step(zg.cuda(), read=False)
                torch.cuda.synchronize()
            elif mode == "x":
                tr.train_iteration_host([xh], [zh], [eh], zg, read=False)
            else:
                tr.train_iteration_host_indices([idx], [zh], [eh], zg, read=False)
        torch.cuda.synchronize()
        outs.append(tr.get_params())
        tr.close()
    for k in PARAM_NAMES:
        assert np.array_equal(outs[0][k], outs[1][k]), k
        assert np.array_equal(outs[0][k], outs[2][k]), k

"""world_size-2 gloo test (CPU) of the data-parallel partition: every rank evaluates
its batch shard with the GLOBAL 1/B in the per-sample coefficients, a single SUM all-reduce of
the flat gradient + scalar buffer reproduces the 1-process gradient of the full batch, and the
replicated Adam update keeps the ranks bit-identical."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from oracle import wgan_oracle as O  # noqa: E402

DIMS = dict(noise_input_size=6, inflate_to_size=10, gex_size=13, disc_internal_size=7)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = O.OracleConfig(lambda_gp=10.0, learn_rate=1e-3, **DIMS)
    P = O.random_params(cfg, seed=11)
    B = 12
    rng = np.random.default_rng(0)
    x = rng.random((B, cfg.gex_size))
    z = np.abs(rng.normal(0, 0.1, (B, cfg.noise_input_size)) + rng.poisson(1.0, (B, cfg.noise_input_size)))
    eps = rng.random(B)
    sl = slice(rank * B // world, (rank + 1) * B // world)
    tr = O.OracleTrainer(cfg, P)
    for step in range(2):
        s, g = tr.critic_step(x[sl], z[sl], eps[sl], global_batch=B, apply=False)
        flat = np.concatenate([g[k].ravel() for k in O.CRITIC_NAMES] + [[s["wasserstein"], s["gp"]]])
        t = torch.tensor(flat)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)            # ONE collective per optimizer step
        flat = t.numpy()
        off = 0
        gsum = {}
        for k in O.CRITIC_NAMES:
            n = g[k].size
            gsum[k] = flat[off:off + n].reshape(g[k].shape)
            off += n
        tr.opt_critic.apply(tr.P, gsum)
        s2, g2 = tr.generator_step(z[sl], global_batch=B, apply=False)
        t = torch.tensor(np.concatenate([g2[k].ravel() for k in O.GEN_NAMES] + [[s2["gen_loss"]]]))
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        flat2 = t.numpy()
        off = 0
        gsum2 = {}
        for k in O.GEN_NAMES:
            n = g2[k].size
            gsum2[k] = flat2[off:off + n].reshape(g2[k].shape)
            off += n
        tr.opt_gen.apply(tr.P, gsum2)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), scal=flat[-2:], gen_loss=flat2[-1:],
             **{k.replace("/", "__"): v for k, v in tr.P.items()})
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    # single-process reference on the full batch
    cfg = O.OracleConfig(lambda_gp=10.0, learn_rate=1e-3, **DIMS)
    P = O.random_params(cfg, seed=11)
    B = 12
    rng = np.random.default_rng(0)
    x = rng.random((B, cfg.gex_size))
    z = np.abs(rng.normal(0, 0.1, (B, cfg.noise_input_size)) + rng.poisson(1.0, (B, cfg.noise_input_size)))
    eps = rng.random(B)
    tr = O.OracleTrainer(cfg, P)
    for _ in range(2):
        s, _ = tr.critic_step(x, z, eps)
        sg, _ = tr.generator_step(z)
    assert np.allclose(r0["scal"], [s["wasserstein"], s["gp"]], rtol=1e-9)
    assert np.allclose(r0["gen_loss"], sg["gen_loss"], rtol=1e-9)
    for k in O.PARAM_NAMES:
        kk = k.replace("/", "__")
        assert np.array_equal(r0[kk], r1[kk]), k                       # replicas stay identical
        assert np.allclose(r0[kk], tr.P[k], rtol=1e-6, atol=1e-9), k   # == 1-process, sum order aside

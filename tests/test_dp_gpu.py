"""Data-parallel path behind the C ABI (wgan_dp_export / wgan_dp_connect / wgan_allreduce_grads): every rank runs
the real CUDA step on its batch shard, the library's own kernel sums the flat gradient arenas over peer memory
(one exchange per optimizer step [W:154-155]), Adam runs replicated.

* in-process groups (several handles, one host thread; all ranks on cuda:0 when the box has a single GPU, one rank
  per GPU otherwise): the ALL-REDUCED GRADIENT ARENA of every rank is compared directly, before Adam, with the
  arena of a single-handle run on the full batch -- replicas bit-identical, sums equal up to summation order;
* one process per GPU over cudaIpc handles, rendezvous through torch.distributed (needs >= 2 GPUs)."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

DIMS = dict(noise_input_size=20, inflate_to_size=40, gex_size=133, disc_internal_size=24)
REF = dict(noise_input_size=100, inflate_to_size=600, gex_size=6605, disc_internal_size=200)


def _inputs(cfg, B, dev):
    g = np.random.default_rng(0)
    x = torch.tensor(g.random((B, cfg.gex_size)), dtype=torch.float32, device=dev)
    z = torch.tensor(np.abs(g.normal(1, 0.5, (B, cfg.noise_input_size))), dtype=torch.float32, device=dev)
    eps = torch.tensor(g.random(B), dtype=torch.float32, device=dev)
    return x, z, eps


def _group(cfg, B, world, mode):
    """`world` connected handles in this process; rank r on GPU r % device_count."""
    from arshamg_scrnaseq_wgan_b200 import WGANTrainer
    ndev = torch.cuda.device_count()
    # ranks that share a GPU must all fit on it at once (their exchange kernels wait for each other)
    per_dev = -(-world // ndev)
    os.environ["WGAN_B200_DP_BLOCKS"] = str(max(1, min(32, 148 // per_dev // 2)))
    trs = [WGANTrainer(cfg, max_batch=B // world, device=r % ndev, gemm_mode=mode, dp_rank=r, dp_world=world)
           for r in range(world)]
    blobs = [t.dp_export() for t in trs]
    for t in trs:
        t.init_params(seed=4)
        t.dp_connect(blobs)
    os.environ.pop("WGAN_B200_DP_BLOCKS", None)
    streams = [torch.cuda.Stream(device=f"cuda:{t.device}") for t in trs]
    return trs, streams


def _close(trs):
    torch.cuda.synchronize()
    for t in trs:
        t.close()


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("form", ["explicit", "gram"])
@pytest.mark.parametrize("mode", [1, 0])
def test_allreduced_gradient_arena_matches_single_handle(mode, form, world):
    from arshamg_scrnaseq_wgan_b200 import WGANConfig, WGANTrainer
    from arshamg_scrnaseq_wgan_b200.trainer import NET_CRITIC, NET_GENERATOR, KIND_GRAD
    B = 64
    cfg = WGANConfig(lambda_gp=10.0, learn_rate=1e-3, gp_form=form, **DIMS)
    one = WGANTrainer(cfg, max_batch=B, device=0, gemm_mode=mode, distributed=False)
    one.init_params(seed=4)
    trs, streams = _group(cfg, B, world, mode)
    tol = 2e-2 if mode == 0 else 2e-5
    for it in range(3):
        x, z, eps = _inputs(cfg, B, "cuda:0")
        x, z, eps = x + 0.01 * it, z + 0.01 * it, eps
        # ---- critic: gradients -> all-reduce; compare the arenas BEFORE Adam
        one.critic_step(x, z, eps, apply=False, read=False)
        ref = one.arena(NET_CRITIC, KIND_GRAD).cpu().numpy().copy()
        torch.cuda.synchronize()
        # shards first, then enqueue every rank's step without any host synchronisation in between: rank r's
        # exchange kernel spins on its stream until the other ranks' kernels have been launched
        shards = []
        for r, t in enumerate(trs):
            dev = f"cuda:{t.device}"
            sl = slice(r * B // world, (r + 1) * B // world)
            shards.append((x[sl].to(dev).contiguous(), z[sl].to(dev).contiguous(), eps[sl].to(dev).contiguous()))
        torch.cuda.synchronize()
        for (xs, zs, es), t, s in zip(shards, trs, streams):
            with torch.cuda.device(t.device), torch.cuda.stream(s):
                t.critic_step(xs, zs, es, apply=False, read=False, global_batch=B)
        torch.cuda.synchronize()
        got = [t.arena(NET_CRITIC, KIND_GRAD).cpu().numpy() for t in trs]
        for g in got[1:]:
            assert np.array_equal(got[0], g)                                   # replicas bit-identical
        # scalar tail (loss sums) and every gradient entry: equal to the single-handle arena up to summation order
        # (and, in TF32 mode, the slope flips of units within rounding of zero)
        assert np.linalg.norm(got[0] - ref) <= tol * np.linalg.norm(ref), (mode, form, world, it)
        if mode == 1:
            assert np.abs(got[0] - ref).max() <= 1e-5 * np.abs(ref).max()
        # ---- generator
        one.generator_step(z, apply=False, read=False)
        refg = one.arena(NET_GENERATOR, KIND_GRAD).cpu().numpy().copy()
        torch.cuda.synchronize()
        for (_, zs, _), t, s in zip(shards, trs, streams):
            with torch.cuda.device(t.device), torch.cuda.stream(s):
                t.generator_step(zs, apply=False, read=False, global_batch=B)
        torch.cuda.synchronize()
        gotg = [t.arena(NET_GENERATOR, KIND_GRAD).cpu().numpy() for t in trs]
        for g in gotg[1:]:
            assert np.array_equal(gotg[0], g)
        assert np.linalg.norm(gotg[0] - refg) <= tol * np.linalg.norm(refg), (mode, form, world, it)
        # ---- now apply Adam everywhere: parameters stay bit-identical across the replicas
        one.lib.wgan_critic_apply(one._h, None)
        one.lib.wgan_generator_apply(one._h, None)
        for t in trs:
            with torch.cuda.device(t.device):
                t.lib.wgan_critic_apply(t._h, None)
                t.lib.wgan_generator_apply(t._h, None)
        torch.cuda.synchronize()
    from arshamg_scrnaseq_wgan_b200 import PARAM_NAMES
    for k in PARAM_NAMES:
        p0 = trs[0].get_tensor(k)
        for t in trs[1:]:
            assert np.array_equal(p0, t.get_tensor(k)), k
    _close(trs)
    one.close()


def test_allreduce_reference_dims_is_exact_sum():
    """Reference-size arenas (critic 5.4 MB, generator 17.6 MB), 8 in-process ranks: the exchanged arena equals the
    rank-ordered FP32 sum of the per-rank arenas bit for bit, and a second exchange (epoch 2) works."""
    from arshamg_scrnaseq_wgan_b200 import WGANConfig
    from arshamg_scrnaseq_wgan_b200.trainer import NET_CRITIC, NET_GENERATOR, KIND_GRAD
    cfg = WGANConfig(lambda_gp=10.0, **REF)
    world = 8
    trs, streams = _group(cfg, 8 * world, world, 0)
    for net in (NET_CRITIC, NET_GENERATOR, NET_CRITIC):
        arenas = [t.arena(net, KIND_GRAD) for t in trs]
        gen = torch.Generator(device="cpu").manual_seed(net + 1)
        vals = [torch.randn(arenas[0].numel(), generator=gen) for _ in trs]
        for a, v in zip(arenas, vals):
            a.copy_(v.to(a.device))
        expect = vals[0].clone()
        for v in vals[1:]:
            expect += v                                                        # rank order, FP32
        torch.cuda.synchronize()
        for t, s in zip(trs, streams):
            with torch.cuda.device(t.device), torch.cuda.stream(s):
                t._allreduce(net)
        torch.cuda.synchronize()
        for a in arenas:
            assert torch.equal(a.cpu(), expect)
    _close(trs)


# ------------------------------------------------------------------------------------------------------------
# one process per GPU (cudaIpc)
# ------------------------------------------------------------------------------------------------------------
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir, mode, form):
    import torch.distributed as dist
    from arshamg_scrnaseq_wgan_b200 import WGANConfig, WGANTrainer, PARAM_NAMES
    from arshamg_scrnaseq_wgan_b200.trainer import NET_CRITIC, KIND_GRAD
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    B = 64
    cfg = WGANConfig(lambda_gp=10.0, learn_rate=1e-3, gp_form=form, **DIMS)
    tr = WGANTrainer(cfg, max_batch=B, device=rank, gemm_mode=mode)
    assert tr.world == world and tr.dp_backend == "engine"
    tr.init_params(seed=4)
    x, z, eps = _inputs(cfg, B, f"cuda:{rank}")
    sl = slice(rank * B // world, (rank + 1) * B // world)
    tr.critic_step(x[sl], z[sl], eps[sl], apply=False, read=False)
    torch.cuda.synchronize()
    arena = tr.arena(NET_CRITIC, KIND_GRAD).cpu().numpy().copy()
    for _ in range(3):
        sc = tr.critic_step(x[sl], z[sl], eps[sl])
        sg = tr.generator_step(z[sl])
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), scal=np.array([sc["wasserstein"], sc["gp"], sg["gen_loss"]]),
             arena=arena, **{k.replace("/", "__"): tr.get_tensor(k) for k in PARAM_NAMES})
    dist.barrier()
    tr.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("form", ["explicit", "gram"])
@pytest.mark.parametrize("mode", [1, 0])
def test_multi_process_ipc_matches_single_gpu(tmp_path, mode, form):
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    from arshamg_scrnaseq_wgan_b200 import WGANConfig, WGANTrainer, PARAM_NAMES
    from arshamg_scrnaseq_wgan_b200.trainer import NET_CRITIC, KIND_GRAD
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), mode, form), nprocs=world, join=True)
    ranks = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    B = 64
    cfg = WGANConfig(lambda_gp=10.0, learn_rate=1e-3, gp_form=form, **DIMS)
    tr = WGANTrainer(cfg, max_batch=B, device=0, gemm_mode=mode, distributed=False)
    tr.init_params(seed=4)
    x, z, eps = _inputs(cfg, B, "cuda:0")
    tr.critic_step(x, z, eps, apply=False, read=False)
    ref_arena = tr.arena(NET_CRITIC, KIND_GRAD).cpu().numpy().copy()
    for _ in range(3):
        sc = tr.critic_step(x, z, eps)
        sg = tr.generator_step(z)
    tol = 2e-2 if mode == 0 else 2e-5
    for r in ranks[1:]:
        assert np.array_equal(ranks[0]["arena"], r["arena"])                      # all-reduced gradients identical
    assert np.linalg.norm(ranks[0]["arena"] - ref_arena) <= tol * np.linalg.norm(ref_arena)
    assert np.allclose(ranks[0]["scal"], [sc["wasserstein"], sc["gp"], sg["gen_loss"]], rtol=max(tol, 1e-4), atol=tol * 0.1)
    for k in PARAM_NAMES:
        kk = k.replace("/", "__")
        for r in ranks[1:]:
            assert np.array_equal(ranks[0][kk], r[kk]), k                         # replicas identical after Adam

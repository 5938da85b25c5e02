"""GPU parity of the tcgen05 TF32 GEMM (and the FP32 verification GEMM) against numpy fp64,
through the C ABI (wgan_gemm).  Covers the three operand layouts of the dense layers
(forward / dgrad / wgrad), ragged shapes, split-K and every fused epilogue."""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

# Tolerances, normalised by max|ref|: TF32 has a 10-bit mantissa (operands rounded to nearest
# by TMA), FP32 accumulation.  FP32 mode differs from fp64 numpy only by summation rounding.
TOL = {0: 2e-3, 1: 2e-5}


def _trainer(mode, max_batch=64):
    from arshamg_scrnaseq_wgan_b200 import WGANConfig, WGANTrainer
    cfg = WGANConfig(noise_input_size=8, inflate_to_size=8, gex_size=8, disc_internal_size=8)
    return WGANTrainer(cfg, max_batch=max_batch, device=0, gemm_mode=mode)


def _pad(a, ld):
    out = np.zeros((a.shape[0], ld), np.float32)
    out[:, :a.shape[1]] = a
    return out


def _run(tr, layout, M, N, K, rng, bias=False, act=0, aux=False, row_scale=False, sumsq=False, splits=0):
    a_mn, b_mn = {"fwd": (0, 1), "dgrad": (0, 0), "wgrad": (1, 1)}[layout]
    A = rng.standard_normal((M, K))
    Bm = rng.standard_normal((N, K))
    r4 = lambda v: (v + 3) // 4 * 4
    A_store = _pad(A.T if a_mn else A, r4(M if a_mn else K))
    B_store = _pad(Bm.T if b_mn else Bm, r4(N if b_mn else K))
    dev = torch.device("cuda:0")
    Ad, Bd = torch.tensor(A_store, device=dev), torch.tensor(B_store, device=dev)
    ldd = r4(N)
    Dd = torch.full((M, ldd), float("nan"), device=dev)
    ref = A.astype(np.float32).astype(np.float64) @ Bm.astype(np.float32).astype(np.float64).T
    bias_d = aux_d = rs_d = sq_d = None
    alpha = tr.cfg.alpha
    if bias:
        b = rng.standard_normal(N)
        bias_d = torch.tensor(b, dtype=torch.float32, device=dev)
        ref = ref + b.astype(np.float32)
    if act:
        ref = np.maximum(ref, alpha * ref)
    if aux:
        h = rng.standard_normal((M, N)).astype(np.float32)
        aux_d = torch.tensor(_pad(h, ldd), device=dev)
        ref = ref * np.where(h > 0, 1.0, alpha)
    if row_scale:
        s = rng.standard_normal(M).astype(np.float32)
        rs_d = torch.tensor(s, device=dev)
        ref = ref * s[:, None]
    if sumsq:
        sq_d = torch.zeros((128 * M,), device=dev)
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p()
    tiles = C.c_int(0)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = tr.lib.wgan_gemm(tr._h, a_mn, b_mn, p(Ad), Ad.shape[1], p(Bd), Bd.shape[1], p(Dd), ldd, M, N, K,
                          p(bias_d), act, p(aux_d), ldd, p(rs_d), p(sq_d), C.byref(tiles), splits, st)
    assert rc == 0, tr.lib.wgan_last_error(tr._h)
    torch.cuda.synchronize()
    out = Dd.cpu().numpy()[:, :N].astype(np.float64)
    scale = np.abs(ref).max() + 1e-30
    err = np.abs(out - ref).max() / scale
    res = {"err": err, "out": Dd.cpu().numpy()[:, :N].copy()}
    if sumsq:
        got = sq_d.cpu().numpy()[:tiles.value * M].reshape(tiles.value, M).sum(axis=0)
        want = (ref * ref).sum(axis=1)
        res["sq_err"] = np.abs(got - want).max() / (np.abs(want).max() + 1e-30)
    return res


SHAPES = [
    # layout, M, N, K
    ("fwd", 128, 32, 32), ("fwd", 128, 224, 96), ("fwd", 100, 200, 133), ("fwd", 300, 600, 100),
    ("fwd", 257, 6605 // 8, 600), ("fwd", 5, 3, 7),
    ("dgrad", 128, 32, 32), ("dgrad", 130, 200, 200), ("dgrad", 96, 700, 200), ("dgrad", 77, 100, 600),
    ("wgrad", 128, 32, 32), ("wgrad", 200, 200, 300), ("wgrad", 661, 200, 512), ("wgrad", 100, 600, 96),
]


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("layout,M,N,K", SHAPES)
def test_gemm_plain(mode, layout, M, N, K):
    tr = _trainer(mode)
    rng = np.random.default_rng(hash((layout, M, N, K)) % (2 ** 32))
    r = _run(tr, layout, M, N, K, rng)
    assert r["err"] < TOL[mode], r["err"]


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("layout,M,N,K,splits", [("fwd", 256, 200, 2048, 4), ("wgrad", 300, 200, 4096, 7),
                                                   ("dgrad", 64, 96, 1000, 3), ("fwd", 128, 200, 6605, 0)])
def test_gemm_splitk(mode, layout, M, N, K, splits):
    tr = _trainer(mode)
    rng = np.random.default_rng(7)
    r = _run(tr, layout, M, N, K, rng, bias=True, act=1, splits=splits)
    assert r["err"] < TOL[mode], r["err"]


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("splits", [0, 3])
def test_gemm_epilogues(mode, splits):
    tr = _trainer(mode)
    rng = np.random.default_rng(11)
    r = _run(tr, "fwd", 200, 200, 300, rng, bias=True, act=1, splits=splits)
    assert r["err"] < TOL[mode], r["err"]
    r = _run(tr, "dgrad", 150, 200, 200, rng, aux=True, splits=splits)
    assert r["err"] < TOL[mode], r["err"]
    r = _run(tr, "fwd", 150, 200, 700, rng, aux=True, row_scale=True, splits=splits)
    assert r["err"] < TOL[mode], r["err"]
    r = _run(tr, "dgrad", 150, 700, 200, rng, sumsq=True, splits=splits)
    assert r["err"] < TOL[mode] and r["sq_err"] < 2 * TOL[mode], r


def test_gemm_many_tiles_persistent():
    """More tiles than SMs: exercises the persistent loop, TMEM double buffering and phase flips."""
    tr = _trainer(0)
    rng = np.random.default_rng(3)
    r = _run(tr, "fwd", 2048, 3000, 256, rng, bias=True, act=1)
    assert r["err"] < TOL[0], r
    r = _run(tr, "wgrad", 1500, 2000, 512, rng)
    assert r["err"] < TOL[0], r
    r = _run(tr, "dgrad", 4096, 1024, 128, rng, aux=True)
    assert r["err"] < TOL[0], r


def test_gemm_random_ragged_shapes():
    """30 random ragged shapes per layout (tails in M, N and K, tiny and multi-tile sizes, random epilogues)."""
    tr = _trainer(0)
    rng = np.random.default_rng(2024)
    for i in range(30):
        layout = ("fwd", "dgrad", "wgrad")[i % 3]
        M = int(rng.integers(1, 700)); N = int(rng.integers(1, 900)); K = int(rng.integers(1, 1500))
        kw = dict(bias=bool(rng.integers(2)), act=int(rng.integers(2)), aux=bool(rng.integers(2)),
                  row_scale=bool(rng.integers(2)), sumsq=bool(rng.integers(2)), splits=int(rng.integers(0, 4)))
        r = _run(tr, layout, M, N, K, rng, **kw)
        assert r["err"] < 3e-3, (layout, M, N, K, kw, r)
        if kw["sumsq"]:
            assert r["sq_err"] < 6e-3, (layout, M, N, K, kw, r)


# Stream-K (splits = -1 forces it): every CTA pair streams one contiguous range of (tile, K block) units; partial
# accumulators meet in a workspace and the tile's owner applies the fused epilogue.
STREAMK_SHAPES = [
    # few tiles, long K: many contributors per tile (1 tile: all 74 pairs feed one owner)
    ("fwd", 256, 200, 6605), ("fwd", 300, 200, 4000), ("wgrad", 661, 200, 4096), ("dgrad", 512, 96, 3000),
    # the two layer-1 products of the critic update at batch 1024 (policy picks stream-K there, see below)
    ("fwd", 3072, 200, 6605), ("wgrad", 6605, 200, 3072),
    # 512-row pair tiles (two sub-tiles per CTA: M >= 1024, K >= 2048, one N tile) with ragged M / N / K
    ("fwd", 1100, 200, 2100), ("wgrad", 1500, 130, 2500), ("dgrad", 1300, 96, 3000), ("fwd", 1025, 7, 2049),
    # ragged everything, multiple N tiles, units not divisible by the pair count
    ("fwd", 1000, 700, 1111), ("dgrad", 777, 300, 2500), ("wgrad", 900, 520, 1300),
]


@pytest.mark.parametrize("layout,M,N,K", STREAMK_SHAPES)
def test_gemm_streamk_forced(layout, M, N, K):
    tr = _trainer(0)
    rng = np.random.default_rng(hash((layout, M, N, K, 5)) % (2 ** 32))
    r = _run(tr, layout, M, N, K, rng, splits=-1)
    assert r["err"] < TOL[0], r
    r = _run(tr, layout, M, N, K, rng, bias=True, act=1, splits=-1)
    assert r["err"] < TOL[0], r
    if layout != "wgrad":
        r = _run(tr, layout, M, N, K, rng, aux=True, row_scale=True, sumsq=True, splits=-1)
        assert r["err"] < TOL[0] and r["sq_err"] < 2 * TOL[0], r


def test_gemm_streamk_policy_and_determinism():
    """The shapes the critic update runs at the benchmark batch: split-K 0 lets the policy choose (stream-K for both),
    the result matches the forced split-K + finish path to TF32 summation-order noise, and repeated launches are
    bit-identical (partials are added in pair order; flags are re-armed by the owner)."""
    tr = _trainer(0)
    for layout, M, N, K in [("fwd", 12288, 200, 6605), ("wgrad", 6605, 200, 12288)]:
        outs = []
        for splits in (0, 0, 0, 3):
            rng = np.random.default_rng(99)
            l0 = tr.launch_count
            r = _run(tr, layout, M, N, K, rng, bias=(layout == "fwd"), act=int(layout == "fwd"), splits=splits)
            outs.append((r, tr.launch_count - l0))
            assert r["err"] < TOL[0], (layout, splits, r)
        assert outs[0][1] == 1, "stream-K is one launch (no finish pass)"
        assert outs[3][1] >= 2
        assert np.array_equal(outs[0][0]["out"], outs[1][0]["out"]) and np.array_equal(outs[0][0]["out"], outs[2][0]["out"])

"""CPU tests of the host utilities next to the hot path: the per-gene scaler is pinned
against the reference's own third-party dependency (sklearn.preprocessing.MinMaxScaler, [W:4,34-39]),
the CSV reader against the R script's output contract ([R:86]: header row, no row names)."""
import os

import numpy as np
import pytest

from arshamg_scrnaseq_wgan_b200.data import GeneMinMaxScaler, load_ltpm_csv, prepare_training_matrix


def test_scaler_matches_sklearn_minmax():
    sk = pytest.importorskip("sklearn.preprocessing")
    rng = np.random.default_rng(0)
    m = rng.gamma(2.0, 2.0, size=(50, 37)) * (rng.random((50, 37)) > 0.4)
    m[7] = 3.25                                         # a constant gene
    ref = sk.MinMaxScaler().fit_transform(m.T).T        # exactly [W:34-39]
    got = GeneMinMaxScaler().fit_transform(m)
    assert np.allclose(got, ref, rtol=0, atol=1e-15)
    assert got[7].max() == 0.0
    s = GeneMinMaxScaler().fit(m)
    assert np.allclose(s.inverse_transform_cells(s.transform(m).T).T, m)


def test_csv_loader_and_training_matrix(tmp_path):
    rng = np.random.default_rng(1)
    m = np.round(rng.gamma(2.0, 2.0, size=(13, 40)), 6)
    p = tmp_path / "m.csv"
    with open(p, "w") as f:
        f.write(",".join(f"cell{i}" for i in range(40)) + "\n")
        for row in m:
            f.write(",".join(repr(float(v)) for v in row) + "\n")
    got = load_ltpm_csv(str(p))
    assert got.shape == (13, 40) and np.allclose(got, m)
    train, scaler, dmax = prepare_training_matrix(got, num_cells_train=25)
    assert train.shape == (25, 13) and train.dtype == np.float32
    assert 0.0 <= train.min() and dmax <= 1.0


def test_lfs_pointer_is_reported(tmp_path):
    p = tmp_path / "ptr.csv"
    p.write_text("version https://git-lfs.github.com/spec/v1\noid sha256:52950c14\nsize 132846935\n")
    with pytest.raises(FileNotFoundError):
        load_ltpm_csv(str(p))


class _FakeTrainer:
    """Duck-typed stand-in for WGANTrainer's state I/O (checkpoint.py only needs these five members)."""

    def __init__(self, rank=0, seed=0):
        from arshamg_scrnaseq_wgan_b200.trainer import PARAM_NAMES
        rng = np.random.default_rng(seed)
        self.rank, self.world = rank, 1
        self.t = {(k, kind): rng.standard_normal((3, 4)).astype(np.float32) for k in PARAM_NAMES
                  for kind in ("param", "m", "v")}
        self.pw = {0: (0.5, 0.9), 1: (0.25, 0.81)}

    def get_tensor(self, name, kind="param"):
        return self.t[(name, kind)]

    def set_tensor(self, name, value, kind="param"):
        self.t[(name, kind)] = np.asarray(value, np.float32)

    def adam_powers(self, net):
        return self.pw[net]

    def set_adam_powers(self, net, b1, b2):
        self.pw[net] = (b1, b2)


def test_checkpoint_paths_are_normalised_and_writes_are_atomic(tmp_path):
    """'.npz' is appended for save AND load, a save never leaves a torn file behind (written next to the
    destination, then renamed), `step_suffix` appends -<global_step> like the reference's Saver [W:203], and only
    rank 0 writes."""
    import os
    from arshamg_scrnaseq_wgan_b200.checkpoint import save_checkpoint, load_checkpoint
    from arshamg_scrnaseq_wgan_b200.trainer import PARAM_NAMES
    a = _FakeTrainer(seed=1)
    base = str(tmp_path / "run")                       # no suffix given
    dest = save_checkpoint(a, base, global_step=7)
    assert dest == base + ".npz" and os.path.exists(dest)
    assert not [f for f in os.listdir(tmp_path) if ".tmp" in f]          # the temporary file was renamed away
    b = _FakeTrainer(seed=2)
    assert load_checkpoint(b, base) == 7                # same path string as the save
    for k in PARAM_NAMES:
        for kind in ("param", "m", "v"):
            assert np.array_equal(a.t[(k, kind)], b.t[(k, kind)])
    assert b.pw[0] == pytest.approx(a.pw[0]) and b.pw[1] == pytest.approx(a.pw[1])
    # TF slot names in the file
    ck = np.load(dest)
    assert PARAM_NAMES[0] + "/Adam" in ck and PARAM_NAMES[0] + "/Adam_1" in ck and "generator_opt/beta1_power" in ck
    # step-suffixed intermediate saves do not overwrite each other
    d1 = save_checkpoint(a, base, global_step=1000, step_suffix=True)
    d2 = save_checkpoint(a, base, global_step=2000, step_suffix=True)
    assert d1.endswith("run-1000.npz") and d2.endswith("run-2000.npz") and os.path.exists(d1) and os.path.exists(d2)
    # ranks other than 0 do not write
    other = str(tmp_path / "other")
    save_checkpoint(_FakeTrainer(rank=1), other, global_step=1)
    assert not os.path.exists(other + ".npz")

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

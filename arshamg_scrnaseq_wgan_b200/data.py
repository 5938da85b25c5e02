"""Host data plane of the reference script [W:32-54].

The reference reads a genes x cells log2(TPM+1) CSV with a header row and no row names
([R:86], `np.genfromtxt(..., skip_header=1)` [W:32]), transposes it, fits sklearn's
MinMaxScaler (per-gene min-max over all cells [W:34-39]), transposes back and keeps the first
`num_cells_train` cells [W:53-54].  It never inverts the scaling of generated cells; the
inverse transform here does.  One-off host work, deliberately numpy (not a kernel).
"""
from __future__ import annotations

import numpy as np


def load_ltpm_csv(path: str, delimiter: str = ",") -> np.ndarray:
    """genes x cells float64 matrix; first row is the header [W:32]."""
    with open(path, "rb") as f:
        head = f.read(200)
    if head.startswith(b"version https://git-lfs.github.com"):
        raise FileNotFoundError(f"{path} is a Git-LFS pointer, not the expression matrix (fetch the LFS object first)")
    m = np.loadtxt(path, delimiter=delimiter, skiprows=1, dtype=np.float64, ndmin=2)
    return m


class GeneMinMaxScaler:
    """Per-gene min-max to [0, 1] over all cells == MinMaxScaler().fit_transform(matrix.T).T [W:34-39];
    constant genes map to 0 (sklearn replaces a zero range by 1)."""

    def fit(self, genes_by_cells: np.ndarray) -> "GeneMinMaxScaler":
        m = np.asarray(genes_by_cells, dtype=np.float64)
        self.data_min_ = m.min(axis=1)
        rng = m.max(axis=1) - self.data_min_
        rng[rng == 0.0] = 1.0
        self.data_range_ = rng
        return self

    def transform(self, genes_by_cells: np.ndarray) -> np.ndarray:
        m = np.asarray(genes_by_cells, dtype=np.float64)
        return (m - self.data_min_[:, None]) / self.data_range_[:, None]

    def fit_transform(self, genes_by_cells: np.ndarray) -> np.ndarray:
        return self.fit(genes_by_cells).transform(genes_by_cells)

    def inverse_transform_cells(self, cells_by_genes: np.ndarray) -> np.ndarray:
        """Generated cells [n, genes] in [0, 1] units back to log2(TPM+1)."""
        c = np.asarray(cells_by_genes, dtype=np.float64)
        return c * self.data_range_[None, :] + self.data_min_[None, :]


def prepare_training_matrix(genes_by_cells: np.ndarray, num_cells_train: int = 1500):
    """[W:34-54]: scale, then the first num_cells_train cells, returned CELLS-major float32 (the layout the
    device keeps resident) together with the scaler and data_max_value [W:89]."""
    scaler = GeneMinMaxScaler()
    scaled = scaler.fit_transform(genes_by_cells)
    train = scaled[:, :num_cells_train]
    return np.ascontiguousarray(train.T, dtype=np.float32), scaler, float(np.amax(train))

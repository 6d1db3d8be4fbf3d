"""Input formats of the reference's documented workflows (docs/src/man/inputdata.md:5-10, 45-91), host side.

  read_matrix_networks(path)      vectorised form: n rows x (V(V+1)/2 + 1) columns, header row, response last
                                  (examples/matrix_networks.csv)                      -> Fit(X, y, R, x_transform=False)
  read_adjacency_csvs(dir, n)     data1.csv .. data<n>.csv, one V x V adjacency matrix each, header row
                                  (examples/data*.csv)                                -> Fit(X, y, R, x_transform=True)
  read_responses(path)            one-column CSV with a header (examples/responses.csv)

JLD2 containers (examples/vector_networks.jld2) are a Julia-side format and stay with the Julia wrapper."""
import os

import numpy as np


def _loadcsv(path):
    return np.loadtxt(path, delimiter=",", skiprows=1, ndmin=2, dtype=np.float64)


def read_matrix_networks(path, response_last=True):
    """-> (X, y): X is n x q with q = V(V+1)/2 columns in lower_triangle order (src/utils.jl:40-57)."""
    a = _loadcsv(path)
    if not response_last:
        return a, None
    X, y = a[:, :-1], a[:, -1]
    q = X.shape[1]
    V = int(round((-1 + np.sqrt(1 + 8 * q)) / 2))
    if V * (V + 1) // 2 != q:
        raise ValueError("%s: %d predictor columns is not V(V+1)/2 for an integer V" % (path, q))
    return np.ascontiguousarray(X), np.ascontiguousarray(y)


def read_adjacency_csvs(directory, n, pattern="data%d.csv"):
    """-> list of n square matrices (what Fit!(X, ...; x_transform=true) takes, docs/src/man/inputdata.md:52-66)."""
    mats = []
    for i in range(1, n + 1):
        m = _loadcsv(os.path.join(directory, pattern % i))
        if m.shape[0] != m.shape[1]:
            raise ValueError("%s is not square: %s" % (pattern % i, m.shape))
        if mats and m.shape != mats[0].shape:
            raise ValueError("all adjacency matrices must have the same size")
        mats.append(m)
    return mats


def read_responses(path):
    return _loadcsv(path)[:, 0].copy()

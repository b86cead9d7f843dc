"""Shared helpers for the test-suite."""
import numpy as np


def csc_dense(g, key):
    """Densify a SparseMatrixCSC stored by tests/golden/extract_o3_jld.py."""
    m, n = int(g[key + "__m"]), int(g[key + "__n"])
    colptr, rowval, nzval = g[key + "__colptr"], g[key + "__rowval"], g[key + "__nzval"]
    a = np.zeros((m, n), dtype=nzval.dtype)
    for j in range(n):
        for k in range(colptr[j] - 1, colptr[j + 1] - 1):
            a[rowval[k] - 1, j] = nzval[k]
    return a


def maxabs(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))))

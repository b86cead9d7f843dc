"""Extract the reference's JLD fixtures into plain .npz files (run in the build container).

    python tests/golden/extract_o3_jld.py [/root/reference]

Writes tests/golden/O3.npz and tests/golden/linalg.npz.  Dense arrays keep Julia's index order;
a SparseMatrixCSC ``X`` becomes ``X__m, X__n, X__colptr, X__rowval, X__nzval`` (1-based as stored).
/root/reference does not exist on the GPU box, so tests read only the .npz files.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from jld_scan import JLDFile  # noqa: E402

SKIP = {"ENDIAN_BOM", "JULIA_MAJOR", "JULIA_MINOR", "JULIA_PATCH", "WORD_SIZE"}


def extract(src, dst):
    f = JLDFile(src)
    out = {}
    for k in f.names():
        if k in SKIP:
            continue
        v = f.load(k)
        if isinstance(v, dict):
            for kk, vv in v.items():
                out[f"{k}__{kk}"] = np.asarray(vv)
        else:
            out[k] = v
    np.savez_compressed(dst, **out)
    return sorted(out)


if __name__ == "__main__":
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    here = os.path.dirname(os.path.abspath(__file__))
    for name in ("O3", "linalg"):
        keys = extract(os.path.join(ref, "test", "data", name + ".jld"), os.path.join(here, name + ".npz"))
        print(name, len(keys), "arrays")

"""Shared test helpers: copy a product hierarchy into the oracle's plain container."""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import oracle as orc  # noqa: E402
import omp_amg_b200 as amg  # noqa: E402
from omp_amg_b200 import api  # noqa: E402

EMU_SO = os.path.join(ROOT, "tests", "_emu", "libamgb_emu.so")


def fetch(H):
    """omp_amg_b200.Hierarchy -> oracle.Hierarchy (numpy copies on the host)."""
    out = orc.Hierarchy()
    out.nlevels = H.nlevels
    out.nullspace = H.nullspace
    for l in range(out.nlevels):
        info = H.level_info(l)
        par = H.level_params(l)
        lev = {"info": [info[k] for k in ("n", "nnz", "nf", "nc", "nnzf", "nnzw", "nnzfp", "coarsen_rounds",
                                            "lanczos_iters", "interp_rounds")],
               "m": par["m"], "rho": par["rho"], "lmin": par["lambda_min"], "lmax": par["lambda_max"]}
        names = (("A", api.A),) if l == out.nlevels - 1 else (("A", api.A), ("Af", api.AF), ("W", api.W), ("AfP", api.AFP))
        for name, which in names:
            ro, col, a, shape = H.csr(l, which)
            lev[name] = (ro.astype(np.int64), col.astype(np.int64), a.copy(), shape)
        if l < out.nlevels - 1:
            lev["C"] = H.vec(l, api.VEC_C).copy()
            lev["D"] = H.vec(l, api.VEC_D).copy()
            lev["idc"] = H.vec(l, api.VEC_IDC).copy()
            lev["idf"] = H.vec(l, api.VEC_IDF).copy()
        out.levels.append(lev)
    return out


def product_trace(L):
    out = []
    tag = C.create_string_buffer(96)
    hs, nb = C.c_uint64(), C.c_int64()
    for i in range(L.amgb_trace_count()):
        L.amgb_trace_get(i, tag, 96, C.byref(hs), C.byref(nb))
        out.append((tag.value.decode(), hs.value, nb.value))
    return out


def first_trace_mismatch(ta, tb):
    for i, (x, y) in enumerate(zip(ta, tb)):
        if x != y:
            return i, x, y
    if len(ta) != len(tb):
        return min(len(ta), len(tb)), (ta + [None])[min(len(ta), len(tb))], (tb + [None])[min(len(ta), len(tb))]
    return None

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


def load_trace_fixture(path):
    """tests/golden/trace_*.json.gz (make_trace_fixtures.py): trace as (tag, hash, bytes) tuples."""
    import gzip
    import json
    with gzip.open(path, "rb") as f:
        fx = json.loads(f.read().decode())
    fx["trace"] = [(t, int(h, 16), nb) for t, h, nb in fx["trace"]]
    return fx


def first_trace_mismatch(ta, tb):
    for i, (x, y) in enumerate(zip(ta, tb)):
        if x != y:
            return i, x, y
    if len(ta) != len(tb):
        return min(len(ta), len(tb)), (ta + [None])[min(len(ta), len(tb))], (tb + [None])[min(len(ta), len(tb))]
    return None


# ---------------------------------------------------------------------------------------
# mxm (amg_setup.c:1894) in isolation: operands that reach every SpGEMM bin, exact reference
# ---------------------------------------------------------------------------------------
def spgemm_reference(A, B):
    """X = A*B exactly as mxm forms it: for every row, the rows of B are visited in the order of the
    entries of A (columns ascending), each product is rounded, then added to the running value of
    its column; entries that end exactly 0.0 are dropped; columns ascending."""
    aro, acol, aa, (arn, _) = A
    bro, bcol, ba, (_, bcn) = B
    xro = np.zeros(arn + 1, np.int64)
    cols, vals = [], []
    for i in range(arn):
        acc = {}
        for ja in range(aro[i], aro[i + 1]):
            k, av = int(acol[ja]), float(aa[ja])
            for jb in range(bro[k], bro[k + 1]):
                c = int(bcol[jb])
                acc[c] = acc.get(c, 0.0) + float(ba[jb]) * av
        row = sorted((c, v) for c, v in acc.items() if v != 0.0)
        cols.extend(c for c, _ in row)
        vals.extend(v for _, v in row)
        xro[i + 1] = len(cols)
    return xro, np.array(cols, np.int64), np.array(vals, np.float64), (arn, bcn)


def spgemm_adversarial_operands(seed=0):
    """A (rows of every size class, some empty) and B (short/medium/long rows over narrow to very
    wide column ranges) such that the rows of A*B fall into every tier of spgemm.cu: 8-thread tiles,
    32-thread tiles, warps with 512 and 2048 slots (fitting and handed down), blocks with 4096 and
    8192 slots in shared memory, blocks with a table in HBM and the global-hash kernel.  Values mix
    small signed powers of two (exact cancellations, so entries vanish) with random doubles (so the
    order of the additions shows in the last bit)."""
    rng = np.random.default_rng(seed)
    bcn = 1_200_000
    classes = [(1000, (0, 7), 20_000), (1000, (20, 61), 24_000), (400, (20, 61), 380_000),
               (300, (20, 61), 790_000), (300, (200, 401), 790_000), (300, (200, 401), bcn),
               (600, (130, 161), 6_000), (600, (140, 201), 21_000), (500, (20, 61), 3_000)]      # rows that fill most of a narrow span (dense tiers)
    bro, bcol, ba, first = [0], [], [], []
    for count, (lo, hi), width in classes:
        first.append(len(bro) - 1)
        for _ in range(count):
            n = int(rng.integers(lo, hi))
            c = np.sort(rng.choice(width, size=n, replace=False)) if n else np.zeros(0, np.int64)
            v = np.where(rng.random(n) < 0.5, rng.choice([-2.0, -1.0, -0.5, 0.5, 1.0, 2.0], size=n), rng.standard_normal(n))
            bcol.extend(c.tolist()); ba.extend(v.tolist()); bro.append(len(bcol))
    brn = len(bro) - 1
    plan = [(0, (1, 5)), (0, (10, 16)), (1, (3, 5)), (1, (20, 61)), (2, (10, 15)), (2, (40, 51)),
            (4, (18, 23)), (4, (30, 36)), (5, (4, 9)), (3, (60, 90)), (1, (70, 91)), (1, (110, 141)), (1, (230, 281)), (6, (30, 41)), (7, (60, 81)), (8, (50, 81))]
    aro, acol, aa = [0], [], []
    for cls, (lo, hi) in plan:
        count = classes[cls][0]
        for rep in range(12):
            n = int(rng.integers(lo, hi))
            ks = np.sort(first[cls] + rng.choice(count, size=n, replace=False))
            v = np.where(rng.random(n) < 0.5, rng.choice([-2.0, -1.0, 1.0, 2.0], size=n), rng.standard_normal(n))
            acol.extend(ks.tolist()); aa.extend(v.tolist()); aro.append(len(acol))
            if rep % 5 == 4:
                aro.append(len(acol))                       # an empty row of A in between
    # a row that multiplies only empty rows of B, and exact duplicates of a row with opposite sign
    empty_b = [k for k in range(brn) if bro[k + 1] == bro[k]][:3]
    if empty_b:
        acol.extend(empty_b); aa.extend([1.0] * len(empty_b)); aro.append(len(acol))
    arn = len(aro) - 1
    A = (np.array(aro, np.int32), np.array(acol, np.int32), np.array(aa, np.float64), (arn, brn))
    B = (np.array(bro, np.int32), np.array(bcol, np.int32), np.array(ba, np.float64), (brn, bcn))
    return A, B


def same_csr_bits(X, Y):
    return (np.array_equal(np.asarray(X[0], np.int64), np.asarray(Y[0], np.int64))
            and np.array_equal(np.asarray(X[1], np.int64), np.asarray(Y[1], np.int64))
            and np.array_equal(np.asarray(X[2], np.float64).view(np.uint64), np.asarray(Y[2], np.float64).view(np.uint64)))

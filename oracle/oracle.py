"""TEST INFRASTRUCTURE ONLY: ctypes access to the CPU checkers.

* ``Oracle``   -- oracle/libamg_oracle.so, the restatement (amg_oracle.c)
* ``Ref``      -- oracle/_ref/libamg_ref.so, the unmodified reference compiled by
                  oracle/Makefile from /root/reference (amg_setup.c, amg_tools.c, ...)
* ``RefVcycle`` -- oracle/_ref/libvcycle_ref.so: the reference's own amg_exec / crs_solve
                  (amg.c:85-189, cut out at build time and compiled unchanged inside
                  oracle/vcycle_ref_harness.c) on a hierarchy in this module's container

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libamg_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libamg_ref.so")
REF_BIN = os.path.join(HERE, "_ref", "serial_amg")
VREF_SO = os.path.join(HERE, "_ref", "libvcycle_ref.so")

SEQ, TREE = 0, 1
CSR_A, CSR_AF, CSR_W, CSR_AFP = 0, 1, 2, 3
VEC_C, VEC_D, VEC_IDC, VEC_IDF = 0, 1, 2, 3

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build(ref=True):
    """Compile the checkers (oracle/Makefile).  Building the checker is not using it."""
    subprocess.run(["make", "-C", HERE, "libamg_oracle.so"], check=True, stdout=subprocess.DEVNULL)
    if ref:
        subprocess.run(["make", "-C", HERE, "ref"], check=True, stdout=subprocess.DEVNULL)


class _OCSR(C.Structure):
    _fields_ = [("rn", C.c_int), ("cn", C.c_int), ("ro", C.POINTER(C.c_int)),
                ("col", C.POINTER(C.c_int)), ("a", C.POINTER(C.c_double))]


class Hierarchy:
    """Plain-numpy copy of a hierarchy; the same container is filled from the oracle, the
    reference and the CUDA product so that tests compare like with like."""

    def __init__(self):
        self.nlevels = 0
        self.nullspace = 0
        self.levels = []      # dicts: A, Af, W, AfP (ro,col,a,shape), C, D, idc, idf, m, rho, info

    def __repr__(self):
        return "Hierarchy(nlevels=%d, n=%s)" % (self.nlevels, [l["A"][3][0] for l in self.levels])


def _csr_tuple(ro, col, a, rn, cn):
    return (np.asarray(ro, np.int64), np.asarray(col, np.int64), np.asarray(a, np.float64), (rn, cn))


class Oracle:
    def __init__(self, path=ORACLE_SO):
        if not os.path.exists(path):
            build(ref=False)
        L = self.L = C.CDLL(path)
        L.amgo_setup.argtypes = [C.c_int64, _i32p, _i32p, _f64p, C.c_int, C.POINTER(C.c_void_p)]
        L.amgo_setup.restype = C.c_int
        L.amgo_free.argtypes = [C.c_void_p]
        L.amgo_nlevels.argtypes = [C.c_void_p]
        L.amgo_nullspace.argtypes = [C.c_void_p]
        L.amgo_level_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64)]
        L.amgo_level_params.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        L.amgo_get_csr.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                   C.POINTER(C.c_int64), C.c_void_p, C.c_void_p, C.c_void_p]
        L.amgo_get_vec.argtypes = [C.c_void_p, C.c_int, C.c_int, _f64p]
        L.amgo_export.argtypes = [C.c_void_p, C.c_char_p]
        L.amgo_solve.argtypes = [C.c_void_p, _f64p, _f64p]
        L.amgo_dot.argtypes = [_f64p, _f64p, C.c_int64, C.c_int]
        L.amgo_dot.restype = C.c_double
        L.amgo_trace_enable.argtypes = [C.c_int]
        L.amgo_trace_get.argtypes = [C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_int64)]
        for f in ("amgo_build_csr", "amgo_transpose", "amgo_spgemm", "amgo_mpm", "amgo_mxmpoint",
                  "amgo_sub_mat", "amgo_interpolation", "amgo_csr_new"):
            getattr(L, f).restype = C.POINTER(_OCSR)
        L.amgo_build_csr.argtypes = [C.c_int64, _i32p, _i32p, _f64p]
        L.amgo_transpose.argtypes = [C.POINTER(_OCSR)]
        L.amgo_spgemm.argtypes = [C.POINTER(_OCSR), C.POINTER(_OCSR)]
        L.amgo_mpm.argtypes = [C.c_double, C.POINTER(_OCSR), C.c_double, C.POINTER(_OCSR)]
        L.amgo_mxmpoint.argtypes = [C.POINTER(_OCSR), C.POINTER(_OCSR)]
        L.amgo_sub_mat.argtypes = [C.POINTER(_OCSR), _f64p, _f64p]
        L.amgo_coarsen.argtypes = [_f64p, C.POINTER(_OCSR), C.c_double]
        L.amgo_interpolation.argtypes = [C.POINTER(_OCSR)] * 3 + [C.c_double, C.c_double, C.c_int,
                                                                  C.POINTER(C.c_int)]
        L.amgo_csr_free.argtypes = [C.POINTER(_OCSR)]
        L.amgo_csr_new.argtypes = [C.c_int, C.c_int, C.c_int64]
        L.amgo_rng_seed.argtypes = [C.c_void_p, C.c_uint32]
        L.amgo_rng_next.argtypes = [C.c_void_p]
        L.amgo_rng_next.restype = C.c_int32

    # ---- csr marshalling for the single-stage entry points ----
    def to_csr(self, ro, col, a, shape):
        ro = np.asarray(ro, np.int32); col = np.asarray(col, np.int32); a = np.asarray(a, np.float64)
        p = self.L.amgo_csr_new(shape[0], shape[1], len(col))
        C.memmove(p.contents.ro, ro.ctypes.data, ro.nbytes)
        if len(col):
            C.memmove(p.contents.col, col.ctypes.data, col.nbytes)
            C.memmove(p.contents.a, a.ctypes.data, a.nbytes)
        return p

    def from_csr(self, p, free=True):
        m = p.contents
        ro = np.ctypeslib.as_array(m.ro, (m.rn + 1,)).copy()
        nnz = int(ro[-1])
        col = np.ctypeslib.as_array(m.col, (max(nnz, 1),))[:nnz].copy()
        a = np.ctypeslib.as_array(m.a, (max(nnz, 1),))[:nnz].copy()
        out = _csr_tuple(ro, col, a, m.rn, m.cn)
        if free:
            self.L.amgo_csr_free(p)
        return out

    def rand_stream(self, n, seed=1):
        buf = C.create_string_buffer(256)
        self.L.amgo_rng_seed(buf, seed)
        return np.array([self.L.amgo_rng_next(buf) for _ in range(n)], dtype=np.int64)

    def dot(self, a, b, mode):
        a = np.ascontiguousarray(a, np.float64); b = np.ascontiguousarray(b, np.float64)
        return self.L.amgo_dot(a, b, len(a), mode)

    # ---- full setup ----
    def setup_raw(self, Ai, Aj, Av, mode=SEQ, trace=False):
        Ai = np.ascontiguousarray(Ai, np.int32); Aj = np.ascontiguousarray(Aj, np.int32)
        Av = np.ascontiguousarray(Av, np.float64)
        h = C.c_void_p()
        self.L.amgo_trace_enable(1 if trace else 0)
        rc = self.L.amgo_setup(len(Av), Ai, Aj, Av, mode, C.byref(h))
        if rc != 0:
            raise RuntimeError("amgo_setup failed: %d" % rc)
        return h

    def trace(self):
        out = []
        tag = C.create_string_buffer(80)
        hs, nb = C.c_uint64(), C.c_int64()
        for i in range(self.L.amgo_trace_count()):
            self.L.amgo_trace_get(i, tag, 80, C.byref(hs), C.byref(nb))
            out.append((tag.value.decode(), hs.value, nb.value))
        return out

    def fetch(self, h):
        H = Hierarchy()
        H.nlevels = self.L.amgo_nlevels(h)
        H.nullspace = self.L.amgo_nullspace(h)
        for l in range(H.nlevels):
            info = (C.c_int64 * 10)()
            par = (C.c_double * 4)()
            self.L.amgo_level_info(h, l, info)
            self.L.amgo_level_params(h, l, par)
            lev = {"info": list(info), "m": par[0], "rho": par[1], "lmin": par[2], "lmax": par[3]}
            for name, which in (("A", CSR_A), ("Af", CSR_AF), ("W", CSR_W), ("AfP", CSR_AFP)):
                if l == H.nlevels - 1 and which != CSR_A:
                    continue
                rn, cn, nnz = C.c_int(), C.c_int(), C.c_int64()
                self.L.amgo_get_csr(h, l, which, C.byref(rn), C.byref(cn), C.byref(nnz), None, None, None)
                ro = np.zeros(rn.value + 1, np.int32); col = np.zeros(max(nnz.value, 1), np.int32)
                a = np.zeros(max(nnz.value, 1), np.float64)
                self.L.amgo_get_csr(h, l, which, None, None, None, ro.ctypes.data, col.ctypes.data, a.ctypes.data)
                lev[name] = _csr_tuple(ro, col[:nnz.value], a[:nnz.value], rn.value, cn.value)
            if l < H.nlevels - 1:
                n, nf, nc = info[0], info[2], info[3]
                for name, which, ln in (("C", VEC_C, n), ("D", VEC_D, nf), ("idc", VEC_IDC, nc), ("idf", VEC_IDF, nf)):
                    v = np.zeros(max(ln, 1), np.float64)
                    self.L.amgo_get_vec(h, l, which, v)
                    lev[name] = v[:ln].copy()
            H.levels.append(lev)
        return H

    def setup(self, Ai, Aj, Av, mode=SEQ):
        h = self.setup_raw(Ai, Aj, Av, mode)
        try:
            return self.fetch(h)
        finally:
            self.L.amgo_free(h)

    def export(self, h, dirname):
        rc = self.L.amgo_export(h, dirname.encode())
        if rc:
            raise RuntimeError("amgo_export failed: %d" % rc)

    def solve(self, h, b):
        b = np.ascontiguousarray(b, np.float64)
        x = np.zeros_like(b)
        self.L.amgo_solve(h, x, b)
        return x

    def free(self, h):
        self.L.amgo_free(h)


# ------------------------------------------------------------------------------------------
# the reference's V-cycle (amg.c:114-189) on a Hierarchy
# ------------------------------------------------------------------------------------------
def level_sorted_layout(H):
    """The storage amg_exec works on (amg.c:117, :438-446): the unknowns sorted by the level at
    which they become F (the last level's single unknown at the end), ascending inside a level.
    Returns (off, g): off[l] = first position of level l (len nlevels+1), g[l][i] = position of
    unknown i of level l."""
    nl = H.nlevels
    nf = [int(H.levels[l]["W"][3][0]) for l in range(nl - 1)]
    nlast = int(H.levels[nl - 1]["A"][3][0])
    off = np.concatenate([[0], np.cumsum(nf + [nlast])]).astype(np.int64)
    g = [None] * nl
    g[nl - 1] = off[nl - 1] + np.arange(nlast, dtype=np.int64)
    for l in range(nl - 2, -1, -1):
        C_ = np.asarray(H.levels[l]["C"]) != 0.0
        gl = np.empty(len(C_), np.int64)
        gl[~C_] = off[l] + np.arange(int((~C_).sum()), dtype=np.int64)
        gl[C_] = g[l + 1]
        g[l] = gl
    return off, g


class RefVcycle:
    """crs_solve of the reference (one process, every id unique) on a Hierarchy.  The matrices keep
    the storage order of their rows (level-local ascending columns); only the column numbers are
    mapped into the level-sorted layout, as amg_setup_mats does (amg.c:295-377)."""

    def __init__(self, path=VREF_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle ref` where /root/reference exists)")
        self.L = C.CDLL(path)
        self.L.vref_solve.restype = C.c_int

    @staticmethod
    def available():
        return os.path.exists(VREF_SO)

    def solve(self, H, b, null_space=None):
        nl = H.nlevels
        off, g = level_sorted_layout(H)
        tot = int(off[nl])
        n0 = int(H.levels[0]["A"][3][0])
        assert tot == n0, (tot, n0)
        Dff = np.zeros(tot + 1, np.float64)
        m = np.ones(max(nl, 1), np.uint32)
        rho = np.zeros(max(nl, 1), np.float64)
        keep, ros, cols, vals = [], [], [], []
        for l in range(nl - 1):
            lev = H.levels[l]
            Dff[off[l]:off[l + 1]] = lev["D"]
            m[l] = int(lev["m"]); rho[l] = lev["rho"]
            tail = g[l + 1] - off[l + 1]            # level l+1 numbering -> position behind off[l+1]
            for name in ("W", "AfP", "Af"):
                ro, col, a, _ = lev[name]
                c2 = col if name == "Af" else tail[col]
                arrs = (np.ascontiguousarray(ro, np.uint64), np.ascontiguousarray(np.append(c2, 0), np.uint64),
                        np.ascontiguousarray(np.append(a, 0.0), np.float64))
                keep.append(arrs)
                ros.append(arrs[0].ctypes.data); cols.append(arrs[1].ctypes.data); vals.append(arrs[2].ctypes.data)
        last = H.levels[nl - 1]["A"]
        if off[nl] - off[nl - 1] == 1:              # dvec of the last level as amg_export writes it (amg_setup.c:166, :452)
            Dff[off[nl - 1]] = 0.0 if (H.nullspace or len(last[2]) == 0) else 1.0 / last[2][0]
        umap = np.empty(tot, np.uint64)
        umap[g[0]] = np.arange(n0, dtype=np.uint64)
        P = C.c_void_p * max(len(ros), 1)
        offu = np.ascontiguousarray(off, np.uint64)
        bb = np.ascontiguousarray(b, np.float64).copy()
        x = np.zeros(n0, np.float64)
        ns = H.nullspace if null_space is None else null_space
        rc = self.L.vref_solve(C.c_uint(nl), offu.ctypes.data_as(C.c_void_p), Dff.ctypes.data_as(C.c_void_p),
                               m.ctypes.data_as(C.c_void_p), rho.ctypes.data_as(C.c_void_p), P(*ros), P(*cols), P(*vals),
                               C.c_ulong(n0), umap.ctypes.data_as(C.c_void_p), C.c_int(int(ns)),
                               x.ctypes.data_as(C.c_void_p), bb.ctypes.data_as(C.c_void_p))
        assert rc == 0
        del keep
        return x


# ------------------------------------------------------------------------------------------
# the compiled reference (uint = unsigned long: the reference Makefile's -DUSE_LONG)
# ------------------------------------------------------------------------------------------
class _RCSR(C.Structure):
    _fields_ = [("rn", C.c_ulong), ("cn", C.c_ulong), ("row_off", C.POINTER(C.c_ulong)),
                ("col", C.POINTER(C.c_ulong)), ("a", C.POINTER(C.c_double))]


class _RData(C.Structure):     # struct amg_setup_data, amg_tools.h:29
    _fields_ = [("tolc", C.c_double), ("gamma", C.c_double), ("n", C.POINTER(C.c_double)),
                ("nnz", C.POINTER(C.c_double)), ("nnzf", C.POINTER(C.c_double)),
                ("nnzfp", C.POINTER(C.c_double)), ("m", C.POINTER(C.c_double)),
                ("rho", C.POINTER(C.c_double)), ("A", C.POINTER(C.POINTER(_RCSR))),
                ("id", C.POINTER(C.c_ulong)), ("idc", C.POINTER(C.POINTER(C.c_ulong))),
                ("idf", C.POINTER(C.POINTER(C.c_ulong))), ("C", C.POINTER(C.POINTER(C.c_double))),
                ("F", C.POINTER(C.POINTER(C.c_double))), ("D", C.POINTER(C.POINTER(C.c_double))),
                ("Af", C.POINTER(C.POINTER(_RCSR))), ("W", C.POINTER(C.POINTER(_RCSR))),
                ("AfP", C.POINTER(C.POINTER(_RCSR))), ("nlevels", C.c_ulong), ("nullspace", C.c_ulong)]


class _Quiet:
    """The reference prints progress to stdout; silence fd 1 around calls into it."""

    def __enter__(self):
        sys.stdout.flush()
        self.libc = C.CDLL(None)
        self.libc.fflush(None)
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)

    def __exit__(self, *a):
        self.libc.fflush(None)
        os.dup2(self.saved, 1)
        os.close(self.null)
        os.close(self.saved)


class Ref:
    def __init__(self, path=REF_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle ref` where /root/reference exists)")
        self.L = C.CDLL(path)
        self.libc = C.CDLL(None)
        L = self.L
        P = C.POINTER(_RCSR)
        L.amg_setup.argtypes = [C.c_ulong, C.c_void_p, C.c_void_p, _f64p, C.POINTER(_RData)]
        L.coarsen.argtypes = [_f64p, P, C.c_double]
        L.interpolation.argtypes = [P, P, P, P, C.c_double, C.c_double]
        L.mxm.argtypes = [P, P, P, C.c_double]
        L.mpm.argtypes = [P, C.c_double, P, C.c_double, P]
        L.mxmpoint.argtypes = [P, P, P]
        L.transpose.argtypes = [P, P]
        L.sub_mat.argtypes = [P, P, _f64p, _f64p]
        L.amg_export.argtypes = [C.POINTER(_RData)]

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def _keep(self, *objs):
        self._alive = getattr(self, "_alive", []) + list(objs)

    def to_csr(self, ro, col, a, shape):
        ro = np.ascontiguousarray(ro, np.uint64); col = np.ascontiguousarray(col, np.uint64)
        a = np.ascontiguousarray(a, np.float64)
        if len(col) == 0:
            col = np.zeros(1, np.uint64); a = np.zeros(1, np.float64)
        m = _RCSR(shape[0], shape[1], ro.ctypes.data_as(C.POINTER(C.c_ulong)),
                  col.ctypes.data_as(C.POINTER(C.c_ulong)), a.ctypes.data_as(C.POINTER(C.c_double)))
        self._keep(ro, col, a, m)
        return C.pointer(m)

    @staticmethod
    def from_csr(m):
        if isinstance(m, C.POINTER(_RCSR)):
            m = m.contents
        rn, cn = int(m.rn), int(m.cn)
        ro = np.ctypeslib.as_array(m.row_off, (rn + 1,)).astype(np.int64)
        nnz = int(ro[-1])
        col = np.ctypeslib.as_array(m.col, (max(nnz, 1),))[:nnz].astype(np.int64)
        a = np.ctypeslib.as_array(m.a, (max(nnz, 1),))[:nnz].copy()
        return _csr_tuple(ro, col, a, rn, cn)

    def new_csr(self):
        m = _RCSR()
        self._keep(m)
        return C.pointer(m)

    def setup(self, Ai, Aj, Av, seed=1):
        """amg_setup (amg_setup.c:60) on 0-based COO input; rand() reseeded as in a fresh
        run of serial_amg."""
        Ai = np.ascontiguousarray(Ai, np.uint64); Aj = np.ascontiguousarray(Aj, np.uint64)
        Av = np.ascontiguousarray(Av, np.float64)
        data = _RData()
        self.libc.srand(seed)
        with _Quiet():
            self.L.amg_setup(len(Av), Ai.ctypes.data, Aj.ctypes.data, Av, C.byref(data))
        H = Hierarchy()
        H.nlevels = int(data.nlevels)
        H.nullspace = int(data.nullspace)
        for l in range(H.nlevels):
            lev = {"A": self.from_csr(data.A[l])}
            if l < H.nlevels - 1:
                n = lev["A"][3][0]
                lev["Af"] = self.from_csr(data.Af[l]); lev["W"] = self.from_csr(data.W[l])
                lev["AfP"] = self.from_csr(data.AfP[l])
                nf, nc = lev["W"][3]
                lev["C"] = np.ctypeslib.as_array(data.C[l], (n,)).copy()
                lev["D"] = np.ctypeslib.as_array(data.D[l], (nf,)).copy()
                lev["idc"] = np.ctypeslib.as_array(data.idc[l], (max(nc, 1),))[:nc].astype(np.float64)
                lev["idf"] = np.ctypeslib.as_array(data.idf[l], (max(nf, 1),))[:nf].astype(np.float64)
                lev["m"] = data.m[l]; lev["rho"] = data.rho[l]
            H.levels.append(lev)
        self._last = data
        return H

    def export_last(self, dirname):
        """amg_export (amg_setup.c:405) of the last setup; the reference writes into the cwd."""
        cwd = os.getcwd()
        os.chdir(dirname)
        try:
            with _Quiet():
                self.L.amg_export(C.byref(self._last))
        finally:
            os.chdir(cwd)

    # single stages
    def coarsen(self, A, ctol=0.7):
        p = self.to_csr(*A)
        vc = np.zeros(A[3][0], np.float64)
        with _Quiet():
            self.L.coarsen(vc, p, ctol)
        return vc

    def mxm(self, A, B, iftrsp):
        X = self.new_csr()
        with _Quiet():
            self.L.mxm(X, self.to_csr(*A), self.to_csr(*B), float(iftrsp))
        return self.from_csr(X)

    def mpm(self, alpha, A, beta, B):
        X = self.new_csr()
        self.L.mpm(X, alpha, self.to_csr(*A), beta, self.to_csr(*B))
        return self.from_csr(X)

    def mxmpoint(self, A, B):
        X = self.new_csr()
        self.L.mxmpoint(X, self.to_csr(*A), self.to_csr(*B))
        return self.from_csr(X)

    def transpose(self, A):
        X = self.new_csr()
        self.L.transpose(X, self.to_csr(*A))
        return self.from_csr(X)

    def sub_mat(self, A, vr, vc):
        X = self.new_csr()
        self.L.sub_mat(X, self.to_csr(*A), np.ascontiguousarray(vr, np.float64), np.ascontiguousarray(vc, np.float64))
        return self.from_csr(X)

    def interpolation(self, Af, Ac, Ar, gamma2, tol):
        X = self.new_csr()
        with _Quiet():
            self.L.interpolation(X, self.to_csr(*Af), self.to_csr(*Ac), self.to_csr(*Ar), gamma2, tol)
        return self.from_csr(X)


def compare(H1, H2, rtol=0.0, what=("A", "Af", "W", "AfP"), params_rtol=None, names=("a", "b")):
    """Compare two hierarchies: structure must be identical; values identical (rtol=0) or within
    rtol relative to the largest magnitude in the matrix.  Returns a list of mismatch strings."""
    bad = []
    if H1.nlevels != H2.nlevels:
        return ["nlevels %d != %d" % (H1.nlevels, H2.nlevels)]
    if H1.nullspace != H2.nullspace:
        bad.append("nullspace %d != %d" % (H1.nullspace, H2.nullspace))
    for l, (a, b) in enumerate(zip(H1.levels, H2.levels)):
        for k in what:
            if k not in a and k not in b:
                continue
            ra, ca, va, sa = a[k]; rb, cb, vb, sb = b[k]
            if sa != sb:
                bad.append("L%d %s shape %s != %s" % (l, k, sa, sb)); continue
            if not np.array_equal(ra, rb) or not np.array_equal(ca, cb):
                bad.append("L%d %s pattern differs (nnz %d vs %d)" % (l, k, len(ca), len(cb))); continue
            if rtol == 0.0:
                if not np.array_equal(va, vb):
                    d = np.abs(va - vb).max() / max(np.abs(va).max(), 1e-300)
                    bad.append("L%d %s values differ (max rel %.3e)" % (l, k, d))
            elif len(va):
                d = np.abs(va - vb).max() / max(np.abs(va).max(), 1e-300)
                if not d <= rtol:
                    bad.append("L%d %s values rel err %.3e > %.1e" % (l, k, d, rtol))
        for k in ("C", "idc", "idf"):
            if k in a and not np.array_equal(a[k], b[k]):
                bad.append("L%d %s differs" % (l, k))
        if "D" in a:
            pr = rtol if params_rtol is None else params_rtol
            if pr == 0.0:
                if not np.array_equal(a["D"], b["D"]):
                    bad.append("L%d D differs (max rel %.3e)" % (l, np.abs(a["D"] / b["D"] - 1).max()))
                if a["m"] != b["m"] or a["rho"] != b["rho"]:
                    bad.append("L%d m/rho differ: %r %r vs %r %r" % (l, a["m"], a["rho"], b["m"], b["rho"]))
            else:
                d = np.abs(a["D"] / b["D"] - 1).max() if len(a["D"]) else 0.0
                if not d <= pr:
                    bad.append("L%d D rel err %.3e > %.1e" % (l, d, pr))
                if a["m"] != b["m"]:
                    bad.append("L%d m differs: %r vs %r" % (l, a["m"], b["m"]))
                if abs(a["rho"] - b["rho"]) > pr * max(abs(b["rho"]), 1e-300):
                    bad.append("L%d rho differs: %r vs %r" % (l, a["rho"], b["rho"]))
    return bad


# ------------------------------------------------------------------------------------------
# hierarchy fingerprint: the same number the product computes on the device
# (amgb_hierarchy_hash, omp_amg_b200/csrc/capi.cu) -- bench.py prints the product's, the tests
# compare it with this one computed from the oracle's hierarchy.
# ------------------------------------------------------------------------------------------
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _array_hash(words):
    """sum_i splitmix64(w[i] ^ ((i+1) * 0x9E3779B97F4A7C15)) mod 2^64 over 8-byte words."""
    w = np.ascontiguousarray(words).astype(np.uint64, copy=False)
    with np.errstate(over="ignore"):
        i = (np.arange(1, len(w) + 1, dtype=np.uint64)) * np.uint64(0x9E3779B97F4A7C15)
        z = w ^ i
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
        return int(np.add.reduce(z, dtype=np.uint64)) if len(z) else 0


def _fnv_words(h, *words):
    for x in words:
        x = int(x) & 0xFFFFFFFFFFFFFFFF
        for b in range(8):
            h ^= (x >> (8 * b)) & 0xFF
            h = (h * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def _bits(v):
    return np.ascontiguousarray(v, np.float64).view(np.uint64)


def hierarchy_hash(H):
    """Fingerprint of a whole hierarchy: nlevels, nullspace and per level the matrices A (and Af,
    W, AfP above the last level: row offsets, columns, value bits), C, D (bits), idc, idf, m and
    rho (bits)."""
    h = _fnv_words(0xCBF29CE484222325, H.nlevels, H.nullspace)
    for l, lev in enumerate(H.levels):
        last = l == H.nlevels - 1
        for which, name in enumerate(("A", "Af", "W", "AfP")):
            if last and which:
                continue
            ro, col, a, shape = lev[name]
            h = _fnv_words(h, l, which, shape[0], shape[1], len(col), _array_hash(ro), _array_hash(col),
                           _array_hash(_bits(a)))
        if not last:
            h = _fnv_words(h, _array_hash(_bits(lev["C"])), _array_hash(_bits(lev["D"])),
                           _array_hash(np.asarray(lev["idc"]).astype(np.uint64)),
                           _array_hash(np.asarray(lev["idf"]).astype(np.uint64)),
                           int(_bits([lev["m"]])[0]), int(_bits([lev["rho"]])[0]))
    return h

"""ctypes binding of include/omp_amg_b200.h and the host-side mirror of the reference API.

Reference interfaces mirrored here:
  * ``amg_setup(n, Ai, Aj, Av, data)``      amg_setup.h:5   -> :func:`amg_setup`
  * ``amg_export(data)``                    amg_setup.h:9   -> :meth:`Hierarchy.export`
  * ``struct amg_setup_data`` fields        amg_tools.h:29  -> :class:`Hierarchy` accessors
  * ``crs_setup / crs_solve / crs_stats / crs_free``  crs.h:14-22 -> functions of the same name
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_LIB_PATH = os.path.join(_HERE, "libomp_amg_b200.so")

A, AF, W, AFP = 0, 1, 2, 3
VEC_C, VEC_D, VEC_IDC, VEC_IDF = 0, 1, 2, 3


class AmgError(RuntimeError):
    pass


# int fn(void *buf, const long long *off, int size, void *user): in-place all-gather of byte segments
ALLGATHERV_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_longlong), C.c_int, C.c_void_p)


def lib_path():
    return _LIB_PATH


def _declare(L):
    i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
    f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
    vp = C.c_void_p
    L.amgb_last_error.restype = C.c_char_p
    L.amgb_build_info.restype = C.c_char_p
    L.amgb_init.argtypes = [C.c_int]
    L.amgb_setup.argtypes = [C.c_int64, i32p, i32p, f64p, C.POINTER(vp)]
    L.amgb_setup_device.argtypes = [C.c_int64, vp, vp, vp, C.POINTER(vp)]
    L.amgb_setup_from_dump.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.amgb_free.argtypes = [vp]
    L.amgb_free.restype = None
    L.amgb_nlevels.argtypes = [vp]
    L.amgb_nullspace.argtypes = [vp]
    L.amgb_level_info.argtypes = [vp, C.c_int, C.POINTER(C.c_int64)]
    L.amgb_level_params.argtypes = [vp, C.c_int, C.POINTER(C.c_double)]
    L.amgb_get_csr.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                               C.POINTER(C.c_int64), vp, vp, vp]
    L.amgb_get_vec.argtypes = [vp, C.c_int, C.c_int, f64p]
    L.amgb_export.argtypes = [vp, C.c_char_p]
    L.amgb_solve.argtypes = [vp, f64p, f64p]
    L.amgb_solve_device.argtypes = [vp, vp, vp]
    L.amgb_timing.argtypes = [vp, C.POINTER(C.c_double)]
    L.amgb_spmv_stats_enable.argtypes = [C.c_int]
    L.amgb_spmv_stats.argtypes = [vp, C.POINTER(C.c_double)]
    L.amgb_set_reduce_mode.argtypes = [C.c_int]
    L.amgb_release_memory.restype = None
    L.amgb_peak_device_bytes.restype = C.c_int64
    L.amgb_debug_dot.argtypes = [f64p, C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_double)]
    L.amgb_debug_spgemm.argtypes = [C.c_int32, C.c_int32, i32p, i32p, f64p, C.c_int32, C.c_int32, i32p, i32p, f64p,
                                    C.c_int64, C.POINTER(C.c_int64), i32p, i32p, f64p]
    L.amgb_trace_enable.argtypes = [C.c_int]
    L.amgb_trace_enable.restype = None
    L.amgb_trace_get.argtypes = [C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_int64)]
    u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
    u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
    L.crs_amg_setup_u32.argtypes = [C.c_uint32, u64p, C.c_uint32, u32p, u32p, f64p, C.c_uint32, vp]
    L.crs_amg_setup_u32.restype = vp
    L.crs_amg_setup_u64.argtypes = [C.c_uint64, u64p, C.c_uint64, u64p, u64p, f64p, C.c_uint64, vp]
    L.crs_amg_setup_u64.restype = vp
    L.amgb_launch_count.restype = C.c_int64
    L.amgb_sync_count.restype = C.c_int64
    L.amgb_hierarchy_hash.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.amgb_solve_device_repeat.argtypes = [vp, vp, vp, C.c_int]
    L.crs_amg_solve.argtypes = [f64p, vp, f64p]
    L.crs_amg_solve.restype = None
    L.crs_amg_stats.argtypes = [vp]
    L.crs_amg_stats.restype = None
    L.crs_amg_free.argtypes = [vp]
    L.crs_amg_free.restype = None
    L.crs_amg_hierarchy.argtypes = [vp]
    L.crs_amg_hierarchy.restype = vp
    L.amgb_comm_unique_id.argtypes = [C.c_char_p]
    L.amgb_comm_init.argtypes = [C.c_int, C.c_int, C.c_char_p]
    L.amgb_comm_init_host.argtypes = [C.c_int, C.c_int, ALLGATHERV_FN, vp]
    L.amgb_comm_stats.argtypes = [C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.amgb_partition_solve_storage.argtypes = [vp, C.POINTER(C.c_int64)]
    return L


def lib(path=None):
    """Load the CUDA library.  Raises if it has not been built: there is no fallback."""
    global _LIB
    if path is not None:
        return _declare(C.CDLL(path))
    if _LIB is None:
        if not os.path.exists(_LIB_PATH):
            raise AmgError("%s is missing: run `make -C omp_amg_b200/csrc` (or __graft_entry__.build()); "
                           "omp_amg_b200 has no CPU implementation" % _LIB_PATH)
        _LIB = _declare(C.CDLL(_LIB_PATH))
    return _LIB


REDUCE_TREE, REDUCE_SEQUENTIAL = 0, 1


def set_reduce_mode(mode, L=None):
    """0: parallel tree (fast); 1: left-to-right like the reference (bit-identical results)."""
    L = L or lib()
    _check(L, L.amgb_set_reduce_mode(int(mode)))


def debug_dot(a, b=None, mode=REDUCE_SEQUENTIAL, L=None):
    """The library's reduction kernel on host vectors (diagnostics / tests)."""
    L = L or lib()
    a = np.ascontiguousarray(a, np.float64)
    out = C.c_double()
    if b is None:
        _check(L, L.amgb_debug_dot(a, None, len(a), mode, C.byref(out)))
    else:
        b = np.ascontiguousarray(b, np.float64)
        _check(L, L.amgb_debug_dot(a, b.ctypes.data, len(a), mode, C.byref(out)))
    return out.value


def debug_spgemm(A, B, L=None):
    """The library's SpGEMM on host CSR operands ``(ro, col, a, (rn, cn))`` -> the same tuple for X
    (diagnostics / tests of the ``mxm`` primitive in isolation)."""
    L = L or lib()
    aro, acol, aa, (arn, acn) = A
    bro, bcol, ba, (brn, bcn) = B
    c32 = lambda v: np.ascontiguousarray(v, np.int32) if len(v) else np.zeros(1, np.int32)
    c64 = lambda v: np.ascontiguousarray(v, np.float64) if len(v) else np.zeros(1, np.float64)
    aro, acol, aa, bro, bcol, ba = c32(aro), c32(acol), c64(aa), c32(bro), c32(bcol), c64(ba)
    # every product could survive: bound of the output size
    cap = int(sum(int(bro[k + 1] - bro[k]) for k in acol[:int(aro[arn])])) + 1
    xro = np.zeros(arn + 1, np.int32); xcol = np.zeros(cap, np.int32); xa = np.zeros(cap, np.float64)
    nnz = C.c_int64()
    _check(L, L.amgb_debug_spgemm(arn, acn, aro, acol, aa, brn, bcn, bro, bcol, ba, cap, C.byref(nnz), xro, xcol, xa))
    return xro, xcol[:nnz.value], xa[:nnz.value], (arn, bcn)


def debug_spgemm_tiers(L=None):
    """Rows per SpGEMM tier of the last :func:`debug_spgemm` call: (as binned, after hand-downs)."""
    L = L or lib()
    out = (C.c_int32 * 22)()
    L.amgb_debug_spgemm_tiers(out)
    return list(out[:11]), list(out[11:])


def build_info(L=None):
    return (L or lib()).amgb_build_info().decode()


def device_count(L=None):
    return (L or lib()).amgb_device_count()


def _check(L, rc):
    if rc != 0:
        raise AmgError("omp_amg_b200 error %d: %s" % (rc, L.amgb_last_error().decode()))


def comm_init(L=None, device=None):
    """Join the ranks of the current ``torch.distributed`` job (one process per GPU) so that
    ``amg_setup`` partitions its SpGEMM rows and local solves over them and exchanges the blocks
    through NCCL.  torch.distributed only carries the 128-byte NCCL id from rank 0 to the others;
    the data path is the library's own communicator.  Every rank must then call ``amg_setup`` with
    the same matrix."""
    import torch
    import torch.distributed as dist
    L = L or lib()
    rank, size = dist.get_rank(), dist.get_world_size()
    if device is not None:
        _check(L, L.amgb_init(int(device)))
    buf = C.create_string_buffer(128)
    if rank == 0:
        _check(L, L.amgb_comm_unique_id(buf))
    on_gpu = dist.get_backend() == "nccl"
    t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    if on_gpu:
        t = t.cuda()
    dist.broadcast(t, src=0)
    ident = bytes(t.cpu().numpy().tobytes())
    _check(L, L.amgb_comm_init(rank, size, ident))
    return rank, size


def comm_init_host_gloo(L):
    """Host transport over the current (gloo) process group for the host-emulation build -- a test
    tool for the partitioning logic on CPU.  The product build refuses it (error -112)."""
    import torch
    import torch.distributed as dist
    rank, size = dist.get_rank(), dist.get_world_size()

    def gather(buf, off, n, user):
        try:
            for r in range(n):
                nb = off[r + 1] - off[r]
                if nb <= 0:
                    continue
                seg = np.ctypeslib.as_array((C.c_ubyte * nb).from_address(buf + off[r]))
                dist.broadcast(torch.from_numpy(seg), src=r)
            return 0
        except Exception:           # never unwind through the C frames
            import traceback
            traceback.print_exc()
            return 1

    cb = ALLGATHERV_FN(gather)
    L._amgb_host_cb = cb            # keep the trampoline alive as long as the library handle
    _check(L, L.amgb_comm_init_host(rank, size, cb, None))
    return rank, size


def comm_finalize(L=None):
    L = L or lib()
    _check(L, L.amgb_comm_finalize())


class Hierarchy:
    """Handle to a hierarchy resident in HBM (the reference's ``struct amg_setup_data``)."""

    def __init__(self, L, handle):
        self._L = L
        self._h = handle

    # -- struct amg_setup_data --
    @property
    def nlevels(self):
        return self._L.amgb_nlevels(self._h)

    @property
    def nullspace(self):
        return self._L.amgb_nullspace(self._h)

    def level_info(self, lvl):
        info = (C.c_int64 * 10)()
        _check(self._L, self._L.amgb_level_info(self._h, lvl, info))
        keys = ("n", "nnz", "nf", "nc", "nnzf", "nnzw", "nnzfp", "coarsen_rounds", "lanczos_iters", "interp_rounds")
        return dict(zip(keys, list(info)))

    def level_params(self, lvl):
        par = (C.c_double * 4)()
        _check(self._L, self._L.amgb_level_params(self._h, lvl, par))
        return {"m": par[0], "rho": par[1], "lambda_min": par[2], "lambda_max": par[3]}

    def csr(self, lvl, which):
        """(row_off, col, a, (rn, cn)) of data->A/Af/W/AfP[lvl], copied to host."""
        rn, cn, nnz = C.c_int32(), C.c_int32(), C.c_int64()
        _check(self._L, self._L.amgb_get_csr(self._h, lvl, which, C.byref(rn), C.byref(cn), C.byref(nnz),
                                              None, None, None))
        ro = np.zeros(rn.value + 1, np.int32)
        col = np.zeros(max(nnz.value, 1), np.int32)
        a = np.zeros(max(nnz.value, 1), np.float64)
        _check(self._L, self._L.amgb_get_csr(self._h, lvl, which, None, None, None, ro.ctypes.data,
                                              col.ctypes.data, a.ctypes.data))
        return ro, col[:nnz.value], a[:nnz.value], (rn.value, cn.value)

    def vec(self, lvl, which):
        info = self.level_info(lvl)
        ln = {VEC_C: info["n"], VEC_D: info["nf"], VEC_IDC: info["nc"], VEC_IDF: info["nf"]}[which]
        out = np.zeros(max(ln, 1), np.float64)
        _check(self._L, self._L.amgb_get_vec(self._h, lvl, which, out))
        return out[:ln]

    # -- amg_export --
    def export(self, dirname):
        _check(self._L, self._L.amgb_export(self._h, os.fsencode(dirname)))

    # -- amg_exec + crs_solve projection --
    def solve(self, b):
        b = np.ascontiguousarray(b, np.float64)
        x = np.zeros_like(b)
        _check(self._L, self._L.amgb_solve(self._h, x, b))
        return x

    def solve_device(self, x_ptr, b_ptr):
        _check(self._L, self._L.amgb_solve_device(self._h, x_ptr, b_ptr))

    def solve_device_repeat(self, x_ptr, b_ptr, repeat):
        """``repeat`` V-cycles on device vectors, one host synchronisation at the end."""
        _check(self._L, self._L.amgb_solve_device_repeat(self._h, x_ptr, b_ptr, int(repeat)))

    def hash(self):
        """Fingerprint of the whole hierarchy, computed on the device (amgb_hierarchy_hash)."""
        out = C.c_uint64()
        _check(self._L, self._L.amgb_hierarchy_hash(self._h, C.byref(out)))
        return out.value

    def timing(self):
        t = (C.c_double * 16)()
        _check(self._L, self._L.amgb_timing(self._h, t))
        keys = ("total", "build_csr", "coarsen", "smoother", "lanczos", "interp", "galerkin",
                "spgemm_device_s", "spgemm_bytes", "spgemm_calls", "launches", "syncs", "device_total_s",
                "comm_calls", "comm_bytes", "comm_device_s")
        return dict(zip(keys, list(t)))

    def partition_solve_storage(self):
        """Several ranks (collective): keep only this rank's row blocks of the matrices the V-cycle
        applies row-partitioned, release the rest; returns the device bytes released.  The hierarchy
        then serves ``solve`` only."""
        f = C.c_int64()
        _check(self._L, self._L.amgb_partition_solve_storage(self._h, C.byref(f)))
        return int(f.value)

    def spmv_stats(self):
        """(device seconds, algorithmic bytes, calls) of the long-row SpMV kernels of this setup
        (zeros unless amgb_spmv_stats_enable(1) was called before it)."""
        t = (C.c_double * 3)()
        _check(self._L, self._L.amgb_spmv_stats(self._h, t))
        return float(t[0]), int(t[1]), int(t[2])

    def free(self):
        if self._h:
            self._L.amgb_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def amg_setup(Ai, Aj, Av, L=None, device_ptrs=False, nnz=None):
    """``amg_setup`` (amg_setup.c:60).  ``Ai, Aj`` 0-based int32, ``Av`` float64 host arrays; with
    ``device_ptrs=True`` they are raw device addresses (ints) of arrays already in HBM."""
    L = L or lib()
    h = C.c_void_p()
    if device_ptrs:
        _check(L, L.amgb_setup_device(int(nnz), C.c_void_p(Ai), C.c_void_p(Aj), C.c_void_p(Av), C.byref(h)))
    else:
        Ai = np.ascontiguousarray(Ai, np.int32)
        Aj = np.ascontiguousarray(Aj, np.int32)
        Av = np.ascontiguousarray(Av, np.float64)
        _check(L, L.amgb_setup(len(Av), Ai, Aj, Av, C.byref(h)))
    return Hierarchy(L, h)


def amg_setup_from_dump(dirname, L=None):
    """The reference driver's input path (serial_amg.c:76-96): amgdmp_{i,j,p}.dat in ``dirname``."""
    L = L or lib()
    h = C.c_void_p()
    _check(L, L.amgb_setup_from_dump(os.fsencode(dirname), C.byref(h)))
    return Hierarchy(L, h)


class _Crs:
    def __init__(self, L, ptr, n):
        self.L, self.ptr, self.n = L, ptr, n


def crs_setup(n, id, nz, Ai, Aj, A, null_space, comm=None, L=None, uint_bits=64):
    """``crs_setup`` (crs.h:14): same argument order and meaning as the reference.  ``uint_bits``
    is the width of gslib's build-time ``uint``: 64 for the reference's own build (-DUSE_LONG,
    types.h:52-64), 32 for a gslib built without it.  ``comm`` is the address of a ``struct comm``
    (comm.h:85) or None."""
    L = L or lib()
    id = np.ascontiguousarray(id, np.uint64)
    ut = np.uint64 if uint_bits == 64 else np.uint32
    Ai = np.ascontiguousarray(Ai, ut)
    Aj = np.ascontiguousarray(Aj, ut)
    A = np.ascontiguousarray(A, np.float64)
    fn = L.crs_amg_setup_u64 if uint_bits == 64 else L.crs_amg_setup_u32
    p = fn(n, id, nz, Ai, Aj, A, null_space, comm)
    if not p:
        raise AmgError("crs_setup failed: %s" % L.amgb_last_error().decode())
    return _Crs(L, p, n)


def crs_solve(x, data, b):
    """``crs_solve`` (crs.h:18): x and b are local vectors of length n."""
    b = np.ascontiguousarray(b, np.float64)
    assert x.dtype == np.float64 and x.flags.c_contiguous
    data.L.crs_amg_solve(x, data.ptr, b)


def crs_stats(data):
    data.L.crs_amg_stats(data.ptr)


def crs_free(data):
    if data.ptr:
        data.L.crs_amg_free(data.ptr)
        data.ptr = None

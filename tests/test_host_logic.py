"""CPU tests of the product's host side: the C ABI surface, the loud failure without a GPU,
the host orchestration (through the host-emulation build of the same sources), the reference-
format files, the crs_* mirror and the matrix generators."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from util import ROOT, EMU_SO, orc, amg, api, fetch, product_trace, first_trace_mismatch
from omp_amg_b200 import matrices as M

GOLDEN = os.path.join(ROOT, "tests", "golden")
HEADER = os.path.join(ROOT, "include", "omp_amg_b200.h")


@pytest.fixture(scope="module")
def emu():
    subprocess.run(["make", "-j4", "-C", os.path.join(ROOT, "omp_amg_b200", "csrc"), "emu"], check=True,
                   stdout=subprocess.DEVNULL)
    L = api.lib(EMU_SO)
    assert "emulation" in api.build_info(L)
    return L


@pytest.fixture(scope="module")
def O():
    orc.build(ref=False)
    return orc.Oracle()


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:amgb|crs_amg)_[a-z_0-9]+)\s*\(", src)))


def test_cuda_library_exports_every_declared_symbol():
    path = amg.lib_path()
    if not os.path.exists(path):
        subprocess.run(["make", "-j4", "-C", os.path.join(ROOT, "omp_amg_b200", "csrc")], check=True,
                       stdout=subprocess.DEVNULL)
    L = ctypes.CDLL(path)
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), "include/omp_amg_b200.h declares %s but the library does not export it" % s


def test_cuda_library_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = amg.lib()
    assert L.amgb_device_count() == 0
    assert L.amgb_init(0) != 0
    with pytest.raises(amg.AmgError, match="no CUDA device"):
        amg.amg_setup(*M.poisson7(3))


def test_product_package_never_touches_the_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "omp_amg_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert "amg_oracle" not in txt and "import oracle" not in txt and "from oracle" not in txt, f


CASES = [("dump", 0), ("poisson7", 6), ("poisson7", 11), ("poisson27", 6), ("aniso7", 9), ("sem_hex", 7)]


@pytest.mark.parametrize("mode", [orc.SEQ, orc.TREE])
@pytest.mark.parametrize("name,n", CASES)
def test_host_orchestration_bit_identical_to_oracle(emu, O, name, n, mode):
    """The host-emulation build runs the product's orchestration (same sources, kernels executed
    as loops) and must agree with the oracle in every traced intermediate array, in both
    reduction modes (sequential = the reference's order, tree = the fast mode)."""
    mat = M.read_amgdmp(GOLDEN) if name == "dump" else M.by_name(name, n)
    h = O.setup_raw(*mat, mode, trace=True)
    want = O.fetch(h); twant = O.trace(); O.free(h)
    api.set_reduce_mode(api.REDUCE_SEQUENTIAL if mode == orc.SEQ else api.REDUCE_TREE, L=emu)
    emu.amgb_trace_enable(1)
    try:
        H = amg.amg_setup(*mat, L=emu)
        tgot = product_trace(emu)
    finally:
        emu.amgb_trace_enable(0)
        api.set_reduce_mode(api.REDUCE_SEQUENTIAL, L=emu)
    assert first_trace_mismatch(tgot, twant) is None
    assert orc.compare(fetch(H), want) == []


def test_ragged_and_degenerate_inputs(emu, O):
    # zeros in the COO list are dropped, empty rows/columns vanish (build_csr, amg_setup.c:3612)
    Ai, Aj, Av = M.poisson7(4, 3, 2)
    extra_i = np.array([30, 40], np.int32); extra_j = np.array([31, 40], np.int32)
    Ai2 = np.concatenate([Ai, extra_i]); Aj2 = np.concatenate([Aj, extra_j]); Av2 = np.concatenate([Av, [0.0, 0.0]])
    a = fetch(amg.amg_setup(Ai2, Aj2, Av2, L=emu))
    b = O.setup(Ai, Aj, Av, orc.SEQ)
    assert orc.compare(a, b) == []
    # unsorted input gives the same hierarchy
    perm = np.random.default_rng(0).permutation(len(Av))
    c = fetch(amg.amg_setup(Ai[perm], Aj[perm], Av[perm], L=emu))
    assert orc.compare(c, b) == []
    # 1x1 and 2x2
    one = fetch(amg.amg_setup(np.array([0], np.int32), np.array([0], np.int32), np.array([2.0]), L=emu))
    assert one.nlevels == 1 and one.nullspace == 0
    two = amg.amg_setup(np.array([0, 0, 1, 1], np.int32), np.array([0, 1, 0, 1], np.int32),
                        np.array([2.0, -1.0, -1.0, 2.0]), L=emu)
    assert orc.compare(fetch(two), O.setup(np.array([0, 0, 1, 1], np.int32), np.array([0, 1, 0, 1], np.int32),
                                           np.array([2.0, -1.0, -1.0, 2.0]), orc.SEQ)) == []
    # errors: duplicates, empty input
    with pytest.raises(amg.AmgError, match="duplicate"):
        amg.amg_setup(np.array([0, 0, 1], np.int32), np.array([0, 0, 1], np.int32), np.array([1.0, 1.0, 1.0]), L=emu)
    with pytest.raises(amg.AmgError):
        amg.amg_setup(np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0), L=emu)


def test_export_matches_reference_files(emu, tmp_path):
    z = np.load(os.path.join(GOLDEN, "ref_dump.npz"))
    H = amg.amg_setup_from_dump(GOLDEN, L=emu)
    H.export(str(tmp_path))
    for f in ("amg_W.dat", "amg_AfP.dat", "amg_Aff.dat", "amg.dat"):
        got = np.fromfile(os.path.join(str(tmp_path), f)); want = z["file_" + f.replace(".", "_")]
        assert got.shape == want.shape
        assert np.array_equal(got, want), f      # sequential reductions: the files are identical


def test_vcycle_matches_oracle_and_reduces_residual(emu, O):
    mat = M.sem_hex(6)
    H = amg.amg_setup(*mat, L=emu)
    h = O.setup_raw(*mat, orc.SEQ)
    n = H.level_info(0)["n"]
    rng = np.random.default_rng(3)
    b = rng.standard_normal(n); b -= b.mean()
    x = H.solve(b); xo = O.solve(h, b)
    O.free(h)
    assert np.abs(x - xo).max() <= 1e-12 * np.abs(xo).max()
    # the host-emulation build runs the same operations in the same order: identical bits (the
    # oracle's V-cycle is itself pinned to the reference's amg_exec, tests/test_oracle.py)
    assert np.array_equal(x, xo)
    import scipy.sparse as sp
    A = sp.coo_matrix((mat[2], (mat[0], mat[1]))).tocsr()
    # one V-cycle as a preconditioner must contract the error of a few Richardson steps
    r0 = np.linalg.norm(b)
    xk = np.zeros(n)
    for _ in range(8):
        r = b - A @ xk; r -= r.mean()
        xk += H.solve(r)
    assert np.linalg.norm(b - A @ xk - (b - A @ xk).mean()) < 1e-3 * r0


def test_crs_interface_mirrors_reference_test(emu):
    """crs_test.c: the 2x2-element assembly with shared ids, null_space=1."""
    A = np.array([2, -1, -1, 0, -1, 2, 0, -1, -1, 0, 2, -1, 0, -1, -1, 2], np.float64)
    Ai = np.repeat(np.arange(4), 4).astype(np.uint32); Aj = np.tile(np.arange(4), 4).astype(np.uint32)
    # two elements sharing an edge: local dofs of both elements in one call
    ids = np.array([1, 2, 4, 5, 2, 3, 5, 6], np.uint64)
    Ai2 = np.concatenate([Ai, Ai + 4]).astype(np.uint32); Aj2 = np.concatenate([Aj, Aj + 4]).astype(np.uint32)
    A2 = np.concatenate([A, A])
    d = amg.crs_setup(8, ids, 32, Ai2, Aj2, A2, 1, None, L=emu, uint_bits=32)
    b = np.array([1.0, -1, 0.5, -0.5, 0, 0, 0, 0])
    x = np.zeros(8)
    amg.crs_solve(x, d, b)
    # shared dofs get one value
    assert x[1] == x[4] and x[3] == x[6]
    assert abs(sum(x[[0, 1, 2, 3, 5, 7]])) < 1e-12     # mean projected out
    amg.crs_stats(d)
    amg.crs_free(d)
    with pytest.raises(amg.AmgError):
        amg.crs_setup(2, np.array([0, 0], np.uint64), 0, np.zeros(0, np.uint32), np.zeros(0, np.uint32),
                      np.zeros(0), 0, None, L=emu, uint_bits=32)


@pytest.mark.parametrize("bits", [32, 64])
def test_crs_setup_reads_gslib_struct_comm(emu, bits):
    """A real gslib caller always passes a struct comm (comm.h:85: ``uint id, np; comm_ext c;`` with
    the build-time uint): np == 1 must be accepted and np == 2 rejected, for both integer widths
    (the reference Makefile's -DUSE_LONG build binds crs_amg_setup_u64)."""
    UI = ctypes.c_uint64 if bits == 64 else ctypes.c_uint32

    class Comm(ctypes.Structure):
        _fields_ = [("id", UI), ("np", UI), ("c", ctypes.c_int)]     # comm_ext is an int without MPI

    Ai, Aj, Av = M.poisson7(3)
    n = 27
    ids = np.arange(1, n + 1, dtype=np.uint64)
    one = Comm(0, 1, 0)
    d = amg.crs_setup(n, ids, len(Av), Ai, Aj, Av, 0, ctypes.addressof(one), L=emu, uint_bits=bits)
    x = np.zeros(n)
    amg.crs_solve(x, d, np.ones(n))
    assert np.isfinite(x).all() and np.abs(x).max() > 0
    amg.crs_free(d)
    two = Comm(1, 2, 0)
    with pytest.raises(amg.AmgError, match="np == 1"):
        amg.crs_setup(n, ids, len(Av), Ai, Aj, Av, 0, ctypes.addressof(two), L=emu, uint_bits=bits)


def test_unmodified_reference_driver_links_against_the_shim(emu, tmp_path):
    """serial_amg.c (and fail.c) of the reference, compiled UNMODIFIED with the reference's own
    flags (-DUSE_LONG: uint = unsigned long) and linked against the reference-shaped entry points
    amg_setup / amg_export / free_data of include/amg_setup_b200.h (host-emulation build here; the
    GPU test runs the same driver against the CUDA library): its four output files must equal the
    ones the reference's own serial_amg wrote (tests/golden/ref_dump.npz)."""
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "drivers"], check=True, stdout=subprocess.DEVNULL)
    exe = os.path.join(ROOT, "oracle", "_ref", "serial_amg_emu")
    if not os.path.exists(exe):
        pytest.skip("needs /root/reference to compile the unmodified driver")
    import shutil
    for k in "ijp":
        shutil.copy(os.path.join(GOLDEN, "amgdmp_%s.dat" % k), str(tmp_path))
    r = subprocess.run([exe], cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "Level 1, dim(A) = 49" in r.stdout and "Nullspace = 1" in r.stdout
    z = np.load(os.path.join(GOLDEN, "ref_dump.npz"))
    for f in ("amg.dat", "amg_W.dat", "amg_AfP.dat", "amg_Aff.dat"):
        assert np.array_equal(np.fromfile(os.path.join(str(tmp_path), f)), z["file_" + f.replace(".", "_")]), f


def test_shim_fills_struct_amg_setup_data(emu):
    """amg_setup of the shim fills every field of struct amg_setup_data (amg_tools.h:29-52) with the
    reference's build-time uint (unsigned long); compared with the fixture the reference produced."""
    from test_oracle import load_golden
    want, mat, _ = load_golden(os.path.join(GOLDEN, "ref_dump.npz"))
    S = ctypes.CDLL(os.path.join(ROOT, "tests", "_emu", "libamg_setup_emu.so"))
    data = orc._RData()
    Ai = np.ascontiguousarray(mat[0], np.uint64); Aj = np.ascontiguousarray(mat[1], np.uint64)
    Av = np.ascontiguousarray(mat[2], np.float64)
    S.amg_setup.argtypes = [ctypes.c_ulong, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(orc._RData)]
    with orc._Quiet():
        S.amg_setup(len(Av), Ai.ctypes.data, Aj.ctypes.data, Av.ctypes.data, ctypes.byref(data))
    assert int(data.nlevels) == want.nlevels and int(data.nullspace) == want.nullspace
    assert data.tolc == 0.7 and abs(data.gamma ** 2 - (1 - np.sqrt(0.5))) < 1e-15
    for l, lev in enumerate(want.levels):
        A = orc.Ref.from_csr(data.A[l])
        assert A[3] == lev["A"][3] and np.array_equal(A[0], lev["A"][0]) and np.array_equal(A[1], lev["A"][1])
        assert np.array_equal(A[2], lev["A"][2])
        assert data.n[l] == lev["A"][3][0] and data.nnz[l] == len(lev["A"][1])
        if l == want.nlevels - 1:
            break
        for name, arr in (("Af", data.Af), ("W", data.W), ("AfP", data.AfP)):
            X = orc.Ref.from_csr(arr[l])
            assert np.array_equal(X[0], lev[name][0]) and np.array_equal(X[1], lev[name][1]) and np.array_equal(X[2], lev[name][2]), name
        n = lev["A"][3][0]; nf, nc = lev["W"][3]
        assert np.array_equal(np.ctypeslib.as_array(data.C[l], (n,)), lev["C"])
        assert np.array_equal(np.ctypeslib.as_array(data.F[l], (n,)), 1.0 - lev["C"])
        assert np.array_equal(np.ctypeslib.as_array(data.D[l], (nf,)), lev["D"])
        assert np.array_equal(np.ctypeslib.as_array(data.idc[l], (nc,)).astype(np.float64), lev["idc"])
        assert np.array_equal(np.ctypeslib.as_array(data.idf[l], (nf,)).astype(np.float64), lev["idf"])
        assert data.m[l] == lev["m"] and data.rho[l] == lev["rho"]
        assert data.nnzf[l] == len(lev["Af"][1]) and data.nnzfp[l] == len(lev["AfP"][1])
    assert np.array_equal(np.ctypeslib.as_array(data.id, (49,)), np.arange(1, 50))
    S.free_data.argtypes = [ctypes.c_void_p]


def test_hierarchy_hash_matches_python_definition(emu, O):
    """amgb_hierarchy_hash (what bench.py prints on every GPU count) equals oracle.hierarchy_hash of
    the oracle's hierarchy for the same input, and differs when one value bit differs."""
    for mat in (M.read_amgdmp(GOLDEN), M.poisson7(7), M.sem_hex(5)):
        H = amg.amg_setup(*mat, L=emu)
        want = O.setup(*mat, orc.SEQ)
        assert H.hash() == orc.hierarchy_hash(want) == orc.hierarchy_hash(fetch(H))
        want.levels[0]["W"][2][0] = np.nextafter(want.levels[0]["W"][2][0], 1e300)
        assert H.hash() != orc.hierarchy_hash(want)


def test_trace_fixtures_are_well_formed():
    """The committed oracle traces at sizes the oracle needs minutes for (make_trace_fixtures.py)."""
    import glob
    from util import load_trace_fixture
    files = sorted(glob.glob(os.path.join(GOLDEN, "trace_*.json.gz")))
    assert len(files) >= 4
    rows = 0
    for f in files:
        fx = load_trace_fixture(f)
        assert fx["nlevels"] == len(fx["levels"]) and len(fx["trace"]) > 100
        assert fx["levels"][-1]["info"][0] <= 1 and len(fx["hierarchy_hash"]) == 16
        rows = max(rows, fx["levels"][0]["info"][0])
    assert rows >= 64 ** 3


def test_matrix_generators():
    Ai, Aj, Av = M.poisson7(5)
    assert len(Av) == 125 * 7 - 6 * 25 and Av.sum() == 6 * 25
    Ai, Aj, Av = M.poisson27(4)
    assert (Av[Ai == Aj] == 26).all()
    import scipy.sparse as sp
    for gen in (lambda: M.aniso7(6), lambda: M.sem_hex(4)):
        Ai, Aj, Av = gen()
        A = sp.coo_matrix((Av, (Ai, Aj))).tocsr()
        assert abs(A - A.T).max() < 1e-12
        key = Ai.astype(np.int64) * (Ai.max() + 1) + Aj
        assert (np.diff(key) > 0).all()
    Ai, Aj, Av = M.sem_hex(4)
    A = sp.coo_matrix((Av, (Ai, Aj))).tocsr()
    assert np.abs(A @ np.ones(A.shape[0])).max() < 1e-12      # Neumann: constants in the null space


def test_amgdmp_round_trip(tmp_path):
    Ai, Aj, Av = M.poisson7(3)
    M.write_amgdmp(str(tmp_path), Ai, Aj, Av)
    Bi, Bj, Bv = M.read_amgdmp(str(tmp_path))
    assert np.array_equal(Ai, Bi) and np.array_equal(Aj, Bj) and np.array_equal(Av, Bv)


def test_transports_are_not_interchangeable(emu):
    """The product build exchanges through NCCL only: it refuses a host transport (-112), so the
    CPU tests of the partitioning logic can never be mistaken for the product path; the
    host-emulation build in turn has no NCCL."""
    import ctypes as C
    L = amg.lib()
    cb = api.ALLGATHERV_FN(lambda buf, off, n, user: 0)
    assert L.amgb_comm_init_host(0, 2, cb, None) == -112
    assert b"NCCL" in L.amgb_last_error()
    assert L.amgb_comm_size() == 1 and L.amgb_comm_rank() == 0
    buf = C.create_string_buffer(128)
    assert emu.amgb_comm_unique_id(buf) == -112
    assert emu.amgb_comm_init(0, 2, buf) == -112
    # a one-rank host communicator is a no-op and must leave the setup untouched
    assert emu.amgb_comm_init_host(0, 1, api.ALLGATHERV_FN(0), None) == 0
    assert emu.amgb_comm_size() == 1
    assert emu.amgb_comm_finalize() == 0


def test_c_host_program_links_against_the_abi(tmp_path):
    """the C drop-in for serial_amg.c compiles against include/ and links
    against the shared library with plain gcc; without a GPU it reports the error and exits 1
    (there is no CPU fallback), with one it writes the four amg*.dat files."""
    import shutil
    import torch
    src = tmp_path / "serial_amg_b200.c"
    src.write_text(r'''
#include <stdio.h>
#include "omp_amg_b200.h"
int main(void)
{
  amgb_hier *h;
  if (amgb_setup_from_dump(".", &h)) {
    fprintf(stderr, "amg setup failed: %s\n", amgb_last_error());
    return 1;
  }
  if (amgb_export(h, ".")) return 1;
  printf("levels %d\n", amgb_nlevels(h));
  amgb_free(h);
  return 0;
}
''')
    libdir = os.path.join(ROOT, "omp_amg_b200")
    exe = tmp_path / "serial_amg_b200"
    subprocess.run(["gcc", str(src), "-I" + os.path.join(ROOT, "include"), "-L" + libdir, "-lomp_amg_b200",
                    "-Wl,-rpath," + libdir, "-o", str(exe)], check=True)
    for f in ("amgdmp_i.dat", "amgdmp_j.dat", "amgdmp_p.dat"):
        shutil.copy(os.path.join(ROOT, "tests", "golden", f), tmp_path / f)
    r = subprocess.run([str(exe)], cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0 and "levels 4" in r.stdout
        assert all((tmp_path / f).exists() for f in ("amg.dat", "amg_W.dat", "amg_AfP.dat", "amg_Aff.dat"))
    else:
        assert r.returncode == 1 and "amg setup failed" in r.stderr


def test_spgemm_primitive_against_exact_order_reference(emu):
    """amgb_debug_spgemm (the mxm primitive in isolation) on operands that reach every SpGEMM bin:
    the host-emulation build must reproduce the exact-order reference bit for bit, zero drops
    included.  (The GPU suite runs the same operands through the CUDA kernels.)"""
    from util import spgemm_adversarial_operands, spgemm_reference, same_csr_bits
    A, B = spgemm_adversarial_operands(0)
    want = spgemm_reference(A, B)
    got = api.debug_spgemm(A, B, L=emu)
    assert got[3] == want[3] and same_csr_bits(got, want)
    dropped = sum(int(B[0][k + 1] - B[0][k]) for k in A[1]) - len(want[1])
    assert dropped > 0 and len(want[1]) > 100000


def test_bench_reference_arm_reports_an_unscaled_measurement():
    """`bench.py --impl reference` (the CPU port on the host cores) must print one JSON line whose
    value is the measured time of the config it names -- no extrapolation to the benchmark size."""
    import json
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                        "--warmup", "1", "--sample-n", "10", "--ref-budget", "30"], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "amg_setup_time" and d["higher_is_better"] is False
    assert d["config"]["workload"] == "poisson7_10^3_full_hierarchy_setup" and d["config"]["rows"] == 1000
    assert d["config"]["same_config_as_gpu_arm"] is False
    assert d["value"] == d["cpu_baseline"]["value"] == d["e2e"]["value"] and d["cpu_baseline"]["cores"] == 1
    assert abs(d["ms_per_step"] - 1e3 * d["value"]) < 1e-6


def test_single_rank_partition_calls_are_no_ops(emu):
    """One rank: amgb_partition_solve_storage releases nothing and leaves the hierarchy whole
    (accessors and fingerprint keep working); the opt-in SpMV statistics are zero unless enabled."""
    H = amg.amg_setup(*M.poisson7(6), L=emu)
    h0 = H.hash()
    assert H.partition_solve_storage() == 0
    assert H.hash() == h0
    ro, col, a, shape = H.csr(0, api.W)
    assert len(col) == ro[-1]
    assert H.spmv_stats() == (0.0, 0, 0)
    cc, cb = ctypes.c_int64(-1), ctypes.c_int64(-1)
    assert emu.amgb_comm_stats(ctypes.byref(cc), ctypes.byref(cb)) == 0
    assert cc.value == 0 and cb.value == 0

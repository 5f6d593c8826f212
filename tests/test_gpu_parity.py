"""GPU parity tests (run on the B200 box): everything goes through the C ABI of
libomp_amg_b200.so and is compared with the oracle, the reference fixtures, or a
size-independent property."""
import os

import numpy as np
import pytest

from util import ROOT, orc, amg, api, fetch, product_trace, first_trace_mismatch
from omp_amg_b200 import matrices as M

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def L():
    lib = amg.lib()
    assert "cuda" in amg.build_info(lib)
    assert amg.device_count(lib) >= 1, "no CUDA device: the product has no CPU path"
    return lib


@pytest.fixture(scope="module")
def O():
    return orc.Oracle()


def pmode(mode):
    return api.REDUCE_SEQUENTIAL if mode == orc.SEQ else api.REDUCE_TREE


CASES = [("dump", 0), ("poisson7", 6), ("poisson7", 13), ("poisson7", 20), ("poisson27", 7), ("poisson27", 10),
         ("aniso7", 10), ("aniso7", 14), ("sem_hex", 8), ("sem_hex", 12)]


@pytest.mark.parametrize("mode", [orc.SEQ, orc.TREE])
@pytest.mark.parametrize("name,n", CASES)
def test_hierarchy_bit_identical_to_oracle(L, O, name, n, mode):
    """C/F split, every sparsity pattern, every fp64 value, D, m, rho, ids: identical bits, and so
    is every traced intermediate array (coarsening rounds, skeletons, lambda, R, ...)."""
    mat = M.read_amgdmp(GOLDEN) if name == "dump" else M.by_name(name, n)
    h = O.setup_raw(*mat, mode, trace=True)
    want = O.fetch(h); twant = O.trace(); O.free(h)
    api.set_reduce_mode(pmode(mode), L=L)
    L.amgb_trace_enable(1)
    try:
        H = amg.amg_setup(*mat, L=L)
        tgot = product_trace(L)
    finally:
        L.amgb_trace_enable(0)
        api.set_reduce_mode(api.REDUCE_SEQUENTIAL, L=L)
    assert first_trace_mismatch(tgot, twant) is None
    assert orc.compare(fetch(H), want) == []
    assert H.timing()["launches"] > 0


@pytest.mark.parametrize("name,n", [("poisson7", 13), ("poisson7", 20), ("poisson27", 10), ("aniso7", 14), ("sem_hex", 12)])
def test_large_level_kernels_bit_identical_to_oracle(L, O, name, n, monkeypatch):
    """The kernels that take over on the large levels of a 128^3 setup -- the long-row SpMV through
    shared memory, block-per-column find_support, the Q builders that keep Q in HBM (one block, and
    the 8-CTA cluster kernel on a second stream), the optimistic block SpGEMM and its HBM overflow
    path -- never run on problems small enough for the oracle.  AMGB_TEST_SMALL_BINS=1 lowers their
    thresholds; every traced intermediate array and the hierarchy must still equal the oracle's."""
    monkeypatch.setenv("AMGB_TEST_SMALL_BINS", "1")
    mat = M.by_name(name, n)
    h = O.setup_raw(*mat, orc.SEQ, trace=True)
    want = O.fetch(h); twant = O.trace(); O.free(h)
    L.amgb_trace_enable(1)
    try:
        H = amg.amg_setup(*mat, L=L)
        tgot = product_trace(L)
    finally:
        L.amgb_trace_enable(0)
    assert first_trace_mismatch(tgot, twant) is None
    assert orc.compare(fetch(H), want) == []


def _trace_fixtures():
    import glob
    return sorted(os.path.basename(f) for f in glob.glob(os.path.join(GOLDEN, "trace_*.json.gz")))


@pytest.mark.parametrize("fname", _trace_fixtures())
def test_trace_fixture_parity(L, fname):
    """Direct oracle comparison far above the sizes the oracle finishes in seconds: the committed
    oracle traces (tests/golden/make_trace_fixtures.py; 64^3..128^3 7-point, 32^3/48^3 27-point,
    Q1 vertex meshes, anisotropic diffusion) hold the FNV-1a hash of every traced intermediate array
    of the setup -- every coarsening round, skeleton, lambda, R, W, AfP, every level's A -- plus the
    per-level sizes, m, rho and the hierarchy fingerprint.  The CUDA path must reproduce all of them
    with the production kernel selection (no test hooks): the large-level kernels, the transposed
    SpGEMM route and the many-block exact reduction are what runs at these sizes."""
    from util import load_trace_fixture
    fx = load_trace_fixture(os.path.join(GOLDEN, fname))
    mat = M.by_name(fx["workload"], fx["n"])
    L.amgb_trace_enable(1)
    try:
        H = amg.amg_setup(*mat, L=L)
        tgot = product_trace(L)
    finally:
        L.amgb_trace_enable(0)
    assert first_trace_mismatch(tgot, fx["trace"]) is None
    assert H.nlevels == fx["nlevels"] and H.nullspace == fx["nullspace"]
    for l, lev in enumerate(fx["levels"]):
        info = H.level_info(l)
        got = [info[k] for k in ("n", "nnz", "nf", "nc", "nnzf", "nnzw", "nnzfp", "coarsen_rounds", "lanczos_iters",
                                 "interp_rounds")]
        assert got == lev["info"], (l, got, lev["info"])
        if l < fx["nlevels"] - 1:
            par = H.level_params(l)
            assert par["m"] == lev["m"] and par["rho"] == float.fromhex(lev["rho"])
    assert "%016x" % H.hash() == fx["hierarchy_hash"]
    # and without tracing (lazy R0, no eager downloads): the same fingerprint
    H2 = amg.amg_setup(*mat, L=L)
    assert "%016x" % H2.hash() == fx["hierarchy_hash"]


@pytest.mark.parametrize("force", ["t", "o", "to", "b", "p"])
@pytest.mark.parametrize("name,n", [("poisson7", 13), ("poisson27", 10), ("aniso7", 14), ("sem_hex", 12)])
def test_forced_routes_bit_identical_to_oracle(L, O, name, n, force, monkeypatch):
    """Routes the size heuristics only take on the large levels, forced on oracle-sized inputs
    (AMGB_TEST_FORCE): 't' every SpGEMM through the transposed product, 'o' an arena that
    overflows so that the second pass runs (with AMGB_TEST_SMALL_BINS also the branch that rebuilds
    the overflow list), 'b' the many-block exact reduction inside a setup, 'p' the panel
    A-orthogonalisation for large supports.  Trace and hierarchy must equal the oracle's."""
    monkeypatch.setenv("AMGB_TEST_FORCE", force)
    if "o" in force:
        monkeypatch.setenv("AMGB_TEST_SMALL_BINS", "1")
    mat = M.by_name(name, n)
    h = O.setup_raw(*mat, orc.SEQ, trace=True)
    want = O.fetch(h); twant = O.trace(); O.free(h)
    L.amgb_trace_enable(1)
    try:
        H = amg.amg_setup(*mat, L=L)
        tgot = product_trace(L)
    finally:
        L.amgb_trace_enable(0)
    assert first_trace_mismatch(tgot, twant) is None
    assert orc.compare(fetch(H), want) == []
    assert H.hash() == orc.hierarchy_hash(want)


def test_unmodified_reference_driver_on_gpu(L, tmp_path):
    """The reference's serial_amg.c + fail.c, compiled unmodified with the reference's flags and
    linked against libamg_setup_b200.so (amg_setup / amg_export / free_data with the reference's
    signatures and its build-time uint): run on the bundled dump, its four files must equal the
    reference's byte for byte.  The binary is built where /root/reference exists (oracle/Makefile
    `drivers`) and travels in oracle/_ref/."""
    import shutil
    import subprocess
    exe = os.path.join(ROOT, "oracle", "_ref", "serial_amg_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/serial_amg_b200 was not built (needs /root/reference at build time)")
    for k in "ijp":
        shutil.copy(os.path.join(GOLDEN, "amgdmp_%s.dat" % k), str(tmp_path))
    r = subprocess.run([exe], cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:]
    z = np.load(os.path.join(GOLDEN, "ref_dump.npz"))
    for f in ("amg.dat", "amg_W.dat", "amg_AfP.dat", "amg_Aff.dat"):
        assert np.array_equal(np.fromfile(os.path.join(str(tmp_path), f)), z["file_" + f.replace(".", "_")]), f


@pytest.mark.parametrize("fixture", ["ref_dump", "ref_sem_hex4", "ref_sem_hex6", "ref_sem_hex_3x4x5"])
def test_identical_to_reference_fixtures(L, fixture):
    """Against struct amg_setup_data produced by the unmodified reference (tests/golden/
    make_golden.py): bit-exact C/F split and patterns; values within 1e-12 relative is the bar
    north_star sets -- they are in fact identical."""
    from test_oracle import load_golden
    want, mat, _ = load_golden(os.path.join(GOLDEN, fixture + ".npz"))
    got = fetch(amg.amg_setup(*mat, L=L))
    assert orc.compare(got, want, rtol=1e-12, params_rtol=1e-12) == []
    assert orc.compare(got, want) == []


def test_export_files_identical_to_reference(L, tmp_path):
    z = np.load(os.path.join(GOLDEN, "ref_dump.npz"))
    H = amg.amg_setup_from_dump(GOLDEN, L=L)
    H.export(str(tmp_path))
    for f in ("amg.dat", "amg_W.dat", "amg_AfP.dat", "amg_Aff.dat"):
        assert np.array_equal(np.fromfile(os.path.join(str(tmp_path), f)), z["file_" + f.replace(".", "_")]), f


def test_edge_cases(L, O):
    Ai, Aj, Av = M.poisson7(5, 4, 3)
    want = O.setup(Ai, Aj, Av, orc.SEQ)
    # explicit zeros, empty rows/cols, unsorted input
    Ai2 = np.concatenate([Ai, np.array([70, 90], np.int32)]); Aj2 = np.concatenate([Aj, np.array([71, 90], np.int32)])
    Av2 = np.concatenate([Av, [0.0, 0.0]])
    assert orc.compare(fetch(amg.amg_setup(Ai2, Aj2, Av2, L=L)), want) == []
    p = np.random.default_rng(1).permutation(len(Av))
    assert orc.compare(fetch(amg.amg_setup(Ai[p], Aj[p], Av[p], L=L)), want) == []
    one = fetch(amg.amg_setup(np.array([0], np.int32), np.array([0], np.int32), np.array([3.0]), L=L))
    assert one.nlevels == 1 and one.nullspace == 0
    two = (np.array([0, 0, 1, 1], np.int32), np.array([0, 1, 0, 1], np.int32), np.array([2.0, -1.0, -1.0, 2.0]))
    assert orc.compare(fetch(amg.amg_setup(*two, L=L)), O.setup(*two, orc.SEQ)) == []
    with pytest.raises(amg.AmgError, match="duplicate"):
        amg.amg_setup(np.array([0, 0, 1], np.int32), np.array([0, 0, 1], np.int32), np.ones(3), L=L)
    with pytest.raises(amg.AmgError):
        amg.amg_setup(np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0), L=L)


def test_vcycle_matches_oracle(L, O):
    for mat in (M.sem_hex(7), M.poisson7(10)):
        H = amg.amg_setup(*mat, L=L)
        h = O.setup_raw(*mat, orc.SEQ)
        n = H.level_info(0)["n"]
        b = np.random.default_rng(5).standard_normal(n)
        x = H.solve(b); xo = O.solve(h, b)
        O.free(h)
        assert np.abs(x - xo).max() <= 1e-12 * np.abs(xo).max()


def test_vcycle_identical_to_reference_fixture(L):
    """tests/golden/vcycle_ref.npz: x = crs_solve(b) computed by the REFERENCE's own amg_exec /
    crs_solve lines (amg.c:85-189, oracle/vcycle_ref_harness.c).  The CUDA V-cycle -- plain and
    replayed as a CUDA graph -- must give the same vector: bit for bit (every row sum, the
    restriction through W' and the mean in the reference's order), and in any case within the
    1e-12 relative of BASELINE.json's north_star."""
    fx = np.load(os.path.join(GOLDEN, "vcycle_ref.npz"))
    names = sorted(k[:-3] for k in fx.files if k.endswith("_Ai"))
    assert len(names) >= 7
    for name in names:
        H = amg.amg_setup(fx[name + "_Ai"], fx[name + "_Aj"], fx[name + "_Av"], L=L)
        want = fx[name + "_x"]
        for rep in range(3):            # first call plain, then captured, then replayed
            x = H.solve(fx[name + "_b"])
            assert np.abs(x - want).max() <= 1e-12 * np.abs(want).max(), (name, rep)
            assert np.array_equal(x, want), (name, rep, np.abs(x - want).max())
        H.free()


def test_crs_interface(L):
    Ai, Aj, Av = M.sem_hex(5)
    n = int(Ai.max()) + 1
    ids = np.arange(1, n + 1, dtype=np.uint64)
    d = amg.crs_setup(n, ids, len(Av), Ai, Aj, Av, 1, None, L=L)       # uint = unsigned long, the reference's build
    import scipy.sparse as sp
    A = sp.coo_matrix((Av, (Ai, Aj))).tocsr()
    b = np.random.default_rng(2).standard_normal(n); b -= b.mean()
    x = np.zeros(n); xk = np.zeros(n)
    for _ in range(10):
        r = b - A @ xk; r -= r.mean()
        amg.crs_solve(x, d, r)
        xk += x
    res = b - A @ xk
    assert np.linalg.norm(res - res.mean()) < 1e-4 * np.linalg.norm(b)
    amg.crs_stats(d)
    amg.crs_free(d)


def _csr(t):
    import scipy.sparse as sp
    ro, col, a, shape = t
    return sp.csr_matrix((a, col, ro), shape=shape)


@pytest.mark.parametrize("name,n", [("poisson7", int(os.environ.get("AMGB_TEST_FULL_N", "40"))), ("poisson27", 24),
                                     ("aniso7", 28), ("sem_hex", 20)])
def test_size_independent_properties(L, name, n):
    """At sizes the oracle no longer finishes in seconds (set AMGB_TEST_FULL_N=128 for the
    BASELINE size): the Galerkin identity A_{l+1} = P^t A_l P with P = [W; I] recomputed
    independently on the host, symmetry of every coarse operator, a proper C/F partition, and
    run-to-run determinism."""
    mat = M.by_name(name, n)
    H1 = amg.amg_setup(*mat, L=L)
    g = fetch(H1)
    for l in range(g.nlevels - 1):
        lev = g.levels[l]
        A = _csr(lev["A"]); W = _csr(lev["W"]); Anext = _csr(g.levels[l + 1]["A"])
        C = lev["C"] != 0
        assert C.sum() == Anext.shape[0] and (~C).sum() == W.shape[0]
        ids = np.concatenate([lev["idc"], lev["idf"]])
        assert len(np.unique(ids)) == A.shape[0]
        Aff = A[~C][:, ~C]; Afc = A[~C][:, C]; Acc = A[C][:, C]
        G = W.T @ (Aff @ W + Afc) + Afc.T @ W + Acc
        scale = abs(Anext).max()
        assert abs(G - Anext).max() <= 1e-12 * scale
        assert abs(Anext - Anext.T).max() <= 1e-12 * scale
        assert abs(_csr(lev["Af"]) - Aff).max() == 0.0
        assert abs(_csr(lev["AfP"]) - (Aff @ W + Afc)).max() <= 1e-12 * abs(Aff).max()
        assert (lev["D"] > 0).all() and 0 <= lev["rho"] < 1 and lev["m"] >= 1
    H2 = amg.amg_setup(*mat, L=L)
    assert orc.compare(fetch(H2), g) == []


def _seq(p):
    return float(np.add.accumulate(np.asarray(p, np.float64))[-1])     # strictly left to right


def test_sequential_reduction_kernel_is_exact(L):
    """The parallel left-to-right sum (runtime.cu: k_eps_dot) must equal the plain loop bit for
    bit on friendly and on adversarial inputs: ties, binade crossings, cancellation, wide ranges,
    sums hovering around zero, denormals, non-finite values."""
    rng = np.random.default_rng(0)
    cases = {
        "squares": rng.standard_normal(300001) ** 2,
        "signed": rng.standard_normal(100003),
        "wide_signed": rng.standard_normal(50000) * 10.0 ** rng.uniform(-8, 8, 50000),
        "wide_positive": np.abs(rng.standard_normal(200000)) * 10.0 ** rng.uniform(-12, 3, 200000),
        "quarter_ties": rng.integers(-8, 9, 60000) * 0.25,
        "ints_plus_ones": rng.integers(1, 1 << 20, 100000) * 2.0 ** -30 + 1.0 * (rng.random(100000) < 0.01),
        "half_ulp_ties": np.concatenate([[1.0], rng.integers(-3, 4, 100000) * 2.0 ** -53]),
        "ties_at_2p52": np.concatenate([[2.0 ** 52], rng.integers(-3, 4, 50000) * 0.5]),
        "cancel": np.concatenate([[1.0, -1.0], rng.standard_normal(1000)]),
        "tenth": np.full(1000000, 0.1),
        "palindrome": (lambda x: np.concatenate([x, -x[::-1]]))(rng.standard_normal(40000)),
        "denormals": rng.integers(-5, 6, 5000) * 5e-324,
        "tiny": np.array([3.0]), "empty_like": np.zeros(17),
        "all_zero_large": np.zeros(1 << 20),
        "zeros_then_values": np.concatenate([np.zeros(300000), rng.standard_normal(1000) ** 2]),
        "sparse_nonzeros": np.where(rng.random(400000) < 1e-3, rng.standard_normal(400000), 0.0),
        "negative_zero_start": np.concatenate([[-0.0, -0.0], np.zeros(10000), [-1.5, 2.5]]),
        "with_inf": np.concatenate([rng.standard_normal(100), [np.inf], rng.standard_normal(100)]),
        # long vectors through the segment planner: crossings in the middle of chunks, a sum that
        # shrinks again, one that changes sign, zeros between values, a late non-finite term
        "growing": np.abs(rng.standard_normal(1 << 20)) * np.linspace(1e-6, 1.0, 1 << 20) ** 3,
        "up_then_down": np.concatenate([np.abs(rng.standard_normal(300000)), -np.abs(rng.standard_normal(299000))]),
        "sign_change": np.concatenate([np.abs(rng.standard_normal(200000)), -2.0 * np.abs(rng.standard_normal(200000))]),
        "blocks_of_zeros": np.where((np.arange(600000) // 5000) % 2 == 0, rng.standard_normal(600000) ** 2, 0.0),
        "around_zero_long": rng.standard_normal(500000) * 1e-3,
        "late_inf": np.concatenate([rng.standard_normal(400000) ** 2, [np.inf], rng.standard_normal(1000)]),
        "ties_long": np.concatenate([[2.0 ** 30], rng.integers(-3, 4, 400000) * 2.0 ** -23]),
        "exact_powers": np.full(1 << 19, 2.0 ** -10),
    }
    for name, p in cases.items():
        got = api.debug_dot(p, None, api.REDUCE_SEQUENTIAL, L=L)
        want = _seq(p)
        assert got == want or (np.isnan(got) and np.isnan(want)), (name, got, want)
    a = rng.standard_normal(1 << 20); b = a * rng.uniform(0.5, 1.5, 1 << 20)
    assert api.debug_dot(a, b, api.REDUCE_SEQUENTIAL, L=L) == _seq(a * b)
    assert api.debug_dot(a, a, api.REDUCE_SEQUENTIAL, L=L) == _seq(a * a)


def test_two_rank_partitioned_setup_identical_to_one_gpu(L):
    """Row-partitioned setup over NCCL (DESIGN.md row e): on a box with >= 2 GPUs, two ranks build
    the hierarchy together (SpGEMM rows and local solves partitioned, every product however small)
    and each rank's result must equal the single-GPU hierarchy bit for bit (tools/dist_check.py)."""
    import subprocess
    import sys
    if amg.device_count(L) < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, AMGB_DIST_MIN_NNZ="0", MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29711", os.path.join(ROOT, "tools", "dist_check.py"), "poisson7:14",
           "aniso7:10", "sem_hex:8"]
    r = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "DIST ALL OK" in r.stdout, r.stdout[-3000:]


@pytest.mark.parametrize("small_bins", ["0", "1"])
def test_spgemm_primitive_every_bin(L, small_bins, monkeypatch):
    """mxm in isolation through the C ABI on operands whose rows fall into every SpGEMM tier (tiles,
    warps with 512/2048 slots and cp.async rings, blocks with 4096/8192 slots and bulk-async rings,
    HBM tables, global hash), with rows handed down the optimistic ladder: bit-identical to the
    exact-order reference, exact-zero drops included.  With the test hook on, every row with more
    than 96 products starts at the smallest warp tier with limits of 20/40/60/80 keys, so that most
    rows walk through every tier."""
    from util import spgemm_adversarial_operands, spgemm_reference, same_csr_bits
    monkeypatch.setenv("AMGB_TEST_SMALL_BINS", small_bins)
    A, B = spgemm_adversarial_operands(0)
    want = spgemm_reference(A, B)
    got = api.debug_spgemm(A, B, L=L)
    assert got[3] == want[3] and same_csr_bits(got, want)
    binned, final = api.debug_spgemm_tiers(L=L)
    assert all(n > 0 for n in final), ("every SpGEMM tier must have processed rows", binned, final)
    assert sum(final) > sum(binned)            # some rows were handed down the ladder

"""CPU tests of the oracle (test infrastructure): golden fixtures generated from the compiled
reference, the live reference when oracle/_ref is present, and the reduction tree."""
import ctypes
import glob
import os

import numpy as np
import pytest

from util import ROOT, orc
from omp_amg_b200 import matrices as M

GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def O():
    orc.build(ref=False)
    return orc.Oracle()


def load_golden(path):
    z = np.load(path)
    H = orc.Hierarchy()
    H.nlevels = int(z["nlevels"]); H.nullspace = int(z["nullspace"])
    for l in range(H.nlevels):
        lev = {}
        for k in ("A", "Af", "W", "AfP"):
            if "L%d_%s_ro" % (l, k) in z:
                lev[k] = (z["L%d_%s_ro" % (l, k)].astype(np.int64), z["L%d_%s_col" % (l, k)].astype(np.int64),
                          z["L%d_%s_a" % (l, k)], tuple(int(x) for x in z["L%d_%s_shape" % (l, k)]))
        for k in ("C", "D", "idc", "idf"):
            if "L%d_%s" % (l, k) in z:
                lev[k] = z["L%d_%s" % (l, k)]
        if "L%d_m" % l in z:
            lev["m"] = float(z["L%d_m" % l]); lev["rho"] = float(z["L%d_rho" % l])
        H.levels.append(lev)
    return H, (z["Ai"], z["Aj"], z["Av"]), z


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "ref_*.npz"))))
def test_oracle_bit_identical_to_reference_fixture(O, path):
    """Sequential-reduction mode reproduces the reference's struct amg_setup_data bit for bit:
    C/F split, every pattern, every value, D, m, rho, ids (fixtures: tests/golden/make_golden.py)."""
    want, mat, _ = load_golden(path)
    got = O.setup(*mat, orc.SEQ)
    assert orc.compare(got, want) == []


def test_golden_facts_of_all_ok_check(O):
    """The reference's own check file (all_ok.check) is stale against its code (it prints
    lam/P of an older revision), but its sizes still hold: 49 rows, 36 F points, 13 C points."""
    H = O.setup(*M.read_amgdmp(GOLDEN), orc.SEQ)
    assert H.levels[0]["A"][3] == (49, 49)
    assert H.levels[0]["W"][3] == (36, 13)
    assert [l["A"][3][0] for l in H.levels] == [49, 13, 4, 1]
    assert H.nullspace == 1


def test_export_files_match_reference(O, tmp_path):
    want, mat, z = load_golden(os.path.join(GOLDEN, "ref_dump.npz"))
    h = O.setup_raw(*mat, orc.SEQ)
    O.export(h, str(tmp_path))
    O.free(h)
    for f in ("amg.dat", "amg_W.dat", "amg_AfP.dat", "amg_Aff.dat"):
        got = np.fromfile(os.path.join(str(tmp_path), f))
        assert np.array_equal(got, z["file_" + f.replace(".", "_")]), f


def test_glibc_rand_stream(O):
    libc = ctypes.CDLL(None)
    libc.srand(1)
    want = np.array([libc.rand() for _ in range(2000)], dtype=np.int64)
    assert np.array_equal(O.rand_stream(2000, 1), want)


def test_tree_reduction_matches_definition(O):
    rng = np.random.default_rng(0)
    for n in (1, 5, 255, 256, 1023, 1024, 1025, 5000, 1024 * 1024 + 3):
        a = rng.standard_normal(n); b = rng.standard_normal(n)
        p = a * b

        def chunk(v):
            v = np.concatenate([v, np.zeros(1024 - len(v))])
            s = ((v[0:256] + v[256:512]) + v[512:768]) + v[768:1024]
            s = s.reshape(8, 32)
            for off in (16, 8, 4, 2, 1):
                s = s[:, :off] + s[:, off:2 * off]
            w = s[:, 0]
            for off in (4, 2, 1):
                w = w[:off] + w[off:2 * off]
            return w[0]

        def tree(v):
            if len(v) <= 1024:
                return chunk(v)
            return tree(np.array([chunk(v[i:i + 1024]) for i in range(0, len(v), 1024)]))

        assert O.dot(a, b, orc.TREE) == tree(p)
        assert O.dot(a, b, orc.SEQ) == float(np.add.reduce(p[:1])) if n == 1 else True


def test_reduction_order_sensitivity_is_real(O):
    """Why the product defaults to sequential reductions: the reference algorithm takes threshold
    decisions (exact-zero drops in mpm/mxm, find_support, coarsen ties) on values that depend on
    the last bit of its dot products.  With a tree-ordered dot the values agree to rounding, the
    hierarchy sizes agree here, but stored patterns differ in entries of relative size 1e-16."""
    mat = M.poisson7(8)
    a = O.setup(*mat, orc.SEQ)
    b = O.setup(*mat, orc.TREE)
    assert [l["A"][3][0] for l in a.levels] == [l["A"][3][0] for l in b.levels]
    for l in range(a.nlevels - 1):
        assert np.array_equal(a.levels[l]["C"], b.levels[l]["C"])
        assert np.array_equal(a.levels[l]["W"][1], b.levels[l]["W"][1])
        assert np.abs(a.levels[l]["W"][2] - b.levels[l]["W"][2]).max() <= 1e-12 * np.abs(a.levels[l]["W"][2]).max()
    assert len(a.levels[0]["AfP"][1]) != len(b.levels[0]["AfP"][1])


@pytest.mark.skipif(not orc.Ref.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("name,n", [("dump", 0), ("sem_hex", 5), ("sem_hex", 7), ("poisson7", 4)])
def test_oracle_vs_live_reference(O, name, n):
    """Against the unmodified reference compiled from /root/reference.  poisson7(4) is an input
    on which the reference's unchecked sp_add (amg_setup.c:1665) writes to wrong entries; the
    oracle's REFSCAN mode reproduces even that, bit for bit."""
    R = orc.Ref()
    mat = M.read_amgdmp(GOLDEN) if name == "dump" else M.by_name(name, n)
    O.L.amgo_set_spadd_mode(1)
    try:
        got = O.setup(*mat, orc.SEQ)
    finally:
        O.L.amgo_set_spadd_mode(0)
    want = R.setup(*mat)
    assert orc.compare(got, want) == []


@pytest.mark.skipif(not orc.Ref.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_single_stages_vs_live_reference(O):
    R = orc.Ref()
    Ai, Aj, Av = M.sem_hex(4)
    n = Ai.max() + 1
    A = O.from_csr(O.L.amgo_build_csr(len(Av), Ai, Aj, Av))
    rng = np.random.default_rng(1)
    vr = (rng.random(n) < 0.6).astype(np.float64); vc = (rng.random(n) < 0.5).astype(np.float64)
    for got, want in (
        (O.from_csr(O.L.amgo_transpose(O.to_csr(*A))), R.transpose(A)),
        (O.from_csr(O.L.amgo_spgemm(O.to_csr(*A), O.to_csr(*A))), R.mxm(A, A, 0.0)),
        (O.from_csr(O.L.amgo_mpm(2.0, O.to_csr(*A), -0.5, O.to_csr(*A))), R.mpm(2.0, A, -0.5, A)),
        (O.from_csr(O.L.amgo_sub_mat(O.to_csr(*A), vr, vc)), R.sub_mat(A, vr, vc)),
    ):
        assert got[3] == want[3]
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
    vc_o = np.zeros(n)
    O.L.amgo_coarsen(vc_o, O.to_csr(*A), 0.7)
    assert np.array_equal(vc_o, R.coarsen(A, 0.7))


def _vcycle_cases():
    fx = np.load(os.path.join(GOLDEN, "vcycle_ref.npz"))
    names = sorted(k[:-3] for k in fx.files if k.endswith("_Ai"))
    return fx, names


def test_oracle_vcycle_bit_identical_to_reference_fixture(O):
    """tests/golden/vcycle_ref.npz holds x = crs_solve(b) of the REFERENCE's own amg_exec/crs_solve
    (amg.c:85-189, compiled unchanged in oracle/vcycle_ref_harness.c; make_vcycle_golden.py).  The
    oracle's restatement of the V-cycle must reproduce every vector bit for bit."""
    fx, names = _vcycle_cases()
    assert len(names) >= 7
    for name in names:
        mat = (fx[name + "_Ai"], fx[name + "_Aj"], fx[name + "_Av"])
        h = O.setup_raw(*mat, orc.SEQ)
        x = O.solve(h, fx[name + "_b"])
        O.free(h)
        assert np.array_equal(x, fx[name + "_x"]), (name, np.abs(x - fx[name + "_x"]).max())


@pytest.mark.skipif(not orc.RefVcycle.available(), reason="oracle/_ref/libvcycle_ref.so not built (needs /root/reference)")
@pytest.mark.parametrize("name,n", [("dump", 0), ("sem_hex", 7), ("poisson7", 9), ("poisson27", 7), ("aniso7", 9)])
def test_oracle_vcycle_vs_live_reference_code(O, name, n):
    """The same comparison against the harness run live, on more inputs and right-hand sides."""
    V = orc.RefVcycle()
    mat = M.read_amgdmp(GOLDEN) if name == "dump" else M.by_name(name, n)
    h = O.setup_raw(*mat, orc.SEQ)
    H = O.fetch(h)
    nrow = H.levels[0]["A"][3][0]
    for seed in range(3):
        b = np.random.default_rng(seed).standard_normal(nrow) * 10.0 ** seed
        assert np.array_equal(O.solve(h, b), V.solve(H, b)), (name, seed)
    O.free(h)

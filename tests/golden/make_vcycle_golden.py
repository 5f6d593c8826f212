"""Generates tests/golden/vcycle_ref.npz: outputs of the REFERENCE's own V-cycle code
(amg_exec + crs_solve, amg.c:85-189, compiled unchanged inside oracle/vcycle_ref_harness.c by
oracle/Makefile) in this container.  The fixture travels to the GPU box, /root/reference does not.

    python tests/golden/make_vcycle_golden.py

Per case: the COO input, a seeded right-hand side b and x = crs_solve(b) of the reference on the
hierarchy the unmodified reference (cases "ref_*": oracle/_ref/libamg_ref.so) or, for inputs on
which the reference's setup is not memory-safe (DESIGN.md "sp_add"), the oracle builds (cases
"orc_*"; the oracle is pinned to the reference bit for bit on the former).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from omp_amg_b200 import matrices as M  # noqa: E402

CASES = {
    "ref_dump": ("ref", lambda: M.read_amgdmp(HERE)),
    "ref_sem_hex6": ("ref", lambda: M.sem_hex(6, seed=1)),
    "ref_sem_hex_3x4x5": ("ref", lambda: M.sem_hex(3, 4, 5, seed=3)),
    "orc_sem_hex5_dirichlet": ("orc", lambda: M.sem_hex(5, seed=2, neumann=False)),   # reference setup unsafe here (make_golden.py skips it)
    "orc_poisson7_12": ("orc", lambda: M.poisson7(12)),
    "orc_poisson27_10": ("orc", lambda: M.poisson27(10)),
    "orc_aniso7_12": ("orc", lambda: M.aniso7(12)),
    "orc_sem_hex9": ("orc", lambda: M.sem_hex(9)),
}


def main():
    R, O, V = orc.Ref(), orc.Oracle(), orc.RefVcycle()
    out = {}
    for name, (kind, gen) in CASES.items():
        Ai, Aj, Av = gen()
        H = R.setup(Ai, Aj, Av) if kind == "ref" else O.setup(Ai, Aj, Av, orc.SEQ)
        n = H.levels[0]["A"][3][0]
        b = np.random.default_rng(len(name)).standard_normal(n)
        x = V.solve(H, b)
        out[name + "_Ai"] = np.asarray(Ai, np.int32); out[name + "_Aj"] = np.asarray(Aj, np.int32)
        out[name + "_Av"] = np.asarray(Av, np.float64)
        out[name + "_b"] = b; out[name + "_x"] = x
        out[name + "_nullspace"] = np.array(H.nullspace)
        print("%s: n %d levels %d nullspace %d |x| %.6g" % (name, n, H.nlevels, H.nullspace, np.linalg.norm(x)))
    np.savez_compressed(os.path.join(HERE, "vcycle_ref.npz"), **out)


if __name__ == "__main__":
    main()

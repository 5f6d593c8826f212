"""Generates tests/golden/ref_*.npz by running the UNMODIFIED reference (oracle/_ref, compiled
from /root/reference by oracle/Makefile) in this container.  The fixtures travel to the GPU
box, /root/reference does not.

    python tests/golden/make_golden.py

Each file holds the reference's struct amg_setup_data for one input: per level A, Af, W, AfP
(row offsets, columns, values), C flags, D, idc, idf, m, rho, plus nlevels and nullspace, and
the four files amg_export() wrote (amg.dat, amg_W.dat, amg_AfP.dat, amg_Aff.dat).
Only inputs on which the reference stays memory-safe are used (see DESIGN.md "sp_add").
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from omp_amg_b200 import matrices as M  # noqa: E402

CASES = {
    "dump": lambda: M.read_amgdmp(HERE),                 # the reference's bundled amgdmp_*.dat
    "sem_hex4": lambda: M.sem_hex(4, seed=0),
    "sem_hex6": lambda: M.sem_hex(6, seed=1),
    "sem_hex5_dirichlet": lambda: M.sem_hex(5, seed=2, neumann=False),
    "sem_hex_3x4x5": lambda: M.sem_hex(3, 4, 5, seed=3),
}


def pack(H, Ai, Aj, Av, exported):
    d = {"nlevels": H.nlevels, "nullspace": H.nullspace, "Ai": Ai, "Aj": Aj, "Av": Av}
    for l, lev in enumerate(H.levels):
        for k in ("A", "Af", "W", "AfP"):
            if k in lev:
                ro, col, a, shape = lev[k]
                d["L%d_%s_ro" % (l, k)] = ro.astype(np.int32)
                d["L%d_%s_col" % (l, k)] = col.astype(np.int32)
                d["L%d_%s_a" % (l, k)] = a
                d["L%d_%s_shape" % (l, k)] = np.array(shape)
        for k in ("C", "D", "idc", "idf"):
            if k in lev:
                d["L%d_%s" % (l, k)] = lev[k]
        if "m" in lev:
            d["L%d_m" % l] = lev["m"]
            d["L%d_rho" % l] = lev["rho"]
    for name, arr in exported.items():
        d["file_" + name] = arr
    return d


def main():
    R = orc.Ref()
    O = orc.Oracle()
    for name, gen in CASES.items():
        Ai, Aj, Av = gen()
        # refuse inputs on which the reference leaves its arrays (REFSCAN reports it)
        O.L.amgo_set_spadd_mode(1)
        O.L.amgo_debug_reset()
        O.setup(Ai, Aj, Av, orc.SEQ)
        O.L.amgo_debug_spadd_miss.restype = __import__("ctypes").c_int64
        miss = O.L.amgo_debug_spadd_miss()
        O.L.amgo_set_spadd_mode(0)
        if miss:
            print("skip %s: reference sp_add misplaces %d updates" % (name, miss))
            continue
        H = R.setup(Ai, Aj, Av)
        exported = {}
        with tempfile.TemporaryDirectory() as td:
            R.export_last(td)
            for f in ("amg.dat", "amg_W.dat", "amg_AfP.dat", "amg_Aff.dat"):
                exported[f.replace(".", "_")] = np.fromfile(os.path.join(td, f))
        np.savez_compressed(os.path.join(HERE, "ref_%s.npz" % name), **pack(H, Ai, Aj, Av, exported))
        print("wrote ref_%s.npz: levels %s" % (name, [l["A"][3][0] for l in H.levels]))


if __name__ == "__main__":
    main()

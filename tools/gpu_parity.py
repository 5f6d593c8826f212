"""Dev helper (GPU box): product vs oracle(TREE) on a list of matrices, with stage traces."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from util import *
from omp_amg_b200 import matrices as M

def main():
    O = orc.Oracle()
    L = api.lib()
    print(api.build_info(L), "devices:", api.device_count(L), flush=True)
    cases = sys.argv[1:] or ["dump:0", "poisson7:6", "poisson7:12", "poisson27:7", "sem_hex:8", "aniso7:10", "poisson7:20"]
    nbad = 0
    for c in cases:
        timeonly = c.startswith("t:")
        if timeonly: c = c[2:]
        name, n = c.split(":"); n = int(n)
        mat = M.read_amgdmp(os.path.join(ROOT, "tests", "golden")) if name == "dump" else M.by_name(name, n)
        if timeonly:
            for rep in range(2):
                t = time.time(); Hp2 = api.amg_setup(*mat, L=L); dt = time.time() - t
                tm = Hp2.timing()
                print(c, "levels", [Hp2.level_info(l)["n"] for l in range(Hp2.nlevels)], flush=True)
                print("   time-only %.4fs" % dt, {k: (round(v, 5) if isinstance(v, float) else v) for k, v in tm.items()}, flush=True)
                Hp2.free()
            continue
        h = O.setup_raw(*mat, orc.TREE, trace=True); Ho = O.fetch(h); to = O.trace(); O.free(h)
        L.amgb_trace_enable(1)
        t = time.time(); Hp = api.amg_setup(*mat, L=L); dt = time.time() - t
        tp = product_trace(L)
        L.amgb_trace_enable(0)
        bad = orc.compare(fetch(Hp), Ho)
        mm = first_trace_mismatch(tp, to)
        nbad += bool(bad) or (mm is not None)
        print(c, [l["A"][3][0] for l in Ho.levels], "gpu %.3fs (traced)" % dt, "IDENTICAL" if not bad else bad[:6],
              "| trace", len(tp), len(to), "mismatch:", mm, flush=True)
        t = time.time(); Hp2 = api.amg_setup(*mat, L=L); dt = time.time() - t
        tm = Hp2.timing()
        print("   untraced %.4fs" % dt, {k: (round(v, 5) if isinstance(v, float) else v) for k, v in tm.items()}, flush=True)
    print("FAILED" if nbad else "ALL OK")
    return 1 if nbad else 0

if __name__ == "__main__":
    sys.exit(main())

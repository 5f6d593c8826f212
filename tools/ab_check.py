"""GPU box: the same setup with an environment switch off and on must give identical hierarchies.

    python tools/ab_check.py AMGB_SPGEMM_OPT8 poisson7 128 [first_level_compared]
The switch must be one the library reads at every call (not cached)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def main():
    import omp_amg_b200 as amg
    from omp_amg_b200 import api, matrices
    var, name, n = sys.argv[1], sys.argv[2], int(sys.argv[3])
    lo = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    L = amg.lib()
    api._check(L, L.amgb_init(0))
    mat = matrices.by_name(name, n)
    res = []
    for val in ("0", "1"):
        os.environ[var] = val
        amg.amg_setup(*mat, L=L).free()
        t0 = time.time(); H = amg.amg_setup(*mat, L=L); dt = time.time() - t0
        tm = H.timing()
        lev = []
        for l in range(H.nlevels):
            info = H.level_info(l)
            par = H.level_params(l)
            arrs = None
            if l >= lo:
                arrs = [H.csr(l, api.A)] + ([H.csr(l, w) for w in (api.W, api.AFP)] if l < H.nlevels - 1 else [])
            lev.append((info, par, arrs))
        res.append(lev)
        print("%s=%s: %.3f s  spgemm %.3f s  levels %s" % (var, val, dt, tm["spgemm_device_s"], [x[0]["n"] for x in lev]), flush=True)
        H.free()
    bad = 0
    for l, (a, b) in enumerate(zip(*res)):
        if a[0] != b[0] or a[1] != b[1]:
            print("level", l, "info/params differ", a[0], b[0]); bad += 1
        if a[2] is not None:
            for (r1, c1, v1, s1), (r2, c2, v2, s2) in zip(a[2], b[2]):
                if not (np.array_equal(r1, r2) and np.array_equal(c1, c2) and np.array_equal(v1.view(np.uint64), v2.view(np.uint64))):
                    print("level", l, "matrix differs", s1, s2); bad += 1
    print("AB FAILED" if bad or len(res[0]) != len(res[1]) else "AB IDENTICAL")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())

"""One setup of a large workload on one GPU, then V-cycles on the resident hierarchy.

    python tools/big_run.py sem_hex 158        # Q1 vertex mesh, 159^3 = 4.02 M rows (BASELINE config 5 size)
    python tools/big_run.py poisson27 128

Prints the level sizes, the setup time (CUDA events inside the library), the peak device memory,
the hierarchy fingerprint, and the time of a V-cycle (amg_exec, replayed as a CUDA graph).  Sizes
that do not fit end with the library's error -102 (out of device memory), not with a dead box.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    import omp_amg_b200 as amg
    from omp_amg_b200 import api, matrices
    name, n = sys.argv[1], int(sys.argv[2])
    cycles = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    L = amg.lib()
    api._check(L, L.amgb_init(0))
    t0 = time.time()
    mat = matrices.by_name(name, n)
    rows, nnz = int(mat[0].max()) + 1, len(mat[2])
    print("generated %s %d: %d rows, %d entries in %.1f s" % (name, n, rows, nnz, time.time() - t0), flush=True)
    out = {"workload": "%s_%d" % (name, n), "rows": rows, "nnz": nnz}
    try:
        t0 = time.time()
        H = amg.amg_setup(*mat, L=L)
        wall = time.time() - t0
    except amg.AmgError as e:
        out["error"] = str(e)
        out["peak_device_bytes"] = int(L.amgb_peak_device_bytes())
        print(json.dumps(out), flush=True)
        return 1
    t = H.timing()
    out.update({"levels": [H.level_info(l)["n"] for l in range(H.nlevels)], "setup_wall_s": wall,
                "setup_device_s": t["device_total_s"], "stage_s": {k: t[k] for k in ("coarsen", "lanczos", "interp", "galerkin")},
                "spgemm_device_s": t["spgemm_device_s"], "launches": int(t["launches"]),
                "peak_device_bytes": int(L.amgb_peak_device_bytes()), "hierarchy_hash": "%016x" % H.hash()})
    print(json.dumps(out), flush=True)
    n0 = H.level_info(0)["n"]
    b = torch.sin(torch.arange(n0, dtype=torch.float64, device="cuda") * 0.37) + 0.1
    x = torch.zeros(n0, dtype=torch.float64, device="cuda")
    H.solve_device_repeat(x.data_ptr(), b.data_ptr(), 3)
    torch.cuda.synchronize()
    t0 = time.time()
    H.solve_device_repeat(x.data_ptr(), b.data_ptr(), cycles)
    torch.cuda.synchronize()
    cyc = (time.time() - t0) / cycles
    out.update({"vcycle_ms": cyc * 1e3, "vcycles_per_s": 1.0 / cyc, "x_norm": float(torch.linalg.norm(x))})
    print(json.dumps(out), flush=True)
    H.free()
    return 0


if __name__ == "__main__":
    sys.exit(main())

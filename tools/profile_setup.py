"""Run under ncu with --profile-from-start off: profiles exactly ONE warm setup.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/profile_setup.py poisson7 128
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import omp_amg_b200 as amg
    from omp_amg_b200 import api, matrices
    name, n = sys.argv[1], int(sys.argv[2])
    warm = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    L = amg.lib()
    api._check(L, L.amgb_init(0))
    mat = matrices.by_name(name, n)
    for _ in range(warm):
        amg.amg_setup(*mat, L=L).free()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    H = amg.amg_setup(*mat, L=L)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    t = H.timing()
    print("profiled setup: %s %d levels %d launches %d total %.3fs" % (name, n, H.nlevels, t["launches"], t["total"]))
    H.free()


if __name__ == "__main__":
    main()

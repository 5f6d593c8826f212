"""GPU box, N >= 2 ranks under torchrun: the row-partitioned setup must produce, on every rank,
the hierarchy one GPU builds, and the row-partitioned V-cycle the solution one GPU computes -- bit
for bit.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29700 tools/dist_check.py poisson7:24 aniso7:12 sem_hex:10 t:poisson7:64
Cases prefixed with t: also print the single-GPU and partitioned times.
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import numpy as np
    import torch
    import torch.distributed as dist
    from util import api, amg, fetch, orc
    from omp_amg_b200 import matrices as M
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = amg.lib()
    api._check(L, L.amgb_init(local))
    cases = sys.argv[1:] or ["poisson7:16", "aniso7:10", "sem_hex:8"]
    mats = []
    for c in cases:
        timed = c.startswith("t:")
        name, n = c[2:].split(":") if timed else c.split(":")
        mats.append((c, timed, M.by_name(name, int(n))))
    single = []
    for c, timed, mat in mats:                      # one GPU, before joining
        if timed:
            amg.amg_setup(*mat, L=L).free()
        t0 = time.time(); H = amg.amg_setup(*mat, L=L); dt = time.time() - t0
        n0 = H.level_info(0)["n"]
        b = np.sin(np.arange(n0) * 0.37) + 0.1
        x1 = H.solve(b)
        tv = 0.0
        if timed:
            H.solve(b); torch.cuda.synchronize()
            t0 = time.time()
            for _ in range(10):
                H.solve(b)
            tv = (time.time() - t0) / 10
        single.append((fetch(H), dt, H.timing(), b, x1, tv))
        H.free()
    api.comm_init(L)
    nbad = 0
    for (c, timed, mat), (want, dt1, tm1, b, x1, tv1) in zip(mats, single):
        if timed:
            amg.amg_setup(*mat, L=L).free()
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.time(); H = amg.amg_setup(*mat, L=L); dtp = time.time() - t0
        tm = H.timing()
        bad = orc.compare(fetch(H), want)
        x = H.solve(b)                               # the ranks apply one cycle together (solve.cu: spmv_dist)
        if not np.array_equal(x, x1):
            bad.append("V-cycle differs from one GPU's by %g" % np.abs(x - x1).max())
        tvp = 0.0
        if timed:
            H.solve(b); dist.barrier(); torch.cuda.synchronize()
            t0 = time.time()
            for _ in range(10):
                H.solve(b)
            tvp = (time.time() - t0) / 10
        # partitioned storage for the solve phase: the same cycle from this rank's row blocks only
        freed = H.partition_solve_storage()
        x2 = H.solve(b)
        if not np.array_equal(x2, x1):
            bad.append("V-cycle on partitioned storage differs from one GPU's by %g" % np.abs(x2 - x1).max())
        print("rank %d %-16s solve-phase storage partitioned: %.1f MB released on this rank, V-cycle %s"
              % (rank, c, freed / 1e6, "IDENTICAL" if np.array_equal(x2, x1) else "DIFFERS"), flush=True)
        H.free()
        flag = torch.tensor([1 if bad else 0], device="cuda")
        dist.all_reduce(flag)
        nbad += int(flag[0])
        print("rank %d %-16s levels %s  single %.3fs  x%d %.3fs  comm: %d exchanges %.1f MB %.4fs | spgemm %.3fs -> %.3fs | "
              "V-cycle through host vectors %.2f ms -> %.2f ms  %s"
              % (rank, c, [l["A"][3][0] for l in want.levels], dt1, world, dtp, tm["comm_calls"], tm["comm_bytes"] / 1e6,
                 tm["comm_device_s"], tm1["spgemm_device_s"], tm["spgemm_device_s"], tv1 * 1e3, tvp * 1e3,
                 "IDENTICAL (hierarchy and V-cycle)" if not bad else bad[:4]), flush=True)
    api.comm_finalize(L)
    dist.destroy_process_group()
    if rank == 0:
        print("DIST FAILED" if nbad else "DIST ALL OK", flush=True)
    return 1 if nbad else 0


if __name__ == "__main__":
    sys.exit(main())

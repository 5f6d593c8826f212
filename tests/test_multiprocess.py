"""N > 1 launch contract on CPU (gloo, world_size 2).  The setup path does not shard in this round
("replicas only", DESIGN.md row e): every rank runs its own setup and only the timing is reduced.
What can be tested without a GPU is the launch protocol of bench.py's reference arm (rank 0
prints one JSON line, the other ranks exit 0 without work) and the max-over-ranks reduction."""
import json
import os
import subprocess
import sys

import pytest

from util import ROOT


def test_reference_arm_under_torchrun_world2():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29631", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "1", "--warmup", "0", "--size", "16", "--sample-n", "8"]
    r = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "amg_setup_time" and d["unit"] == "s"
    assert d["higher_is_better"] is False and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the reduction bench.py applies to its per-rank step times: MAX over ranks after a barrier
    t = torch.tensor([1.0 + rank, 10.0 - rank], dtype=torch.float64)
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    q.put((rank, t.tolist()))
    dist.destroy_process_group()


def test_max_over_ranks_reduction_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, 29632, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    assert out[0][1] == [2.0, 10.0] and out[1][1] == [2.0, 10.0]


# ---------------------------------------------------------------------------------------
# Row-partitioned setup and V-cycle on 2, 3 and 4 ranks (DESIGN.md row e).  The host-emulation build
# runs the same partitioning code as the product (spgemm_partitioned, q_partition, spmv_dist) with a
# gloo transport in place of NCCL; every rank must end with the hierarchy -- and the solution of a
# V-cycle -- one rank builds, bit for bit.
# ---------------------------------------------------------------------------------------
def _dist_worker(rank, world, port, q, case, dist_spmv="0"):
    import numpy as np
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      AMGB_DIST_MIN_NNZ="0",     # partition every product, however small
                      AMGB_DIST_SPMV=dist_spmv)  # "1" (the default): also the matrix-vector products of the setup loops
    from util import EMU_SO, api, amg, fetch, orc
    from omp_amg_b200 import matrices
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L = api.lib(EMU_SO)
        mat = {"poisson7_8": lambda: matrices.poisson7(8), "aniso_6": lambda: matrices.aniso7(6),
               "amgdmp": lambda: matrices.read_amgdmp(os.path.join(ROOT, "tests", "golden"))}[case]()
        H1 = amg.amg_setup(*mat, L=L)                        # before joining: one rank, no exchange
        single = fetch(H1)
        n = single.levels[0]["A"][3][0]
        b = np.random.default_rng(7).standard_normal(n)
        x1 = H1.solve(b)
        api.comm_init_host_gloo(L)
        assert L.amgb_comm_size() == world and L.amgb_comm_rank() == rank
        H = amg.amg_setup(*mat, L=L)
        t = H.timing()
        got = fetch(H)
        bad = orc.compare(got, single)
        # the V-cycle on the ranks together: row blocks of every matrix-vector product, the result
        # vectors exchanged -- the same bits as one rank's cycle, on every rank
        import ctypes
        c0, c1, nb = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        L.amgb_comm_stats(ctypes.byref(c0), ctypes.byref(nb))
        x = H.solve(b)
        L.amgb_comm_stats(ctypes.byref(c1), ctypes.byref(nb))
        if not np.array_equal(x, x1):
            bad.append("partitioned V-cycle differs from the single-rank one by %g" % np.abs(x - x1).max())
        if c1.value <= c0.value:
            bad.append("the V-cycle exchanged nothing")
        # partitioned STORAGE for the solve phase: every rank keeps its row blocks of W', W, AfP, Af
        # and releases the rest; the cycle must not change by a bit, and the hierarchy is solve-only
        freed = H.partition_solve_storage()
        if freed <= 0:
            bad.append("partition_solve_storage released nothing")
        x2 = H.solve(b)
        if not np.array_equal(x2, x1):
            bad.append("V-cycle on partitioned storage differs by %g" % np.abs(x2 - x1).max())
        x3, x3ref = H.solve(2.5 * b), H1.solve(2.5 * b)      # H1: whole storage (collective as well)
        if not np.array_equal(x3, x3ref):
            bad.append("second V-cycle on partitioned storage differs by %g" % np.abs(x3 - x3ref).max())
        try:
            H.csr(0, api.W)
            bad.append("accessor of a partitioned matrix did not refuse")
        except amg.AmgError:
            pass
        try:
            H.hash()
            bad.append("fingerprint of a solve-only hierarchy did not refuse")
        except amg.AmgError:
            pass
        api.comm_finalize(L)
        q.put((rank, bad[:3], int(t["comm_calls"]), int(t["comm_bytes"])))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case,world,port,dist_spmv",
                         [("poisson7_8", 2, 29641, "0"), ("aniso_6", 3, 29642, "0"), ("amgdmp", 2, 29643, "0"),
                          ("amgdmp", 4, 29644, "0"),     # 49 rows over 4 ranks: blocks of 1-4 rows on the coarse levels
                          ("poisson7_8", 3, 29646, "1"), ("amgdmp", 2, 29647, "1"),    # setup SpMVs partitioned too
                          ("poisson7_8", 8, 29648, "1")])                             # the rank count of one B200 box
def test_row_partitioned_setup_matches_single_rank(case, world, port, dist_spmv):
    import torch.multiprocessing as mp
    subprocess.run(["make", "-j4", "-C", os.path.join(ROOT, "omp_amg_b200", "csrc"), "emu"], check=True,
                   stdout=subprocess.DEVNULL)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_dist_worker, args=(r, world, port, q, case, dist_spmv)) for r in range(world)]
    for p in ps:
        p.start()
    out = sorted(q.get(timeout=600) for _ in ps)
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    for rank, bad, calls, nbytes in out:
        assert not bad, "rank %d differs from the single-rank hierarchy: %s" % (rank, bad)
        assert calls > 0 and nbytes > 0, "rank %d exchanged nothing: the stages were not partitioned" % rank


def _threshold_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    os.environ.pop("AMGB_DIST_MIN_NNZ", None)          # default threshold: 2^20 entries
    from util import EMU_SO, api, amg
    from omp_amg_b200 import matrices
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L = api.lib(EMU_SO)
        api.comm_init_host_gloo(L)
        H = amg.amg_setup(*matrices.poisson7(6), L=L)
        q.put((rank, int(H.timing()["comm_calls"])))
        api.comm_finalize(L)
    finally:
        dist.destroy_process_group()


def test_small_products_stay_replicated_gloo_world2():
    """Below the work threshold (AMGB_DIST_MIN_NNZ, default 2^20 entries) a stage is not worth an
    exchange: every rank computes it whole and nothing is sent."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_threshold_worker, args=(r, 2, 29645, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = sorted(q.get(timeout=300) for _ in ps)
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    assert out == [(0, 0), (1, 0)]

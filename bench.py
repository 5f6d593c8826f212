#!/usr/bin/env python
"""bench.py -- AMG hierarchy setup time on B200 (BASELINE.json's headline metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload poisson7 --size 128]

One "step" = one full hierarchy setup (amg_setup, amg_setup.c:60) of the workload.
  value   seconds per setup with the COO input already resident in HBM (device CUDA events
          around the whole setup, max over ranks)
  e2e     the same through the host-buffer entry point amgb_setup(): pinned host COO in,
          H2D inside the timed region, per-level metadata read back (D2H)
  roofline  the SpGEMM kernels (the dominant HBM-bound kernel family): algorithmic bytes /
          device time of those launches, both measured live inside the timed steps
  cpu_baseline  the CPU port of the reference (oracle/, sequential reductions) on a bounded
          sample of the same workload, scaled linearly in rows to the full size
N > 1 (one process per GPU, torchrun): the ranks build ONE hierarchy together -- the SpGEMM rows
and the local solves of the coarse columns are partitioned over the ranks and the blocks exchanged
through the library's NCCL communicator (DESIGN.md row e); every rank ends with the same,
bit-identical hierarchy.  Total work is fixed, so scaling is "strong" and value is the time of that
one setup (max over ranks).  --parallel replicas runs N independent setups instead.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def load_peaks():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.path = "/tmp/amgb_clocks_%d.csv" % os.getpid()
        self.p = None

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in open(self.path):
            c = [x.strip() for x in line.split(",")]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0])); mx.append(float(c[1]))
            except ValueError:
                continue
            for nm, v in zip(names, c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm)}
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


def workload_matrix(name, n):
    from omp_amg_b200 import matrices
    return matrices.by_name(name, n)


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path on the host cores.  oracle/_ref (the
    unmodified reference) cannot run this workload: its mxm is O(rows^2) and its unchecked
    sp_add leaves its arrays on finite-difference Poisson matrices (DESIGN.md "sp_add"), so the
    CPU port in oracle/ is timed, on a bounded sample, single-threaded like the reference."""
    if rank != 0:
        return
    from oracle import oracle as orc
    orc.build(ref=False)
    O = orc.Oracle()
    ns = args.sample_n
    mat = workload_matrix(args.workload, ns)
    rows_s = int(mat[0].max()) + 1
    rows_full = args.n ** 3 if args.workload != "sem_hex" else (args.n + 1) ** 3
    for _ in range(min(args.warmup, 1)):
        O.setup(*workload_matrix(args.workload, max(8, ns // 3)), orc.SEQ)
    ts = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        h = O.setup_raw(*mat, orc.SEQ)
        ts.append(time.perf_counter() - t0)
        O.free(h)
    t = sum(ts) / len(ts)
    scaled = t * rows_full / rows_s
    sample = ("%s %d^3 (%d rows) measured %.3f s per setup; scaled linearly in rows to %d rows "
              "(flatters the CPU: its cost grows faster than linearly)" % (args.workload, ns, rows_s, t, rows_full))
    line = {"impl": "reference", "metric": "amg_setup_time", "value": scaled, "unit": "s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": False,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s_%d^3_full_hierarchy_setup" % (args.workload, args.n), "rows": rows_full,
                       "sample_rows": rows_s},
            "cpu_baseline": {"value": scaled, "unit": "s", "cores": 1, "kind": "port", "sample": sample},
            "e2e": {"value": scaled, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="poisson7")
    ap.add_argument("--size", dest="n", type=int, default=128, help="grid points per side")
    ap.add_argument("--sample-n", type=int, default=40, help="grid size of the CPU baseline's bounded sample")
    ap.add_argument("--reduce", default="seq", choices=["seq", "tree"],
                    help="seq: reference-order dot products (bit-identical hierarchy); tree: fast mode")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--parallel", default="partitioned", choices=["partitioned", "replicas"],
                    help="N > 1: partitioned = one setup, stages row-partitioned over the ranks (NCCL); "
                         "replicas = N independent setups")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import omp_amg_b200 as amg
    from omp_amg_b200 import api

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; omp_amg_b200 has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own messages (its version banner at
        # NCCL_DEBUG=WARN) go to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = amg.lib()
    api._check(L, L.amgb_init(local))
    api.set_reduce_mode(api.REDUCE_SEQUENTIAL if args.reduce == "seq" else api.REDUCE_TREE, L=L)
    partitioned = False
    if world > 1 and args.parallel == "partitioned":
        ok = 1
        try:
            api.comm_init(L)
        except Exception as e:          # no NCCL for the library: say so and run replicas
            sys.stderr.write("rank %d: comm_init failed (%s); running replicas\n" % (rank, e))
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        partitioned = bool(int(flag[0]))
        if not partitioned:
            api.comm_finalize(L)

    Ai, Aj, Av = workload_matrix(args.workload, args.n)
    nnz = len(Av)
    rows = int(Ai.max()) + 1
    # pinned host copies (e2e input) and device copies (value input)
    hAi = torch.from_numpy(Ai).pin_memory(); hAj = torch.from_numpy(Aj).pin_memory(); hAv = torch.from_numpy(Av).pin_memory()
    dAi = hAi.cuda(non_blocking=True); dAj = hAj.cuda(non_blocking=True); dAv = hAv.cuda(non_blocking=True)
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step_device():
        H = api.amg_setup(dAi.data_ptr(), dAj.data_ptr(), dAv.data_ptr(), L=L, device_ptrs=True, nnz=nnz)
        t = H.timing()
        H.free()
        return t

    def step_e2e():
        H = api.amg_setup(hAi.numpy(), hAj.numpy(), hAv.numpy(), L=L)
        meta = [(H.level_info(l), H.level_params(l)) for l in range(H.nlevels)]
        H.free()
        return meta

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    tims = [step_device() for _ in range(args.steps)]
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    dev_s = sum(t["device_total_s"] for t in tims)
    tt = torch.tensor([dev_s, wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dev_s, wall = float(tt[0]), float(tt[1])
    per_step = dev_s / args.steps

    # end to end through the host-buffer entry point
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    metas = [step_e2e() for _ in range(args.steps)]
    barrier()
    e2e_wall = time.perf_counter() - t0
    te = torch.tensor([e2e_wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_step = float(te[0]) / args.steps
    nlev = len(metas[-1])
    h2d = nnz * 16
    d2h = nlev * (10 * 8 + 4 * 8)

    if rank == 0:
        peak, peak_src = load_peaks()
        sp_s = sum(t["spgemm_device_s"] for t in tims)
        sp_b = sum(t["spgemm_bytes"] for t in tims)
        sp_n = sum(t["spgemm_calls"] for t in tims)
        achieved = (sp_b / sp_s / 1e9) if sp_s > 0 else 0.0
        last = tims[-1]
        line = {
            "metric": "amg_setup_time", "value": per_step, "unit": "s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": False,
            "scaling": "strong" if partitioned else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s_%d^3_full_hierarchy_setup" % (args.workload, args.n), "rows": rows, "nnz": nnz,
                       "levels": nlev, "reduce_mode": args.reduce, "l2": "inputs_exceed_l2" if nnz * 16 > 126e6 else "small_input",
                       "parallelism": ("row_partitioned_spgemm_and_local_solves_x%d" % world) if partitioned
                       else ("replicas_x%d" % world if world > 1 else "single_gpu")},
            "rows_per_s": (1 if partitioned else world) * rows / per_step,
            "comm": {"exchanges_per_step": int(last["comm_calls"]), "bytes_received_per_rank_per_step": int(last["comm_bytes"]),
                     "device_s_per_step": last["comm_device_s"], "transport": "nccl grouped broadcast (all-gather of row blocks)"},
            "wall_s_per_step": wall / args.steps,
            "stage_s": {k: last[k] for k in ("build_csr", "coarsen", "smoother", "lanczos", "interp", "galerkin")},
            "gpu_launches": int(sum(t["launches"] for t in tims)),
            "host_syncs": int(sum(t["syncs"] for t in tims)),
            "clocks": clocks,
            "e2e": {"value": e2e_step, "unit": "s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "roofline": {"bound": "hbm", "kernel": "spgemm (fused hash/dense SpGEMM family, all launches of the timed steps)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
                         "peak_source": peak_src,
                         # DRAM bytes (read + write) of the longest launch of the family, ncu, per launch
                         "traffic": 3449320704,
                         "traffic_note": "k_spgemm_bitmap, grid 274882 (the level-1 Galerkin product W'(Af W + Afc): "
                                         "10.8M x 35.9M -> 53.5M entries): 66.7 ms, 2849 MB read + 600 MB written of DRAM "
                                         "against 1.20 GB algorithmic, L2 hit 56%, 12% active warps -- "
                                         "profiles/r1_ncu_spgemm_family_dram_traffic_poisson7_128.txt",
                         "launch_seconds": sp_s / max(sp_n, 1),
                         "algorithmic_bytes_per_call": sp_b / max(sp_n, 1), "calls": int(sp_n),
                         "share_of_step": sp_s / dev_s if dev_s else None},
        }
        if world == 1 and not args.no_cpu_baseline:
            from oracle import oracle as orc
            orc.build(ref=False)
            O = orc.Oracle()
            ns = args.sample_n
            smat = workload_matrix(args.workload, ns)
            srows = int(smat[0].max()) + 1
            t0 = time.perf_counter()
            h = O.setup_raw(*smat, orc.SEQ)
            ts = time.perf_counter() - t0
            O.free(h)
            line["cpu_baseline"] = {
                "value": ts * rows / srows, "unit": "s", "cores": 1, "kind": "port",
                "sample": "%s %d^3 (%d rows): %.3f s measured, scaled linearly in rows to %d rows (flatters the CPU)"
                          % (args.workload, ns, srows, ts, rows)}
        print(json.dumps(line), flush=True)
    if world > 1:
        if partitioned:
            api.comm_finalize(L)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

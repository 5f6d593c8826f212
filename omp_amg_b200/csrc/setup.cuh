// setup.cuh -- the AMG hierarchy in HBM and the stages that build it.
#pragma once
#include "sparse.cuh"
#include "hostmath.h"

namespace amgb {

// One level of struct amg_setup_data (amg_tools.h:29), resident in HBM.
struct Level {
  Csr A;              // data->A[l]
  Csr Af, W, AfP;     // data->Af[l], data->W[l], data->AfP[l]   (absent on the last level)
  Csr Wt;             // W^t, kept for the restriction b_{l+1} += W^t b_l of the V-cycle
  Buf<double> C;      // data->C[l]: 1.0 where the dof stays on the coarse level
  Buf<double> D;      // data->D[l]: diagonal smoother of the F block
  Buf<int> idc, idf;  // data->idc[l], data->idf[l] (1-based ids of the finest level)
  Buf<int> fpos, cpos;  // position of every dof among the F (resp. C) dofs of its level
  int n = 0, nf = 0, nc = 0;
  double m = 0, rho = 0, lmin = 0, lmax = 0;
  int coarsen_rounds = 0, lanczos_k = 0, interp_rounds = 0;
  // V-cycle workspaces (bf, bc, xc, xf, t, c1, c2, r), allocated on the first solve and kept:
  // amg.c keeps b, x, c, c_old, r, buf in struct crs_data for the same reason
  mutable Buf<double> ws[8];
};

struct StageTimes {    // host wall-clock with a stream sync at stage ends, seconds
  double build = 0, coarsen = 0, smoother = 0, lanczos = 0, interp = 0, galerkin = 0, total = 0;
  double device_total = 0;   // CUDA-event time from the first to the last kernel of the setup
  double spgemm = 0;   // device time (CUDA events) inside the SpGEMM kernels
  i64 spgemm_bytes = 0;  // algorithmic bytes moved by those kernels
  i64 spgemm_calls = 0;
  double spmv = 0;     // the same for the long-row SpMV kernels, when spmv_stats_enable(true) (sparse.cuh)
  i64 spmv_bytes = 0, spmv_calls = 0;
  i64 comm_calls = 0, comm_bytes = 0;   // exchanges of the row-partitioned stages (comm.cuh)
  double comm = 0;                      // device seconds inside them
};

// one V-cycle captured as a CUDA graph (solve.cu): the coarse levels are launch-bound, a replay
// costs one launch instead of ~15 per level
struct SolveGraph {
  void *exec = nullptr;           // cudaGraphExec_t
  const double *b = nullptr;
  double *x = nullptr;
  i64 kernels = 0;                // kernels inside the graph
  int calls = 0;                  // plain (uncaptured) solves so far
  SolveGraph() {}
  SolveGraph(const SolveGraph &) = delete;
  SolveGraph &operator=(const SolveGraph &) = delete;
  SolveGraph(SolveGraph &&o) noexcept : exec(o.exec), b(o.b), x(o.x), kernels(o.kernels), calls(o.calls) { o.exec = nullptr; }
  ~SolveGraph();
};

struct Hierarchy {
  std::vector<Level> lv;
  int nullspace = 0;
  int n0 = 0;
  StageTimes t;
  i64 launches = 0, syncs = 0;
  // amgb_partition_solve_storage: the matrices of the V-cycle hold only this rank's row blocks;
  // the hierarchy then serves amgb_solve / crs_amg_solve only (accessors, export and the
  // fingerprint refuse)
  bool solve_only = false;
  mutable Buf<double> mean_scratch;     // device scalar of the mean projection
  // crs_solve's storage order (amg.c:438-446: unknowns sorted by the level at which they become
  // F, ascending inside a level): position of every top-level unknown, and the vector in that
  // order -- the mean of amg.c:182 is summed in it.  Built at the first projection.
  mutable Buf<int> lsort_pos;
  mutable Buf<double> lsort_x;
  mutable SolveGraph graph;
};

// stages (each cites amg_setup.c)
int coarsen(double *vc, const Csr &A, double ctol);                                   // :2737
int lanczos(double *lambda, const Csr &A, hostmath::GlibcRand &rng, int *iters);      // :2435
int pcg(double *x, const Csr &A, double *r, const double *M, double tol, const double *b);  // :2242
Csr interpolation(const Csr &Af, const Csr &Ac, const Csr &Ar, double gamma2, double tol,
                  int *rounds);                                                        // :598
void setup(i64 nnz, const int *dAi, const int *dAj, const double *dAv, Hierarchy &H);  // :60

// V-cycle (amg.c:114 amg_exec + amg.c:171 crs_solve), device vectors of length n0
void vcycle_solve(const Hierarchy &H, double *x, const double *b);
// the same through a CUDA graph captured on the second call with the same vectors (AMGB_SOLVE_GRAPH=0: never)
void vcycle_solve_graph(const Hierarchy &H, double *x, const double *b);
// Several ranks: every rank keeps only its row blocks of the matrices the V-cycle applies row-
// partitioned (Wt, W, AfP, Af of the large levels) and releases the rest; returns the bytes this
// rank released.  Collective; the hierarchy becomes solve-only.
i64 partition_solve_storage(Hierarchy &H);
// x -= mean(x) (amg.c:181-184): the sum runs over crs_solve's level-sorted storage order and
// stays on the device (no host round trip)
void project_mean(const Hierarchy &H, double *x);

}  // namespace amgb

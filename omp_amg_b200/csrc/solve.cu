// solve.cu -- the V-cycle of amg.c:114 (amg_exec) and the projection of amg.c:171 (crs_solve)
// on the hierarchy in HBM.  amg.c keeps all levels' F rows in one vector; here every level has
// its own numbering, which is the same arithmetic up to the column order inside rows of W/AfP.
#include "setup.cuh"
#include "comm.cuh"

namespace amgb {

// Several ranks (one process per GPU, every rank holding the hierarchy and the same b): the
// matrix-vector products of the large levels are row-partitioned (sparse.cu: spmv_vals inside a
// SpmvPartitionScope) -- a cycle streams every large matrix once across the box instead of once
// per GPU, and is bit-identical to one GPU's.  The vector updates and the small levels stay
// replicated: a vector of a large level is a few MB, an exchange costs more than the update.
void vcycle_level(const Hierarchy &H, int l, double *x, const double *b) {
  const Level &L = H.lv[(size_t)l];
  if (l == (int)H.lv.size() - 1) {
    const double *a = L.A.a.p;
    const int ns = H.nullspace;
    const bool has = L.A.nnz > 0;
    parallel_for(1, [=] DEV(i64) { const double d = (ns || !has) ? 0. : 1. / a[0]; x[0] = d * b[0]; });
    return;
  }
  const int n = L.n, nf = L.nf, nc = L.nc;
  if (!L.ws[0].p) {
    const int sz[8] = {nf, nc, nc, nf, nc, nf, nf, nf};
    for (int k = 0; k < 8; k++) L.ws[k].alloc(sz[k]);
  }
  Buf<double> &bf = L.ws[0], &bc = L.ws[1], &xc = L.ws[2], &xf = L.ws[3], &t = L.ws[4], &c1 = L.ws[5], &c2 = L.ws[6], &r = L.ws[7];
  const double *Cf = L.C.p;
  const int *fpos = L.fpos.p, *cpos = L.cpos.p;
  double *bfp = bf.p, *bcp = bc.p, *xcp = xc.p, *xfp = xf.p, *tp = t.p;
  parallel_for(n, [=] DEV(i64 i) { if (Cf[i] != 0.) bcp[cpos[i]] = b[i]; else bfp[fpos[i]] = b[i]; });
  // b_{l+1} += W^t b_l
  spmv(tp, 0, nullptr, 1, L.Wt, bfp);
  parallel_for(nc, [=] DEV(i64 i) { bcp[i] = 1 * bcp[i] + 1 * tp[i]; });
  vcycle_level(H, l + 1, xcp, bcp);
  // x_l = W x_{l+1};  b_l -= AfP x_{l+1}
  spmv(xfp, 0, bfp, 1, L.W, xcp);
  spmv(bfp, 1, bfp, -1, L.AfP, xcp);
  const double *d = L.D.p;
  double *c = c1.p, *co = c2.p, *rp = r.p;
  const unsigned m = (unsigned)L.m;
  double alpha = 0, beta = 0, gamma = 0;
  parallel_for(nf, [=] DEV(i64 i) { c[i] = d[i] * bfp[i]; });
  if (m > 1) {
    alpha = L.rho / 2; alpha *= alpha;
    gamma = 2 * alpha / (1 - 2 * alpha); beta = 1 + gamma;
    spmv(rp, 1, bfp, -1, L.Af, c);
    { double *s = c; c = co; co = s; }
    { double *cc = c, *cco = co; const double bt = beta;
      parallel_for(nf, [=] DEV(i64 i) { cc[i] = bt * (cco[i] + d[i] * rp[i]); }); }
  }
  for (unsigned ci = 3; ci <= m; ci++) {
    gamma = alpha * beta; gamma = gamma / (1 - gamma); beta = 1 + gamma;
    spmv(rp, 1, bfp, -1, L.Af, c);
    { double *s = c; c = co; co = s; }
    { double *cc = c, *cco = co; const double bt = beta, gm = gamma;
      parallel_for(nf, [=] DEV(i64 i) { cc[i] = bt * (cco[i] + d[i] * rp[i]) - gm * cc[i]; }); }
  }
  { double *cc = c;
    parallel_for(n, [=] DEV(i64 i) {
      if (Cf[i] != 0.) x[i] = xcp[cpos[i]];
      else { const int f = fpos[i]; x[i] = xfp[f] + cc[f]; }
    }); }
}

i64 partition_solve_storage(Hierarchy &H) {
  if (!comm_active()) return 0;
  SpmvPartitionScope partition_;
  const int r = comm_rank();
  i64 freed = 0;
  for (Level &L : H.lv)
    for (Csr *M : {&L.Wt, &L.W, &L.AfP, &L.Af})
      if (M->rn > 0 && spmv_is_partitioned(*M)) {
        freed += csr_keep_row_block(*M, (int)row_split(M->rn, r), (int)row_split(M->rn, r + 1));
        H.solve_only = true;
      }
  if (H.solve_only) dev_release_cache();       // hand the released blocks back to the driver
  return freed;
}

void vcycle_solve(const Hierarchy &H, double *x, const double *b) {
  SpmvPartitionScope partition_;
  const int n = H.n0;
  vcycle_level(H, 0, x, b);
  (void)n;
  if (H.nullspace) project_mean(H, x);
}

SolveGraph::~SolveGraph() {
#ifndef AMGB_EMU
  if (exec) cudaGraphExecDestroy((cudaGraphExec_t)exec);
#endif
}

void vcycle_solve_graph(const Hierarchy &H, double *x, const double *b) {
#ifdef AMGB_EMU
  vcycle_solve(H, x, b);
#else
  static int on = -1;
  if (on < 0) { const char *e = getenv("AMGB_SOLVE_GRAPH"); on = (e && *e == '0') ? 0 : 1; }
  SolveGraph &g = H.graph;
  Context &c = ctx();
  // the first solve runs plainly: it allocates the per-level workspaces and warms the allocator's
  // cache, so that the capture below makes no driver call besides the launches
  // several ranks with at least one partitioned product: the exchanges are NCCL calls; the cycle
  // runs uncaptured
  bool exchanges = false;
  if (comm_active()) {
    SpmvPartitionScope partition_;
    for (const Level &L : H.lv)
      for (const Csr *M : {&L.Wt, &L.W, &L.AfP, &L.Af})
        if (spmv_is_partitioned(*M)) exchanges = true;
  }
  if (!on || g.calls < 1 || exchanges) { vcycle_solve(H, x, b); g.calls++; return; }
  if (!g.exec || g.x != x || g.b != b) {
    if (g.exec) { cudaGraphExecDestroy((cudaGraphExec_t)g.exec); g.exec = nullptr; }
    const i64 l0 = c.launches;
    cudaGraph_t graph = nullptr;
    CUDA_CHECK(cudaStreamBeginCapture(c.stream, cudaStreamCaptureModeThreadLocal));
    bool ok = true;
    try { vcycle_solve(H, x, b); }
    catch (...) { ok = false; }
    cudaError_t ce = cudaStreamEndCapture(c.stream, &graph);
    if (!ok || ce != cudaSuccess || !graph) {
      // something inside needed the driver (an allocation the cache could not serve): this
      // hierarchy keeps the plain path
      cudaGetLastError();
      if (graph) cudaGraphDestroy(graph);
      c.launches = l0;
      g.calls = -(1 << 30);
      vcycle_solve(H, x, b);
      return;
    }
    g.kernels = c.launches - l0;
    c.launches = l0;
    cudaGraphExec_t ex = nullptr;
    cudaError_t e = cudaGraphInstantiate(&ex, graph, 0);
    cudaGraphDestroy(graph);
    CUDA_CHECK(e);
    g.exec = ex; g.x = x; g.b = b;
  }
  CUDA_CHECK(cudaGraphLaunch((cudaGraphExec_t)g.exec, c.stream));
  c.launches += g.kernels;
#endif
}

// x -= (1/n) * sum(x) (amg.c:181-184).  The reference sums ux in ITS storage order -- the
// unknowns sorted by the level at which they become F, ascending inside a level, the last level's
// unknown at the end (amg.c:438-446) -- so x is first gathered into that order; the sum (in the
// reduction order the context asks for) stays on the device: a solve never waits for the host.
static void build_level_sorted_positions(const Hierarchy &H) {
  const int nl = (int)H.lv.size();
  std::vector<int> off((size_t)nl + 1, 0);
  for (int l = 0; l < nl; l++) off[(size_t)l + 1] = off[(size_t)l] + (l < nl - 1 ? H.lv[(size_t)l].nf : H.lv[(size_t)l].n);
  Buf<int> g((i64)H.lv[(size_t)nl - 1].n + 1);
  { int *gp = g.p; const int o = off[(size_t)nl - 1]; parallel_for(H.lv[(size_t)nl - 1].n, [=] DEV(i64 i) { gp[i] = o + (int)i; }); }
  for (int l = nl - 2; l >= 0; l--) {
    const Level &L = H.lv[(size_t)l];
    Buf<int> gl((i64)L.n + 1);
    int *glp = gl.p;
    const int *gn = g.p, *cpos = L.cpos.p, *fpos = L.fpos.p;
    const double *Cf = L.C.p;
    const int o = off[(size_t)l];
    parallel_for(L.n, [=] DEV(i64 i) { glp[i] = (Cf[i] != 0.) ? gn[cpos[i]] : o + fpos[i]; });
    g = std::move(gl);
  }
  H.lsort_pos = std::move(g);
  H.lsort_x.alloc(H.n0);
}
void project_mean(const Hierarchy &H, double *x) {
  const i64 n = H.n0;
  if (n <= 0) return;
  if (!H.mean_scratch.p) H.mean_scratch.alloc(1);
  if (!H.lsort_pos.p) build_level_sorted_positions(H);
  const int *pos = H.lsort_pos.p;
  double *ux = H.lsort_x.p;
  parallel_for(n, [=] DEV(i64 i) { ux[pos[i]] = x[i]; });
  vsum_dev(H.mean_scratch.p, ux, n);
  const double *sp = H.mean_scratch.p;
  const double inv = 1 / (double)n;
  parallel_for(n, [=] DEV(i64 i) { const double avg = inv * sp[0]; x[i] = x[i] - avg; });
}

}  // namespace amgb

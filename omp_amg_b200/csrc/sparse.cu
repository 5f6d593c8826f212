// sparse.cu -- sparse primitives (first generation: one logical thread per row; the hot ones
// have warp-cooperative replacements in spgemm.cu / spmv kernels below).
#include "sparse.cuh"
#include "comm.cuh"

namespace amgb {

// ---- SpMV statistics (opt-in: spmv_stats_enable; bench.py's second roofline) ----
// device time (CUDA events around every call's launches) and algorithmic bytes of the kernels
// that apply a matrix with more than 24 entries per row (the k_spmv_pipe / k_spmv_tile family)
namespace {
struct SpmvStats {
  bool on = false;
  i64 bytes = 0, calls = 0;
#ifndef AMGB_EMU
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
#endif
};
SpmvStats g_spmv_stats;
}  // namespace
void spmv_stats_enable(bool on) { g_spmv_stats.on = on; }
void spmv_stats_reset() {
#ifndef AMGB_EMU
  for (auto &e : g_spmv_stats.ev) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
  g_spmv_stats.ev.clear();
#endif
  g_spmv_stats.bytes = 0; g_spmv_stats.calls = 0;
}
void spmv_stats_get(double *seconds, i64 *bytes, i64 *calls) {
  double sec = 0;
#ifndef AMGB_EMU
  for (auto &e : g_spmv_stats.ev) { float ms = 0; if (cudaEventElapsedTime(&ms, e.first, e.second) == cudaSuccess) sec += ms * 1e-3; }
#endif
  *seconds = sec; *bytes = g_spmv_stats.bytes; *calls = g_spmv_stats.calls;
}

void fill(double *p, i64 n, double v) { parallel_for(n, [=] DEV(i64 i) { p[i] = v; }); }
void fill_int(int *p, i64 n, int v) { parallel_for(n, [=] DEV(i64 i) { p[i] = v; }); }

void trace_csr(const char *tag, const Csr &A) {
  if (!ctx().trace_on) return;
  std::string t(tag);
  trace_dev((t + ".ro").c_str(), A.ro.p, sizeof(int) * (size_t)(A.rn + 1));
  trace_dev((t + ".col").c_str(), A.col.p, sizeof(int) * (size_t)A.nnz);
  trace_dev((t + ".a").c_str(), A.a.p, sizeof(double) * (size_t)A.nnz);
}

// ---------------------------------------------------------------------------------------
// SpMV: row sums strictly left to right, separate multiply and add (amg_tools.c:71)
// ---------------------------------------------------------------------------------------
#ifndef AMGB_EMU
// G threads per row: the products of G consecutive entries are formed side by side (coalesced
// loads), then added to the running sum one after the other in entry order, so the row sum has
// the reference's left-to-right rounding.  Every thread of the group carries the same sum.
template <int G>
__global__ void __launch_bounds__(256) k_spmv_tile(int rn, const int *ro, const int *col, const double *vals,
                                                   const double *x, double *z, double alpha, const double *y,
                                                   double beta, bool plain, const double *post, int longrow,
                                                   const int *gen, int want) {
  const int i = blockIdx.x * (256 / G) + threadIdx.x / G;
  if (i >= rn) return;
  if (ro[i + 1] - ro[i] > longrow) return;          // k_spmv_chain
  if (gen && gen[i] != want) return;                // row unchanged since its last sum (find_support)
  const int lane = threadIdx.x % G;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
  const int end = ro[i + 1];
  double t = 0;
  for (int base = ro[i]; base < end; base += G) {
    const int j = base + lane;
    const double p = (j < end) ? (x ? vals[j] * x[col[j]] : vals[j] * 1.0) : 0.0;
    const int m = end - base;
    if (m >= G) {                          // full batch: straight line of G shuffles and adds
#pragma unroll
      for (int l = 0; l < G; l++) t = t + __shfl_sync(gmask, p, l, G);
    } else {
      for (int l = 0; l < m; l++) t = t + __shfl_sync(gmask, p, l, G);
    }
  }
  if (lane == 0) {
    double r = plain ? beta * t : alpha * y[i] + beta * t;
    if (post) r = r * post[i];
    z[i] = r;
  }
}
// Long rows: one warp per row as above, but the 32 products of a batch travel through shared
// memory instead of shuffles.  A 64-bit shuffle is two SHFL instructions and the SM issues one per
// clock, which capped k_spmv_tile<32> at 1.37 TB/s whatever the matrix (per-call log, 128^3);
// here every lane parks its product (one STS) and reads the batch back as 16 broadcast LDS.128,
// then adds the 32 values in entry order.
__global__ void __launch_bounds__(256) k_spmv_row32(int rn, const int *ro, const int *col, const double *vals,
                                                    const double *x, double *z, double alpha, const double *y,
                                                    double beta, bool plain, const double *post, int longrow,
                                                    const int *gen, int want) {
  __shared__ __align__(16) double buf[8][2][32];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + w;
  if (i >= rn) return;
  if (ro[i + 1] - ro[i] > longrow) return;          // k_spmv_chain
  if (gen && gen[i] != want) return;
  const int end = ro[i + 1];
  double t = 0;
  int par = 0;
  for (int base = ro[i]; base < end; base += 32, par ^= 1) {
    const int j = base + lane;
    const double p = (j < end) ? (x ? vals[j] * x[col[j]] : vals[j] * 1.0) : 0.0;
    double *b = buf[w][par];
    b[lane] = p;
    __syncwarp();
    const int m = end - base;
    if (m >= 32) {
#pragma unroll
      for (int l = 0; l < 32; l += 2) {
        const double2 q = *reinterpret_cast<const double2 *>(b + l);
        t = t + q.x;
        t = t + q.y;
      }
    } else {
      for (int l = 0; l < m; l++) t = t + b[l];
    }
  }
  if (lane == 0) {
    double r = plain ? beta * t : alpha * y[i] + beta * t;
    if (post) r = r * post[i];
    z[i] = r;
  }
}
// Long rows, second generation: G lanes per row (32/G rows per warp, so an add instruction of the
// ordered chain serves several rows), U entries per lane and stage.  The products of a stage
// (G*U of them) are parked in a double-buffered slice of shared memory and read back as
// broadcast LDS.128; the loads are software-pipelined three stages deep: while stage s is added,
// x[col] of stage s+1 and (col, val) of stage s+2 are in flight, so the chain of dependent adds
// never waits for a gather.  On the coarse levels (4000 rows of 3600 entries) the run time is the
// chain of the longest row; everything else hides behind it.
template <int G, int U>
__global__ void __launch_bounds__(256) k_spmv_pipe(int rn, const int *ro, const int *col, const double *vals,
                                                   const double *x, double *z, double alpha, const double *y,
                                                   double beta, bool plain, const double *post, int longrow,
                                                   const int *gen, int want) {
  constexpr int S = G * U;
  __shared__ __align__(16) double buf[256 / G][2][S];
  const int grp = threadIdx.x / G, lane = threadIdx.x % G;
  const int i = blockIdx.x * (256 / G) + grp;
  if (i >= rn) return;
  if (ro[i + 1] - ro[i] > longrow) return;          // k_spmv_chain
  if (gen && gen[i] != want) return;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
  const int beg = ro[i], end = ro[i + 1];
  int c1[U], c2[U];
  double v0[U], x0[U], v1[U], v2[U];
#pragma unroll
  for (int u = 0; u < U; u++) {
    const int j1 = beg + u * G + lane, j2 = j1 + S;
    c1[u] = 0; v1[u] = 0.0; c2[u] = 0; v2[u] = 0.0;
    if (j1 < end) { if (x) c1[u] = col[j1]; v1[u] = vals[j1]; }
    if (j2 < end) { if (x) c2[u] = col[j2]; v2[u] = vals[j2]; }
  }
#pragma unroll
  for (int u = 0; u < U; u++) { x0[u] = (beg + u * G + lane < end) ? (x ? x[c1[u]] : 1.0) : 0.0; v0[u] = v1[u]; }
  double t = 0;
  int par = 0;
  for (int base = beg; base < end; base += S, par ^= 1) {
    double *b = buf[grp][par];
#pragma unroll
    for (int u = 0; u < U; u++) b[u * G + lane] = (base + u * G + lane < end) ? v0[u] * x0[u] : 0.0;
    // advance the pipeline: gather x for the next stage (its columns are here), fetch the stage after
#pragma unroll
    for (int u = 0; u < U; u++) {
      x0[u] = (base + S + u * G + lane < end) ? (x ? x[c2[u]] : 1.0) : 0.0;
      v0[u] = v2[u];
      const int j = base + 2 * S + u * G + lane;
      if (j < end) { if (x) c2[u] = col[j]; v2[u] = vals[j]; }
    }
    __syncwarp(gmask);
    const int m = end - base;
    if (m >= S) {
#pragma unroll
      for (int l = 0; l < S; l += 2) {
        const double2 q = *reinterpret_cast<const double2 *>(b + l);
        t = t + q.x;
        t = t + q.y;
      }
    } else {
      for (int l = 0; l < m; l++) t = t + b[l];
    }
  }
  if (lane == 0) {
    double r = plain ? beta * t : alpha * y[i] + beta * t;
    if (post) r = r * post[i];
    z[i] = r;
  }
}
// Very long rows: one block per row.  Warps 1..7 form the products of the next chunk of CH entries
// (coalesced loads, gathers in flight from 224 threads) while thread 0 adds the current chunk from
// shared memory in entry order, eight values ahead in registers: the run time is the chain of
// dependent additions itself (about 8 cycles per entry), not the memory latency per batch that a
// group-per-row kernel pays when only a handful of such rows exist.
#define SPMV_CH 2048
__global__ void __launch_bounds__(256) k_spmv_chain(const int *rows, const int *nrows, const int *ro, const int *col,
                                                    const double *vals, const double *x, double *z, double alpha,
                                                    const double *y, double beta, bool plain, const double *post,
                                                    int r0, int r1, int shift) {
  __shared__ double buf[2][SPMV_CH];
  const int t = threadIdx.x;
  const int n = *nrows;
  for (int q = blockIdx.x; q < n; q += gridDim.x) {
    const int i = rows[q];
    if (i < r0 || i >= r1) continue;              // another rank's row block (block-uniform)
    const int beg = ro[i - shift], end = ro[i - shift + 1];    // ro starts at row `shift` (Csr::partial)
    const int nch = (end - beg + SPMV_CH - 1) / SPMV_CH;
    __syncthreads();
    for (int j = beg + t; j < end && j < beg + SPMV_CH; j += 256) buf[0][j - beg] = x ? vals[j] * x[col[j]] : vals[j] * 1.0;
    double r = 0;
    for (int c = 0; c < nch; c++) {
      __syncthreads();
      const int cur = c & 1;
      if (t >= 32) {
        const int b0 = beg + (c + 1) * SPMV_CH;
        for (int j = b0 + (t - 32); j < end && j < b0 + SPMV_CH; j += 224) buf[cur ^ 1][j - b0] = x ? vals[j] * x[col[j]] : vals[j] * 1.0;
      } else if (t == 0) {
        const int b0 = beg + c * SPMV_CH;
        const int len = min(SPMV_CH, end - b0);
        const double *p = buf[cur];
        int j = 0;
        if (len >= 8) {
          double a0 = p[0], a1 = p[1], a2 = p[2], a3 = p[3], a4 = p[4], a5 = p[5], a6 = p[6], a7 = p[7];
          for (; j + 16 <= len; j += 8) {
            const double b0v = p[j + 8], b1 = p[j + 9], b2 = p[j + 10], b3 = p[j + 11], b4 = p[j + 12], b5 = p[j + 13],
                         b6 = p[j + 14], b7 = p[j + 15];
            r = r + a0; r = r + a1; r = r + a2; r = r + a3; r = r + a4; r = r + a5; r = r + a6; r = r + a7;
            a0 = b0v; a1 = b1; a2 = b2; a3 = b3; a4 = b4; a5 = b5; a6 = b6; a7 = b7;
          }
          r = r + a0; r = r + a1; r = r + a2; r = r + a3; r = r + a4; r = r + a5; r = r + a6; r = r + a7;
          j += 8;
        }
        for (; j < len; j++) r = r + p[j];
      }
    }
    if (t == 0) {
      double v = plain ? beta * r : alpha * y[i] + beta * r;
      if (post) v = v * post[i];
      z[i] = v;
    }
  }
}
__global__ void __launch_bounds__(256) k_find_long_rows(int rn, const int *ro, int longrow, int cap, int *rows, int *count,
                                                        int shift) {       // ro describes rows shift .. shift + rn
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= rn) return;
  if (ro[i + 1] - ro[i] > longrow) { const int p = atomicAdd(count, 1); if (p < cap) rows[p] = i + shift; }
}
#endif

static void spmv_vals_run(double *z, double alpha, const double *y, double beta, const Csr &M, const double *vals,
                          const double *x, const double *post, const int *gen, int want, int r0, int r1);
void spmv_rows(double *z, double alpha, const double *y, double beta, const Csr &M, const double *x, int r0, int r1) {
  if (r1 > r0) spmv_vals_run(z, alpha, y, beta, M, M.a.p, x, nullptr, nullptr, 0, r0, r1);
}
// Several ranks: a product with a large matrix is ROW-PARTITIONED -- rank r forms rows
// [n r/P, n (r+1)/P) with the single-GPU kernels and the blocks of z are exchanged in place
// (comm_allgatherv).  Every row sum is formed by exactly one rank in the single-GPU order, so z is
// bit-identical on every rank.  An exchange costs 15-35 us of NCCL latency, which a row block
// must save first.  Measured on 2 B200, one 128^3 setup (2.38 s with replicated products): 2.33 s
// with every matrix of 2^23 entries or more partitioned (5 300 more exchanges), 2.28 s from
// 20 M entries on (1 900 more) -- hence the default threshold (AMGB_DIST_MIN_NNZ_SOLVE;
// AMGB_DIST_MIN_NNZ when only that is set, as in the tests).  AMGB_DIST_SPMV=0 keeps the
// products of the setup loops replicated; inside the V-cycle (SpmvPartitionScope, solve.cu) the
// partition is always on.
static int g_spmv_partition = -1;      // -1: env not read yet; bit 0 = setup default, bit 1 = forced on by a scope
i64 spmv_partition_min_nnz() {
  static i64 v = -1;
  if (v < 0) {
    const char *e = getenv("AMGB_DIST_MIN_NNZ_SOLVE"), *g = getenv("AMGB_DIST_MIN_NNZ");
    v = e ? atoll(e) : g ? atoll(g) : (i64)20000000;
  }
  return v;
}
static int spmv_partition_state() {
  if (g_spmv_partition < 0) { const char *e = getenv("AMGB_DIST_SPMV"); g_spmv_partition = (e && *e == '0') ? 0 : 1; }
  return g_spmv_partition;
}
SpmvPartitionScope::SpmvPartitionScope() { prev = spmv_partition_state(); g_spmv_partition = prev | 2; }
SpmvPartitionScope::~SpmvPartitionScope() { g_spmv_partition = prev; }
bool spmv_is_partitioned(const Csr &M) {
  return comm_active() && spmv_partition_state() != 0 && M.rn >= comm_size() && M.nnz >= spmv_partition_min_nnz();
}

void spmv_vals(double *z, double alpha, const double *y, double beta, const Csr &M, const double *vals,
               const double *x, const double *post, const int *gen, int want) {
  if (!gen && spmv_is_partitioned(M)) {
    const int P = comm_size(), r = comm_rank();
    const int r0 = (int)row_split(M.rn, r), r1 = (int)row_split(M.rn, r + 1);
    if (r1 > r0) spmv_vals_run(z, alpha, y, beta, M, vals, x, post, nullptr, 0, r0, r1);
    std::vector<i64> off((size_t)P + 1);
    for (int q = 0; q <= P; q++) off[(size_t)q] = (i64)sizeof(double) * row_split(M.rn, q);
    comm_allgatherv(z, off.data(), "comm.spmv", false);
    return;
  }
  if (M.partial) throw Error(-120, "matrix storage is partitioned for the solve phase: the product needs the ranks it was partitioned over");
#ifndef AMGB_EMU
  static int logit = -1;
  if (logit < 0) { const char *e = getenv("AMGB_SPMV_LOG"); logit = (e && *e && *e != '0') ? 1 : 0; }
  if (logit) {                      // per-call device time (diagnostics; synchronises)
    Context &c = ctx();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, c.stream);
    spmv_vals_run(z, alpha, y, beta, M, vals, x, post, gen, want, 0, M.rn);
    cudaEventRecord(e1, c.stream);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    fprintf(stderr, "spmv %d x %d nnz %lld  %.4f ms  %.1f GB/s\n", M.rn, M.cn, (long long)M.nnz, ms,
            (12.0 * M.nnz + 16.0 * M.rn) / (ms * 1e-3) / 1e9);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return;
  }
#endif
  spmv_vals_run(z, alpha, y, beta, M, vals, x, post, gen, want, 0, M.rn);
}
// rows [r0, r1) only (the V-cycle's row blocks, solve.cu): the kernels see the block as a matrix
// of its own -- row offsets, z, y, post and gen shifted by r0; col/vals are indexed absolutely
static void spmv_vals_run(double *zf, double alpha, const double *yf, double beta, const Csr &M, const double *vals,
                          const double *x, const double *postf, const int *genf, int want, int r0, int r1) {
  StageTimer st_("prim.spmv");
  // a matrix whose storage is partitioned holds rows [row_lo, row_lo + own_rows) only
  const int shift = M.partial ? M.row_lo : 0;
  if (M.partial && (r0 < M.row_lo || r1 > M.row_lo + M.own_rows))
    throw Error(-120, "matrix storage is partitioned for the solve phase: rows outside this rank's block were asked for");
  const int *ro = M.ro.p + (r0 - shift), *col = M.col.p;
  double *z = zf + r0;
  const double *y = yf ? yf + r0 : nullptr, *post = postf ? postf + r0 : nullptr;
  const int *gen = genf ? genf + r0 : nullptr;
  const int rn = r1 - r0;
  const bool plain = (alpha == 0. || y == nullptr);
  int longrow = 0x7fffffff;
#ifndef AMGB_EMU
  Context &c = ctx();
  // rows of more than LR entries go to the block-per-row chain kernel; which rows those are is
  // found once per matrix (one read-back), so matrices without any pay nothing afterwards
  const int LR = test_small_bins() ? 40 : 1536;
  static int chain_on = -1;
  if (chain_on < 0) { const char *e = getenv("AMGB_SPMV_CHAIN"); chain_on = (e && *e == '0') ? 0 : 1; }
  if (M.n_long < 0) {
    M.n_long = 0;
    if (chain_on && M.rn > 0 && M.nnz > LR) {
      const int cap = M.rn < 16384 ? M.rn : 16384;
      M.long_rows.alloc((i64)cap + 1);
      dev_memset(M.long_rows.p + cap, 0, sizeof(int));
      const int have = M.partial ? M.own_rows : M.rn;       // rows whose offsets are stored
      if (have > 0) k_find_long_rows<<<(have + 255) / 256, 256, 0, c.stream>>>(have, M.ro.p, LR, cap, M.long_rows.p, M.long_rows.p + cap, shift);
      c.launches++; post_launch("find_long_rows");
      int cnt = 0;
      d2h(&cnt, M.long_rows.p + cap, sizeof(int));
      // many long rows: the matrix is long rows throughout, which the group-per-row kernels
      // stream at full rate anyway (a block per row pays off when a handful of rows set the time)
      static int chain_max = -1;
      if (chain_max < 0) { const char *e = getenv("AMGB_SPMV_CHAIN_MAX"); chain_max = e ? atoi(e) : 1024; }
      if (cnt > 0 && cnt <= cap && cnt <= chain_max) M.n_long = cnt;
      else M.long_rows.release();
    }
  }
  if (M.n_long > 0) longrow = LR;
  static double t1 = -1;       // rows up to this average length: one thread per row (AMGB_SPMV_T1)
  if (t1 < 0) { const char *e = getenv("AMGB_SPMV_T1"); t1 = e ? atof(e) : 24.0; }
  bool done = false;
  cudaEvent_t se0 = nullptr, se1 = nullptr;
  const bool stat = g_spmv_stats.on && rn == M.rn && rn > 0 && (double)M.nnz / (double)M.rn > t1 && !gen;
  if (stat) {
    if (cudaEventCreate(&se0) == cudaSuccess && cudaEventCreate(&se1) == cudaSuccess) cudaEventRecord(se0, c.stream);
    else { if (se0) cudaEventDestroy(se0); se0 = se1 = nullptr; }
  }
  if (rn > 0 && (double)M.nnz / (double)M.rn > t1) {
    // AMGB_SPMV=row32 | pipe16 | auto (default: 16 lanes per row, 8 lanes x 4 entries for a few
    // thousand rows of a few thousand entries; measured per matrix in profiles/r2_spmv_variants_poisson7_128.txt)
    static int spmv_kind = -1;
    if (spmv_kind < 0) { const char *e = getenv("AMGB_SPMV"); spmv_kind = (e && !strcmp(e, "row32")) ? 0 : (e && !strcmp(e, "pipe16")) ? 1 : (e && !strcmp(e, "pipe16x2")) ? 3 : 2; }
    if ((double)M.nnz / (double)M.rn <= 64.0 && !test_small_bins())
      k_spmv_tile<8><<<(rn + 31) / 32, 256, 0, c.stream>>>(rn, ro, col, vals, x, z, alpha, y, beta, plain, post, longrow, gen, want);
    else if (spmv_kind == 0)
      k_spmv_row32<<<(rn + 7) / 8, 256, 0, c.stream>>>(rn, ro, col, vals, x, z, alpha, y, beta, plain, post, longrow, gen, want);
    else if (spmv_kind == 3 && !((double)M.nnz / (double)M.rn >= 2048.0 && M.rn >= 2048))
      k_spmv_pipe<16, 2><<<(rn + 15) / 16, 256, 0, c.stream>>>(rn, ro, col, vals, x, z, alpha, y, beta, plain, post, longrow, gen, want);
    else if (spmv_kind == 1 || (spmv_kind == 2 && !((double)M.nnz / (double)M.rn >= 2048.0 && M.rn >= 2048) && !test_small_bins()))
      k_spmv_pipe<16, 1><<<(rn + 15) / 16, 256, 0, c.stream>>>(rn, ro, col, vals, x, z, alpha, y, beta, plain, post, longrow, gen, want);
    else
      k_spmv_pipe<8, 4><<<(rn + 31) / 32, 256, 0, c.stream>>>(rn, ro, col, vals, x, z, alpha, y, beta, plain, post, longrow, gen, want);
    c.launches++; post_launch("spmv_tile");
    done = true;
  }
  if (!done)
#endif
  parallel_for(rn, [=] DEV(i64 i) {
    if (ro[i + 1] - ro[i] > longrow) return;
    if (gen && gen[i] != want) return;
    double t = 0;
    if (x) for (int j = ro[i]; j < ro[i + 1]; j++) t = t + vals[j] * x[col[j]];
    else for (int j = ro[i]; j < ro[i + 1]; j++) t = t + vals[j] * 1.0;
    double r = plain ? beta * t : alpha * y[i] + beta * t;
    if (post) r = r * post[i];
    z[i] = r;
  });
#ifndef AMGB_EMU
  if (M.n_long > 0) {
    const int cap = M.rn < 16384 ? M.rn : 16384;
    const int grid = M.n_long < c.sm_count * 6 ? M.n_long : c.sm_count * 6;
    k_spmv_chain<<<grid, 256, 0, c.stream>>>(M.long_rows.p, M.long_rows.p + cap, M.ro.p, col, vals, x, zf, alpha, yf, beta, plain, postf, r0, r1, shift);
    c.launches++; post_launch("spmv_chain");
  }
  if (se0 && se1) {
    cudaEventRecord(se1, c.stream);
    g_spmv_stats.ev.emplace_back(se0, se1);
    // matrix once (values, and columns unless x is the implicit vector of ones), row offsets, z, and y when read
    g_spmv_stats.bytes += (x ? 12 : 8) * M.nnz + (plain ? 12 : 20) * (i64)M.rn;
    g_spmv_stats.calls++;
  }
#endif
}
void spmv(double *z, double alpha, const double *y, double beta, const Csr &M, const double *x) {
  spmv_vals(z, alpha, y, beta, M, M.a.p, x);
}

// Solve-phase storage partition: keep rows [r0, r1) of M (row offsets rebased to 0, the entries of
// the block), release the rest.  rn/cn/nnz keep describing the whole matrix.
i64 csr_keep_row_block(Csr &M, int r0, int r1) {
  if (M.partial) return 0;
  if (r0 < 0 || r1 > M.rn || r0 > r1) throw Error(-2, "csr_keep_row_block: bad row range");
  const i64 before = (i64)sizeof(int) * (M.rn + 1) + (i64)(sizeof(int) + sizeof(double)) * M.nnz;
  const int b = M.ro.get(r0), e = M.ro.get(r1);
  const int own = r1 - r0;
  Buf<int> ro2((i64)own + 1), col2((i64)(e - b));
  Buf<double> a2((i64)(e - b));
  { const int *ro = M.ro.p; int *o = ro2.p; parallel_for((i64)own + 1, [=] DEV(i64 i) { o[i] = ro[r0 + i] - b; }); }
  if (e > b) {
    d2d(col2.p, M.col.p + b, sizeof(int) * (size_t)(e - b));
    d2d(a2.p, M.a.p + b, sizeof(double) * (size_t)(e - b));
  }
  stream_sync();                       // the old buffers are released next
  M.ro = std::move(ro2); M.col = std::move(col2); M.a = std::move(a2);
  M.partial = true; M.row_lo = r0; M.own_rows = own; M.own_nnz = e - b;
  M.n_long = -1; M.long_rows.release();
  const i64 after = (i64)sizeof(int) * (own + 1) + (i64)(sizeof(int) + sizeof(double)) * (e - b);
  return before - after;
}

// ---------------------------------------------------------------------------------------
// row-local insertion sort by column (keys unique inside a row)
// ---------------------------------------------------------------------------------------
template <class V>
static HD inline void row_sort(int *key, V *val, int n) {
  for (int i = 1; i < n; i++) {
    int k = key[i]; V v = val[i];
    int j = i - 1;
    while (j >= 0 && key[j] > k) { key[j + 1] = key[j]; val[j + 1] = val[j]; j--; }
    key[j + 1] = k; val[j + 1] = v;
  }
}

#ifndef AMGB_EMU
// rank sort of every row of the transposed matrix by G cooperating threads: keys (source rows)
// are unique inside a row, so the rank of a key is the number of smaller keys
#define TR_LONG 96        // rows longer than this are sorted by a block (bitonic) instead
template <int G>
__global__ void __launch_bounds__(256) k_transpose_rank(int nrows, const int *tro, const int *kin, const int *sin,
                                                        int *kout, int *sout, const double *a, double *ta,
                                                        int *longlist, int *nlong) {
  const int c = blockIdx.x * (256 / G) + threadIdx.x / G;
  if (c >= nrows) return;
  const int r0 = threadIdx.x % G;
  const int b = tro[c], L = tro[c + 1] - b;
  if (L > TR_LONG && L <= 4096) { if (r0 == 0) longlist[atomicAdd(nlong, 1)] = c; return; }
  if (L > 4096 && L <= 16384) { if (r0 == 0) longlist[nrows - 1 - atomicAdd(nlong + 1, 1)] = c; return; }   // from the end of the list
  for (int e = r0; e < L; e += G) {
    const int key = kin[b + e];
    int rank = 0;
    for (int f = 0; f < L; f++) rank += (kin[b + f] < key);
    const int s = sin[b + e];
    kout[b + rank] = key; sout[b + rank] = s; ta[b + rank] = a[s];
  }
}
// long rows: one block, (key, source) pairs sorted by a bitonic network in shared memory
template <int T>
__global__ void __launch_bounds__(T) k_transpose_bitonic(const int *list, int nlist, const int *tro, const int *kin,
                                                         const int *sin, int *kout, int *sout, const double *a,
                                                         double *ta) {
  extern __shared__ int tsm[];
  if ((int)blockIdx.x >= nlist) return;
  const int c = list[blockIdx.x];
  const int b = tro[c], L = tro[c + 1] - b;
  int P = 128;
  while (P < L) P <<= 1;
  int *key = tsm, *src = tsm + P;
  for (int e = threadIdx.x; e < P; e += blockDim.x) { key[e] = e < L ? kin[b + e] : 0x7fffffff; src[e] = e < L ? sin[b + e] : -1; }
  __syncthreads();
  for (int k = 2; k <= P; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int idx = threadIdx.x; idx < P; idx += blockDim.x) {
        const int ixj = idx ^ j;
        if (ixj > idx) {
          const int ka = key[idx], kb = key[ixj];
          const bool up = ((idx & k) == 0);
          if ((ka > kb) == up && ka != kb) { key[idx] = kb; key[ixj] = ka; const int t = src[idx]; src[idx] = src[ixj]; src[ixj] = t; }
        }
      }
      __syncthreads();
    }
  for (int e = threadIdx.x; e < L; e += blockDim.x) { kout[b + e] = key[e]; sout[b + e] = src[e]; ta[b + e] = a[src[e]]; }
}
#endif

// ---------------------------------------------------------------------------------------
// transpose (:2000): A^t rows list the source rows in ascending order
// ---------------------------------------------------------------------------------------
#ifndef AMGB_EMU
// order-free work on every entry (i, j) of a matrix with long rows: one warp per row
template <class F>
__global__ void __launch_bounds__(256) k_row_entries(int rn, const int *ro, F f) {
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= rn) return;
  for (int j = ro[i] + lane; j < ro[i + 1]; j += 32) f(i, j);
}
#endif
// f(i, j) for every entry j of every row i; f must not depend on the order inside a row
template <class F>
static void for_row_entries(const Csr &A, F f) {
  const int *ro = A.ro.p;
#ifndef AMGB_EMU
  if (A.rn > 0 && A.nnz > 12 * (i64)A.rn) {
    Context &c = ctx();
    k_row_entries<<<(A.rn + 7) / 8, 256, 0, c.stream>>>(A.rn, ro, f);
    c.launches++; post_launch("row_entries");
    return;
  }
#endif
  parallel_for(A.rn, [=] DEV(i64 i) { for (int j = ro[i]; j < ro[i + 1]; j++) f((int)i, j); });
}

Csr transpose(const Csr &A, Buf<int> *tpos_out) {
  StageTimer st_("prim.transpose");
  Csr T(A.cn, A.rn, A.nnz);
  Buf<int> cnt(A.cn + 1);
  cnt.zero();
  const int *col = A.col.p;
  const double *a = A.a.p;
  int *cntp = cnt.p;
  parallel_for(A.nnz, [=] DEV(i64 e) { atomic_add(&cntp[col[e]], 1); });
  exclusive_scan(cnt.p, T.ro.p, A.cn);
  Buf<int> cursor(A.cn), src(A.nnz);
  d2d(cursor.p, T.ro.p, sizeof(int) * (size_t)A.cn);
  int *cur = cursor.p, *tcol = T.col.p, *srcp = src.p;
  // any order: the rows of T are sorted by source row below
  for_row_entries(A, [=] DEV(int i, int e) { const int p = atomic_add(&cur[col[e]], 1); tcol[p] = i; srcp[p] = e; });
  const int *tro = T.ro.p;
  double *ta = T.a.p;
#ifdef AMGB_EMU
  parallel_for(A.cn, [=] DEV(i64 c) {
    int b = tro[c], n = tro[c + 1] - b;
    row_sort<int>(tcol + b, srcp + b, n);
    for (int k = 0; k < n; k++) ta[b + k] = a[srcp[b + k]];
  });
#else
  if (A.cn > 0 && A.nnz > 0) {
    Buf<int> kin = T.col.clone(), sin = src.clone(), longlist(A.cn), nlong(2);
    nlong.zero();
    Context &c = ctx();
    if ((double)A.nnz / (double)A.cn <= 12.0)
      k_transpose_rank<8><<<(A.cn + 31) / 32, 256, 0, c.stream>>>(A.cn, tro, kin.p, sin.p, tcol, srcp, a, ta, longlist.p, nlong.p);
    else
      k_transpose_rank<32><<<(A.cn + 7) / 8, 256, 0, c.stream>>>(A.cn, tro, kin.p, sin.p, tcol, srcp, a, ta, longlist.p, nlong.p);
    c.launches++; post_launch("transpose_rank");
    const std::vector<int> hl = nlong.download();
    const int nl = hl[0], nl2 = hl[1];          // rows of 97..4096 and of 4097..16384 entries
    if (nl) {
      k_transpose_bitonic<128><<<nl, 128, 4096 * 8, c.stream>>>(longlist.p, nl, tro, kin.p, sin.p, tcol, srcp, a, ta);
      c.launches++; post_launch("transpose_bitonic");
    }
    if (nl2) {    // the column-0 pile of min_skel transposed: one row of thousands of entries
      static bool attr = false;
      if (!attr) { CUDA_CHECK(cudaFuncSetAttribute((const void *)k_transpose_bitonic<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8)); attr = true; }
      k_transpose_bitonic<512><<<nl2, 512, 16384 * 8, c.stream>>>(longlist.p + (A.cn - nl2), nl2, tro, kin.p, sin.p, tcol, srcp, a, ta);
      c.launches++; post_launch("transpose_bitonic_big");
    }
  }
#endif
  if (tpos_out) {
    tpos_out->alloc(A.nnz);
    int *tp = tpos_out->p;
    parallel_for(A.nnz, [=] DEV(i64 p) { tp[srcp[p]] = (int)p; });
  }
  return T;
}

// ---------------------------------------------------------------------------------------
// sub_mat (:3058)
// ---------------------------------------------------------------------------------------
#ifndef AMGB_EMU
// sub_mat for rows of more than a dozen entries: one warp per kept row, the kept entries keep
// their order (ballot + popcount prefix), loads and stores coalesced
__global__ void __launch_bounds__(256) k_sub_mat_count(int rn, const int *ro, const int *col, const int *rf, const int *rm,
                                                       const int *cf, int *cnt) {
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= rn || !rf[i]) return;
  int c = 0;
  for (int j = ro[i] + lane; j < ro[i + 1]; j += 32) c += cf[col[j]];
  c = __reduce_add_sync(0xffffffffu, c);
  if (lane == 0) cnt[rm[i]] = c;
}
__global__ void __launch_bounds__(256) k_sub_mat_fill(int rn, const int *ro, const int *col, const double *a, const int *rf,
                                                      const int *rm, const int *cf, const int *cm, const int *sr,
                                                      int *scol, double *sa) {
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= rn || !rf[i]) return;
  int p = sr[rm[i]];
  const int end = ro[i + 1];
  for (int base = ro[i]; base < end; base += 32) {
    const int j = base + lane;
    int c = 0;
    bool keep = false;
    if (j < end) { c = col[j]; keep = cf[c] != 0; }
    const unsigned b = __ballot_sync(0xffffffffu, keep);
    if (keep) { const int q = p + __popc(b & ((1u << lane) - 1u)); scol[q] = cm[c]; sa[q] = a[j]; }
    p += __popc(b);
  }
}
#endif

Csr sub_mat(const Csr &A, const double *vr, const double *vc) {
  StageTimer st_("prim.sub_mat");
  const int rn = A.rn, cn = A.cn;
  const int *ro = A.ro.p, *col = A.col.p;
  const double *a = A.a.p;
  Buf<int> rflag(rn + 1), rmap(rn + 1), cflag(cn + 1), cmap(cn + 1);
  int *rf = rflag.p, *cf = cflag.p;
  parallel_for(rn, [=] DEV(i64 i) { rf[i] = (vr == nullptr || vr[i] != 0) ? 1 : 0; });
  parallel_for(cn, [=] DEV(i64 i) { cf[i] = (vc == nullptr || vc[i] != 0) ? 1 : 0; });
  int subrn = (int)exclusive_scan(rflag.p, rmap.p, rn);
  int subcn = (int)exclusive_scan(cflag.p, cmap.p, cn);
  Buf<int> cnt(subrn + 1);
  int *cntp = cnt.p;
  const int *rm = rmap.p, *cm = cmap.p;
#ifndef AMGB_EMU
  const bool warp_rows = rn > 0 && A.nnz > 12 * (i64)rn;
  if (warp_rows) {
    Context &c = ctx();
    k_sub_mat_count<<<(rn + 7) / 8, 256, 0, c.stream>>>(rn, ro, col, rf, rm, cf, cntp);
    c.launches++; post_launch("sub_mat_count");
  } else
#endif
  parallel_for(rn, [=] DEV(i64 i) {
    if (!rf[i]) return;
    int c = 0;
    for (int j = ro[i]; j < ro[i + 1]; j++) c += cf[col[j]];
    cntp[rm[i]] = c;
  });
  Buf<int> sro(subrn + 1);
  i64 nnz = exclusive_scan(cnt.p, sro.p, subrn);
  Csr S(subrn, subcn, nnz);
  S.ro = std::move(sro);
  const int *sr = S.ro.p;
  int *scol = S.col.p;
  double *sa = S.a.p;
#ifndef AMGB_EMU
  if (warp_rows) {
    Context &c = ctx();
    k_sub_mat_fill<<<(rn + 7) / 8, 256, 0, c.stream>>>(rn, ro, col, a, rf, rm, cf, cm, sr, scol, sa);
    c.launches++; post_launch("sub_mat_fill");
    return S;
  }
#endif
  parallel_for(rn, [=] DEV(i64 i) {
    if (!rf[i]) return;
    int p = sr[rm[i]];
    for (int j = ro[i]; j < ro[i + 1]; j++)
      if (cf[col[j]]) { scol[p] = cm[col[j]]; sa[p] = a[j]; p++; }
  });
  return S;
}

// ---------------------------------------------------------------------------------------
// mpm (:1684): X = alpha*A + beta*B
// ---------------------------------------------------------------------------------------
#ifndef AMGB_EMU
// G lanes per row.  Every entry finds its partner in the other row by binary search; its place in
// the merged row is (own index) + (entries of the other row with a smaller column) - (coincident
// pairs before it, which are one entry instead of two) - (those of them whose sum is exactly 0,
// which mpm drops altogether).  The count kernel leaves that running correction per entry of A
// (dropsbefore) so that the fill kernel needs no ordered pass.
__device__ __forceinline__ int lower_bound_col(const int *col, int lo, int hi, int key) {
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (col[mid] < key) lo = mid + 1; else hi = mid; }
  return lo;
}
template <int G>
__global__ void __launch_bounds__(256) k_mpm_count(int rn, double alpha, const int *aro, const int *acol, const double *aa,
                                                   double beta, const int *bro, const int *bcol, const double *ba,
                                                   int *cnt, int *dropsbefore, int *ndrop) {
  const int i = blockIdx.x * (256 / G) + threadIdx.x / G;
  if (i >= rn) return;
  const int lane = threadIdx.x % G;
  const int sub = (threadIdx.x & 31) / G * G;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << sub);
  const int a0 = aro[i], a1 = aro[i + 1], b0 = bro[i], b1 = bro[i + 1];
  int carry = 0;
  for (int base = a0; base < a1; base += G) {
    const int j = base + lane;
    bool drop = false, match = false;
    if (j < a1) {
      const int c = acol[j];
      const int p = lower_bound_col(bcol, b0, b1, c);
      if (p < b1 && bcol[p] == c) { match = true; const double s = alpha * aa[j] + beta * ba[p]; drop = (s == 0.); }
    }
    const unsigned bd = (__ballot_sync(gmask, drop) >> sub) & ((G == 32) ? 0xffffffffu : ((1u << G) - 1u));
    const unsigned bm = (__ballot_sync(gmask, match) >> sub) & ((G == 32) ? 0xffffffffu : ((1u << G) - 1u));
    const unsigned lt = (1u << lane) - 1u;
    if (j < a1) dropsbefore[j] = carry + __popc(bd & lt) + __popc(bm & lt);
    carry += __popc(bd) + __popc(bm);
  }
  if (lane == 0) { cnt[i] = (a1 - a0) + (b1 - b0) - carry; ndrop[i] = carry; }
}
template <int G>
__global__ void __launch_bounds__(256) k_mpm_fill(int rn, double alpha, const int *aro, const int *acol, const double *aa,
                                                  double beta, const int *bro, const int *bcol, const double *ba,
                                                  const int *xro, const int *dropsbefore, const int *ndrop, int *xcol,
                                                  double *xa) {
  const int i = blockIdx.x * (256 / G) + threadIdx.x / G;
  if (i >= rn) return;
  const int lane = threadIdx.x % G;
  const int a0 = aro[i], a1 = aro[i + 1], b0 = bro[i], b1 = bro[i + 1], x0 = xro[i];
  for (int j = a0 + lane; j < a1; j += G) {
    const int c = acol[j];
    const int p = lower_bound_col(bcol, b0, b1, c);
    double v;
    if (p < b1 && bcol[p] == c) { v = alpha * aa[j] + beta * ba[p]; if (v == 0.) continue; }
    else v = alpha * aa[j];
    const int pos = x0 + (j - a0) + (p - b0) - dropsbefore[j];
    xcol[pos] = c; xa[pos] = v;
  }
  const int nd = ndrop[i];
  for (int k = b0 + lane; k < b1; k += G) {
    const int c = bcol[k];
    const int p = lower_bound_col(acol, a0, a1, c);
    if (p < a1 && acol[p] == c) continue;            // written from the A side
    const int pos = x0 + (k - b0) + (p - a0) - (p < a1 ? dropsbefore[p] : nd);
    xcol[pos] = c; xa[pos] = beta * ba[k];
  }
}
#endif

Csr mpm(double alpha, const Csr &A, double beta, const Csr &B) {
  StageTimer st_("prim.mpm");
  if (A.rn != B.rn || A.cn != B.cn) throw Error(-4, "mpm: dimension mismatch");
  const int rn = A.rn;
  const int *aro = A.ro.p, *acol = A.col.p, *bro = B.ro.p, *bcol = B.col.p;
  const double *aa = A.a.p, *ba = B.a.p;
  Buf<int> cnt(rn + 1);
  int *cntp = cnt.p;
#ifndef AMGB_EMU
  if (rn > 0 && ((double)(A.nnz + B.nnz) / rn > 12.0 || test_small_bins())) {
    Context &c = ctx();
    const bool wide = (double)(A.nnz + B.nnz) / rn > 64.0;
    Buf<int> dropsbefore(A.nnz), ndrop(rn);
    if (wide) k_mpm_count<32><<<(rn + 7) / 8, 256, 0, c.stream>>>(rn, alpha, aro, acol, aa, beta, bro, bcol, ba, cntp, dropsbefore.p, ndrop.p);
    else k_mpm_count<8><<<(rn + 31) / 32, 256, 0, c.stream>>>(rn, alpha, aro, acol, aa, beta, bro, bcol, ba, cntp, dropsbefore.p, ndrop.p);
    c.launches++; post_launch("mpm_count");
    Buf<int> xro(rn + 1);
    const i64 nnz = exclusive_scan(cnt.p, xro.p, rn);
    Csr X(rn, A.cn, nnz);
    X.ro = std::move(xro);
    if (wide) k_mpm_fill<32><<<(rn + 7) / 8, 256, 0, c.stream>>>(rn, alpha, aro, acol, aa, beta, bro, bcol, ba, X.ro.p, dropsbefore.p, ndrop.p, X.col.p, X.a.p);
    else k_mpm_fill<8><<<(rn + 31) / 32, 256, 0, c.stream>>>(rn, alpha, aro, acol, aa, beta, bro, bcol, ba, X.ro.p, dropsbefore.p, ndrop.p, X.col.p, X.a.p);
    c.launches++; post_launch("mpm_fill");
    return X;
  }
#endif
  parallel_for(rn, [=] DEV(i64 i) {
    int ja = aro[i], ea = aro[i + 1], jb = bro[i], eb = bro[i + 1], c = 0;
    while (ja < ea || jb < eb) {
      if (ja < ea && jb < eb && acol[ja] == bcol[jb]) {
        double s = alpha * aa[ja] + beta * ba[jb];
        if (s != 0.) c++;
        ja++; jb++;
      } else if (jb == eb || (ja < ea && acol[ja] < bcol[jb])) { c++; ja++; }
      else { c++; jb++; }
    }
    cntp[i] = c;
  });
  Buf<int> xro(rn + 1);
  i64 nnz = exclusive_scan(cnt.p, xro.p, rn);
  Csr X(rn, A.cn, nnz);
  X.ro = std::move(xro);
  const int *xr = X.ro.p;
  int *xcol = X.col.p;
  double *xa = X.a.p;
  parallel_for(rn, [=] DEV(i64 i) {
    int ja = aro[i], ea = aro[i + 1], jb = bro[i], eb = bro[i + 1], p = xr[i];
    while (ja < ea || jb < eb) {
      if (ja < ea && jb < eb && acol[ja] == bcol[jb]) {
        double s = alpha * aa[ja] + beta * ba[jb];
        if (s != 0.) { xcol[p] = acol[ja]; xa[p] = s; p++; }
        ja++; jb++;
      } else if (jb == eb || (ja < ea && acol[ja] < bcol[jb])) {
        xcol[p] = acol[ja]; xa[p] = alpha * aa[ja]; p++; ja++;
      } else {
        xcol[p] = bcol[jb]; xa[p] = beta * ba[jb]; p++; jb++;
      }
    }
  });
  return X;
}

// ---------------------------------------------------------------------------------------
// mxmpoint (:1807): X = A.*B on the intersection pattern (zeros kept)
// ---------------------------------------------------------------------------------------
Csr mxmpoint(const Csr &A, const Csr &B) {
  StageTimer st_("prim.mxmpoint");
  if (A.rn != B.rn || A.cn != B.cn) throw Error(-4, "mxmpoint: dimension mismatch");
  const int rn = A.rn;
  const int *aro = A.ro.p, *acol = A.col.p, *bro = B.ro.p, *bcol = B.col.p;
  const double *aa = A.a.p, *ba = B.a.p;
  Buf<int> cnt(rn + 1);
  int *cntp = cnt.p;
  parallel_for(rn, [=] DEV(i64 i) {
    int ja = aro[i], ea = aro[i + 1], jb = bro[i], eb = bro[i + 1], c = 0;
    while (ja < ea && jb < eb) {
      if (acol[ja] == bcol[jb]) { c++; ja++; jb++; }
      else if (acol[ja] < bcol[jb]) ja++;
      else jb++;
    }
    cntp[i] = c;
  });
  Buf<int> xro(rn + 1);
  i64 nnz = exclusive_scan(cnt.p, xro.p, rn);
  Csr X(rn, A.cn, nnz);
  X.ro = std::move(xro);
  const int *xr = X.ro.p;
  int *xcol = X.col.p;
  double *xa = X.a.p;
  parallel_for(rn, [=] DEV(i64 i) {
    int ja = aro[i], ea = aro[i + 1], jb = bro[i], eb = bro[i + 1], p = xr[i];
    while (ja < ea && jb < eb) {
      if (acol[ja] == bcol[jb]) { xcol[p] = acol[ja]; xa[p] = aa[ja] * ba[jb]; p++; ja++; jb++; }
      else if (acol[ja] < bcol[jb]) ja++;
      else jb++;
    }
  });
  return X;
}

// ---------------------------------------------------------------------------------------
// build_csr_dim (:3656)
// ---------------------------------------------------------------------------------------
Csr coo_to_csr(i64 n, const int *Ai, const int *Aj, const double *Av, int rn, int cn) {
  Buf<int> cnt(rn + 1);
  cnt.zero();
  int *cntp = cnt.p;
  Buf<int> bad(1);
  bad.zero();
  int *badp = bad.p;
  parallel_for(n, [=] DEV(i64 e) {
    if (Av[e] == 0.) return;
    if (Ai[e] < 0 || Ai[e] >= rn || Aj[e] < 0 || Aj[e] >= cn) { *badp = 1; return; }
    atomic_add(&cntp[Ai[e]], 1);
  });
  Buf<int> xro(rn + 1);
  i64 nnz = exclusive_scan(cnt.p, xro.p, rn);
  if (bad.get(0)) throw Error(-5, "COO index out of range");
  Csr X(rn, cn, nnz);
  X.ro = std::move(xro);
  Buf<int> cursor(rn + 1);
  d2d(cursor.p, X.ro.p, sizeof(int) * (size_t)(rn + 1));
  int *cur = cursor.p, *xcol = X.col.p;
  double *xa = X.a.p;
  parallel_for(n, [=] DEV(i64 e) {
    if (Av[e] == 0.) return;
    int p = atomic_add(&cur[Ai[e]], 1);
    xcol[p] = Aj[e]; xa[p] = Av[e];
  });
  const int *xr = X.ro.p;
  parallel_for(rn, [=] DEV(i64 i) {
    int b = xr[i], m = xr[i + 1] - b;
    row_sort<double>(xcol + b, xa + b, m);
    for (int k = 1; k < m; k++) if (xcol[b + k] == xcol[b + k - 1]) *badp = 2;
  });
  if (bad.get(0) == 2) throw Error(-6, "duplicate (row,col) entries: assemble the matrix first");
  return X;
}

// ---------------------------------------------------------------------------------------
// diagonal helpers
// ---------------------------------------------------------------------------------------
void diag_of(double *D, const Csr &A) {
  const int *col = A.col.p;
  const double *a = A.a.p;
  fill(D, A.rn, 0.);                              // rows without a diagonal entry (columns are unique inside a row)
  for_row_entries(A, [=] DEV(int i, int j) { if (col[j] == i) D[i] = a[j]; });
}
void scale_rows(Csr &A, const double *D) {
  double *a = A.a.p;
  for_row_entries(A, [=] DEV(int i, int j) { a[j] = a[j] * D[i]; });
}
void scale_cols(Csr &A, const double *D) {
  const int *col = A.col.p;
  double *a = A.a.p;
  parallel_for(A.nnz, [=] DEV(i64 e) { a[e] = a[e] * D[col[e]]; });
}
void sub_diag(Csr &A, const double *D) {
  const int *col = A.col.p;
  double *a = A.a.p;
  for_row_entries(A, [=] DEV(int i, int j) { if (col[j] == i) a[j] = a[j] - D[i]; });
}
void col_sums(double *s, const Csr &A) {
  Csr T = transpose(A);
  const int *ro = T.ro.p;
  const double *a = T.a.p;
  parallel_for(T.rn, [=] DEV(i64 c) {
    double t = 0.0;
    for (int j = ro[c]; j < ro[c + 1]; j++) t = t + a[j];
    s[c] = t;
  });
}
int max_row_len(const Csr &A) {
  if (A.rn == 0) return 0;
  Buf<double> len(A.rn);
  const int *ro = A.ro.p;
  double *lp = len.p;
  parallel_for(A.rn, [=] DEV(i64 i) { lp[i] = (double)(ro[i + 1] - ro[i]); });
  double m; i64 idx;
  max_first(len.p, A.rn, &m, &idx);
  return (int)m;
}

// ---------------------------------------------------------------------------------------
// SpGEMM, generation 1 (reference semantics of mxm :1894): one logical thread per row, an
// open-addressing table per row in HBM.  Every X[i][c] is accumulated over k ascending.
// ---------------------------------------------------------------------------------------
// Row-partitioned X = A*B (one process per GPU): rank r forms the rows [row_split(r),
// row_split(r+1)) of X from its row block of A with the single-GPU kernel `local`; the row
// lengths and then the entries of the blocks are exchanged in place (comm_allgatherv), so every
// rank ends with the whole X.  Each row is computed by exactly one rank with the same kernel as
// on one GPU, hence X is bit-identical to local(A, B).
Csr spgemm_partitioned(const Csr &A, const Csr &B, Csr (*local)(const Csr &, const Csr &)) {
  const int P = comm_size(), me = comm_rank();
  const int rn = A.rn;
  const int r0 = (int)row_split(rn, me), r1 = (int)row_split(rn, me + 1), ln = r1 - r0;
  Csr Xl;
  {
    // the row block of A (offsets rebased); its entries are one contiguous range of A
    Buf<int> span(2);
    const int *aro = A.ro.p;
    int *sp = span.p;
    parallel_for(1, [=] DEV(i64) { sp[0] = aro[r0]; sp[1] = aro[r1]; });
    std::vector<int> hs = span.download();
    Csr Al(ln, A.cn, (i64)hs[1] - hs[0]);
    int *lro = Al.ro.p;
    const int base = hs[0];
    parallel_for((i64)ln + 1, [=] DEV(i64 i) { lro[i] = aro[r0 + i] - base; });
    if (Al.nnz) {
      d2d(Al.col.p, A.col.p + base, sizeof(int) * (size_t)Al.nnz);
      d2d(Al.a.p, A.a.p + base, sizeof(double) * (size_t)Al.nnz);
    }
    // the block inherits an identity derived from A's, so that the entry-tier hints of the local
    // product (spgemm.cu) carry over from one product with this left operand to the next
    Al.uid = A.uid ? A.uid * 1000003ULL + (unsigned long long)me + 1ULL : 0ULL;
    Xl = local(Al, B);
  }
  // row lengths of all blocks, then the global offsets
  Buf<int> cnt((i64)rn + 1), xro((i64)rn + 1);
  {
    int *cp = cnt.p;
    const int *lro = Xl.ro.p;
    parallel_for(ln, [=] DEV(i64 i) { cp[r0 + i] = lro[i + 1] - lro[i]; });
    std::vector<i64> off((size_t)P + 1);
    for (int r = 0; r <= P; r++) off[(size_t)r] = (i64)sizeof(int) * row_split(rn, r);
    comm_allgatherv(cnt.p, off.data(), "comm.spgemm_counts");
  }
  const i64 nnz = exclusive_scan(cnt.p, xro.p, rn);
  Csr X(rn, B.cn, nnz);
  X.ro = std::move(xro);
  std::vector<i64> seg((size_t)P + 1);
  {
    Buf<int> sb((i64)P + 1);
    int *sp = sb.p;
    const int *xr = X.ro.p;
    const i64 rnl = rn;
    parallel_for((i64)P + 1, [=] DEV(i64 r) { sp[r] = xr[rnl * r / P]; });
    std::vector<int> hs = sb.download();
    for (int r = 0; r <= P; r++) seg[(size_t)r] = hs[(size_t)r];
  }
  if (seg[(size_t)me + 1] - seg[(size_t)me] != Xl.nnz) throw Error(-114, "spgemm_partitioned: block size mismatch");
  if (Xl.nnz) {
    d2d(X.col.p + seg[(size_t)me], Xl.col.p, sizeof(int) * (size_t)Xl.nnz);
    d2d(X.a.p + seg[(size_t)me], Xl.a.p, sizeof(double) * (size_t)Xl.nnz);
  }
  std::vector<i64> off((size_t)P + 1);
  for (int r = 0; r <= P; r++) off[(size_t)r] = (i64)sizeof(int) * seg[(size_t)r];
  comm_allgatherv(X.col.p, off.data(), "comm.spgemm_cols");
  for (int r = 0; r <= P; r++) off[(size_t)r] = (i64)sizeof(double) * seg[(size_t)r];
  comm_allgatherv(X.a.p, off.data(), "comm.spgemm_vals");
  return X;
}

Csr spgemm_rowhash(const Csr &A, const Csr &B) {
  if (A.cn != B.rn) throw Error(-4, "spgemm: dimension mismatch");
  const int rn = A.rn;
  const int *aro = A.ro.p, *acol = A.col.p, *bro = B.ro.p, *bcol = B.col.p;
  const double *aa = A.a.p, *ba = B.a.p;
  Buf<i64> hsz(rn + 1), hoff(rn + 1);
  i64 *hszp = hsz.p;
  parallel_for(rn, [=] DEV(i64 i) {
    i64 ub = 0;
    for (int ja = aro[i]; ja < aro[i + 1]; ja++) ub += bro[acol[ja] + 1] - bro[acol[ja]];
    i64 sz = 0;
    if (ub > 0) { sz = 4; while (sz < 2 * ub) sz <<= 1; }
    hszp[i] = sz;
  });
  i64 total = exclusive_scan64(hsz.p, hoff.p, rn);
  Buf<int> keys(total), cnt(rn + 1), rowlen(rn + 1);
  Buf<double> vals(total);
  dev_memset(keys.p, 0xFF, sizeof(int) * (size_t)total);
  int *kp = keys.p, *cntp = cnt.p, *rl = rowlen.p;
  double *vp = vals.p;
  const i64 *ho = hoff.p;
  parallel_for(rn, [=] DEV(i64 i) {
    const i64 base = ho[i], sz = ho[i + 1] - base;
    if (sz == 0) { cntp[i] = 0; rl[i] = 0; return; }
    int *K = kp + base;
    double *V = vp + base;
    const unsigned mask = (unsigned)(sz - 1);
    for (int ja = aro[i]; ja < aro[i + 1]; ja++) {
      const int k = acol[ja];
      const double av = aa[ja];
      for (int jb = bro[k]; jb < bro[k + 1]; jb++) {
        const int c = bcol[jb];
        unsigned h = ((unsigned)c * 2654435761u) & mask;
        while (K[h] != -1 && K[h] != c) h = (h + 1) & mask;
        if (K[h] == -1) { K[h] = c; V[h] = 0.0; }
        V[h] = V[h] + ba[jb] * av;
      }
    }
    int m = 0;
    for (i64 h = 0; h < sz; h++) if (K[h] != -1) { K[m] = K[h]; V[m] = V[h]; m++; }
    row_sort<double>(K, V, m);
    int nz = 0;
    for (int q = 0; q < m; q++) nz += (V[q] != 0.0);
    rl[i] = m; cntp[i] = nz;
  });
  Buf<int> xro(rn + 1);
  i64 nnz = exclusive_scan(cnt.p, xro.p, rn);
  Csr X(rn, B.cn, nnz);
  X.ro = std::move(xro);
  const int *xr = X.ro.p;
  int *xcol = X.col.p;
  double *xa = X.a.p;
  parallel_for(rn, [=] DEV(i64 i) {
    const i64 base = ho[i];
    int p = xr[i];
    for (int q = 0; q < rl[i]; q++)
      if (vp[base + q] != 0.0) { xcol[p] = kp[base + q]; xa[p] = vp[base + q]; p++; }
  });
  return X;
}

}  // namespace amgb

// localsolve.cu -- cooperative kernels for the local A-orthogonalisation (see localsolve.cuh).
//
// Arithmetic contract (what makes the results bit-identical to the reference): every sum is
// formed in the reference's order -- mv_utt rows left to right, mv_ut columns in ascending j
// starting from 0, the alpha recurrence in ascending m, QQ^t in ascending k -- with separate
// multiply and add (the library is compiled with --fmad=false).  Parallelism is only ever over
// independent outputs, never inside one sum.
#include "localsolve.cuh"
#include "comm.cuh"

#ifndef AMGB_EMU
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
#endif

namespace amgb {

static HD inline i64 tri(i64 k) { return k * (k + 1) / 2; }

// value of the sorted sparse row (xi, x, xn) at index t, 0 if absent (sp_restrict_sorted :2180)
static HD inline double row_at(const int *xi, const double *x, int xn, int t) {
  int lo = 0, hi = xn;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (xi[mid] < t) lo = mid + 1; else hi = mid; }
  return (lo < xn && xi[lo] == t) ? x[lo] : 0.0;
}
static HD inline int pos_of(const int *xi, int xn, int t) {
  int lo = 0, hi = xn;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (xi[mid] < t) lo = mid + 1; else hi = mid; }
  return (lo < xn && xi[lo] == t) ? lo : -1;
}

void q_offsets(QStore &qs, const Csr &Wt) {
  Buf<i64> sz(Wt.rn + 1), len(1);
  qs.qoff.alloc(Wt.rn + 1);
  const int *ro = Wt.ro.p;
  i64 *s = sz.p;
  parallel_for(Wt.rn, [=] DEV(i64 i) { const i64 nz = ro[i + 1] - ro[i]; s[i] = nz * (nz + 1) / 2; });
  qs.total = exclusive_scan64(sz.p, qs.qoff.p, Wt.rn);
  qs.maxnz = max_row_len(Wt);
  qs.Q.alloc(qs.total);
}
void qq_offsets(QQStore &qq, const Csr &Wt) {
  Buf<i64> sz(Wt.rn + 1);
  qq.qqoff.alloc(Wt.rn + 1);
  const int *ro = Wt.ro.p;
  i64 *s = sz.p;
  parallel_for(Wt.rn, [=] DEV(i64 i) { const i64 nz = ro[i + 1] - ro[i]; s[i] = nz * nz; });
  const i64 total = exclusive_scan64(sz.p, qq.qqoff.p, Wt.rn);
  qq.QQ.alloc(total);
}

// Partition of the coarse columns over the ranks (one process per GPU): contiguous column
// ranges holding equal shares of the Q store (its size stands in for the work).  cs[r] is the
// first column of rank r, off[r] the byte offset of its Q blocks; false = not partitioned.
static bool q_partition(const QStore &qs, int n, std::vector<int> &cs, std::vector<i64> &off) {
  const int P = comm_size();
  if (P <= 1 || n < P || qs.total < comm_min_work()) return false;
  Buf<i64> sb(2 * ((i64)P + 1));
  i64 *sp = sb.p;
  const i64 *qo = qs.qoff.p;
  const i64 total = qs.total;
  parallel_for((i64)P + 1, [=] DEV(i64 r) {
    const i64 target = total / P * r;
    int lo = 0, hi = n;
    if (r == P) lo = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (qo[mid] < target) lo = mid + 1; else hi = mid; }
    sp[2 * r] = lo; sp[2 * r + 1] = qo[lo];
  });
  std::vector<i64> h = sb.download();
  cs.resize((size_t)P + 1); off.resize((size_t)P + 1);
  for (int r = 0; r <= P; r++) { cs[(size_t)r] = (int)h[2 * (size_t)r]; off[(size_t)r] = (i64)sizeof(double) * h[2 * (size_t)r + 1]; }
  return true;
}

#ifdef AMGB_EMU
// =======================================================================================
// host emulation: straight restatement, one logical thread per column / row
// =======================================================================================
void build_q_store(QStore &qs, const Csr &Wt, const Csr &At) {
  q_offsets(qs, Wt);
  Buf<double> scratch((i64)2 * (qs.maxnz + 1));
  const int *wro = Wt.ro.p, *wcol = Wt.col.p, *aro = At.ro.p, *acol = At.col.p;
  const double *aa = At.a.p;
  double *Qall = qs.Q.p, *sc = scratch.p;
  const i64 *qo = qs.qoff.p;
  const int mx = qs.maxnz + 1;
  std::vector<int> cs;
  std::vector<i64> off;
  const bool part = q_partition(qs, Wt.rn, cs, off);
  const int c0 = part ? cs[(size_t)comm_rank()] : 0, c1 = part ? cs[(size_t)comm_rank() + 1] : Wt.rn;
  parallel_for(c1 - c0, [=] DEV(i64 ii) {
    const i64 i = c0 + ii;
    const int b = wro[i], nz = wro[i + 1] - b;
    const int *Qj = wcol + b;
    double *Q = Qall + qo[i], *sqv1 = sc, *sqv2 = sc + mx;
    for (int k = 0; k < nz; k++) {
      const int s = Qj[k];
      double *qk = Q + tri(k);
      for (int m = 0; m <= k; m++) sqv1[m] = row_at(acol + aro[s], aa + aro[s], aro[s + 1] - aro[s], Qj[m]);
      for (int r = 0; r < k; r++) { double v = 0; const double *u = Q + tri(r); for (int j = 0; j <= r; j++) v = v + u[j] * sqv1[j]; sqv2[r] = v; }
      for (int r = 0; r < k; r++) { double y = 0; for (int j = r; j < k; j++) y = y + Q[tri(j) + r] * sqv2[j]; qk[r] = y; }
      double alpha = sqv1[k];
      for (int m = 0; m < k; m++) alpha = alpha - sqv1[m] * qk[m];
      alpha = -1.0 / sqrt(alpha);
      for (int m = 0; m < k; m++) qk[m] = qk[m] * alpha;
      qk[k] = -alpha;
    }
  });
  if (part) comm_allgatherv(qs.Q.p, off.data(), "comm.q_blocks");
}
void apply_q(const QStore &qs, Csr &Wt, const Csr &Bt, const double *u, const double *lambda) {
  Buf<double> scratch((i64)2 * (qs.maxnz + 1));
  const int *wro = Wt.ro.p, *wcol = Wt.col.p, *bro = Bt.ro.p, *bcol = Bt.col.p;
  const double *ba = Bt.a.p, *Qall = qs.Q.p;
  double *wa = Wt.a.p, *sc = scratch.p;
  const i64 *qo = qs.qoff.p;
  const int mx = qs.maxnz + 1;
  parallel_for(Wt.rn, [=] DEV(i64 i) {
    const int b = wro[i], nz = wro[i + 1] - b;
    const int *Qj = wcol + b;
    const double *Q = Qall + qo[i];
    double *sqv1 = sc, *sqv2 = sc + mx;
    for (int k = 0; k < nz; k++) {
      const double v = row_at(bcol + bro[i], ba + bro[i], bro[i + 1] - bro[i], Qj[k]);
      sqv1[k] = v + u[i] * lambda[Qj[k]];
    }
    for (int r = 0; r < nz; r++) { double v = 0; const double *uu = Q + tri(r); for (int j = 0; j <= r; j++) v = v + uu[j] * sqv1[j]; sqv2[r] = v; }
    for (int r = 0; r < nz; r++) { double y = 0; for (int j = r; j < nz; j++) y = y + Q[tri(j) + r] * sqv2[j]; wa[b + r] = y; }
  });
}
void form_qq(QQStore &qq, const QStore &qs, const Csr &Wt) {
  qq_offsets(qq, Wt);
  const int *wro = Wt.ro.p;
  const double *Qall = qs.Q.p;
  double *QQ = qq.QQ.p;
  const i64 *qo = qs.qoff.p, *qqo = qq.qqoff.p;
  parallel_for(Wt.rn, [=] DEV(i64 i) {
    const int nz = wro[i + 1] - wro[i];
    const double *Q = Qall + qo[i];
    double *out = QQ + qqo[i];
    for (int m = 0; m < nz; m++)
      for (int j = m; j < nz; j++) {
        double acc = 0;
        for (int k = j; k < nz; k++) acc = acc + Q[tri(k) + m] * Q[tri(k) + j];
        out[(i64)m * nz + j] = acc; out[(i64)j * nz + m] = acc;
      }
  });
}
void lmop_accumulate(Csr &S, const QQStore &qq, const double *u, const Csr &Wskt, const Csr &Wsk,
                     const int *tpos) {
  const int *sro = S.ro.p, *scol = S.col.p, *kro = Wsk.ro.p, *kcol = Wsk.col.p, *tro = Wskt.ro.p, *tcol = Wskt.col.p;
  double *sa = S.a.p;
  const double *QQ = qq.QQ.p;
  const i64 *qqo = qq.qqoff.p;
  parallel_for(S.rn, [=] DEV(i64 j) {
    const int yb = sro[j], yn = sro[j + 1] - yb;
    for (int q = 0; q < yn; q++) sa[yb + q] = 0.0;
    if (yn == 0) return;
    for (int e = kro[j]; e < kro[j + 1]; e++) {
      const int i = kcol[e];
      const int b = tro[i], nz = tro[i + 1] - b, k = tpos[e] - b;
      const double ui = u[i];
      const double *x = QQ + qqo[i] + (i64)k * nz;
      for (int kk = 0; kk < nz; kk++) {
        const int p = pos_of(scol + yb, yn, tcol[b + kk]);
        if (p >= 0) sa[yb + p] = sa[yb + p] + ui * x[kk];
      }
    }
  });
}

#else
// =======================================================================================
// CUDA: G cooperating threads per coarse column (G = 8, 32, or a whole block)
// =======================================================================================
// Ordered sums whose left operand streams from HBM/L2: v (+|-)= a[j]*b[j] for j ascending, one add
// after the other as in the reference, while the loads of a run one batch of U ahead of the adds
// (the sum is a dependent chain; without the explicit batches every add waits for its own load).
// The prefetch is unconditional (indices clamped to the last element) so that it stays a straight
// line of loads ahead of the adds.
template <int U, bool SUB>
__device__ __forceinline__ double chain_dot(double v, const double *a, const double *b, int n) {
  if (n <= 0) return v;
  double cur[U], nxt[U];
  const int last = n - 1;
#pragma unroll
  for (int t = 0; t < U; t++) cur[t] = a[min(t, last)];
  int j = 0;
  while (j + U <= n) {
#pragma unroll
    for (int t = 0; t < U; t++) nxt[t] = a[min(j + U + t, last)];
#pragma unroll
    for (int t = 0; t < U; t++) { const double p = cur[t] * b[j + t]; v = SUB ? v - p : v + p; }
#pragma unroll
    for (int t = 0; t < U; t++) cur[t] = nxt[t];
    j += U;
  }
#pragma unroll
  for (int t = 0; t < U; t++)
    if (j + t < n) { const double p = cur[t] * b[j + t]; v = SUB ? v - p : v + p; }
  return v;
}
// the same for a strided walk: element j sits at offset o_j with o_{j+1} = o_j + inc(j);
// sum over j in [j0, j1) of Q[o_j] * s[j], o_{j0} = o0.  Offsets past the end are clamped to the
// last element's (never used in the sum).
template <int U, class Inc>
__device__ __forceinline__ double chain_walk(const double *Q, i64 o0, int j0, int j1, const double *s, Inc inc) {
  double v = 0;
  if (j1 <= j0) return v;
  double cur[U], nxt[U];
  int j = j0;
  i64 o = o0;                      // offset of element j
  {
    i64 q = o;
#pragma unroll
    for (int t = 0; t < U; t++) { cur[t] = Q[q]; if (j + t + 1 < j1) q += inc(j + t); }
  }
  while (j + U <= j1) {
    i64 on = o;
#pragma unroll
    for (int t = 0; t < U; t++) if (j + t + 1 < j1) on += inc(j + t);      // offset of element j + U (clamped)
    {
      i64 q = on;
#pragma unroll
      for (int t = 0; t < U; t++) { nxt[t] = Q[q]; if (j + U + t + 1 < j1) q += inc(j + U + t); }
    }
#pragma unroll
    for (int t = 0; t < U; t++) v = v + cur[t] * s[j + t];
#pragma unroll
    for (int t = 0; t < U; t++) cur[t] = nxt[t];
    j += U; o = on;
  }
#pragma unroll
  for (int t = 0; t < U; t++)
    if (j + t < j1) v = v + cur[t] * s[j + t];
  return v;
}
// a column of the packed triangle: sum over j in [j0, j1) of Q[tri(j) + r] * s[j]
template <int U>
__device__ __forceinline__ double chain_col(const double *Q, int r, int j0, int j1, const double *s) {
  return chain_walk<U>(Q, tri(j0) + r, j0, j1, s, [](int j) { return (i64)(j + 1); });
}

// On entry Q holds the Gram rows: Q[tri(k) + m] = A[Qj[k]][Qj[m]], m <= k (k_gram_fill) -- the
// slot of column k of Q is exactly the restricted row of A that step k starts from, so the serial
// k loop below touches no matrix data at all.  PIPE: Q lives in HBM/L2 (batched loads).
template <bool PIPE, class Group>
__device__ __forceinline__ void build_q_coop(const Group &g, int nz, double *Q, double *sqv1, double *sqv2) {
  const int r0 = g.thread_rank(), G = g.size();
  for (int k = 0; k < nz; k++) {
    double *qk = Q + tri(k);
    for (int m = r0; m <= k; m += G) sqv1[m] = qk[m];
    g.sync();
    for (int r = r0; r < k; r += G) {                 // mv_utt: row r of Q^t, left to right
      const double *u = Q + tri(r);
      double v = 0;
      if (PIPE) v = chain_dot<8, false>(0.0, u, sqv1, r + 1);
      else for (int j = 0; j <= r; j++) v = v + u[j] * sqv1[j];
      sqv2[r] = v;
    }
    g.sync();
    for (int r = r0; r < k; r += G) {                 // mv_ut: ascending j, starting from 0
      double y = 0;
      if (PIPE) y = chain_col<8>(Q, r, r, k, sqv2);
      else for (int j = r; j < k; j++) y = y + Q[tri(j) + r] * sqv2[j];
      qk[r] = y;
    }
    g.sync();
    double alpha = sqv1[k];                           // every thread forms the same recurrence
    if (PIPE) alpha = chain_dot<8, true>(alpha, qk, sqv1, k);
    else for (int m = 0; m < k; m++) alpha = alpha - sqv1[m] * qk[m];
    alpha = -1.0 / sqrt(alpha);
    g.sync();
    for (int m = r0; m < k; m += G) qk[m] = qk[m] * alpha;
    if (r0 == 0) qk[k] = -alpha;
    g.sync();
  }
}

struct BlockGroup {
  __device__ __forceinline__ int thread_rank() const { return threadIdx.x; }
  __device__ __forceinline__ int size() const { return blockDim.x; }
  __device__ __forceinline__ void sync() const { __syncthreads(); }
};

// small columns: 8 threads per column, everything in shared memory
template <int NZCAP>
__global__ void __launch_bounds__(256) k_build_q_tile8(const int *list, int nlist, const int *wro, const int *wcol,
                                                       const int *aro, const int *acol, const double *aa,
                                                       double *Qall, const i64 *qoff) {
  constexpr int PER = 2 * NZCAP + NZCAP * (NZCAP + 1) / 2;
  __shared__ double sm[32 * PER];
  auto tile = cg::tiled_partition<8>(cg::this_thread_block());
  const int slot = threadIdx.x / 8;
  const int idx = blockIdx.x * 32 + slot;
  if (idx >= nlist) return;
  const int i = list[idx];
  const int b = wro[i], nz = wro[i + 1] - b;
  double *sqv1 = sm + slot * PER, *sqv2 = sqv1 + NZCAP, *Q = sqv2 + NZCAP;
  double *out = Qall + qoff[i];
  const int nq = (int)tri(nz);
  for (int t = tile.thread_rank(); t < nq; t += 8) Q[t] = out[t];
  tile.sync();
  build_q_coop<false>(tile, nz, Q, sqv1, sqv2);
  for (int t = tile.thread_rank(); t < nq; t += 8) out[t] = Q[t];
}

// one block per column; Q in shared memory (QSMEM) or directly in the global store
template <bool QSMEM>
__global__ void k_build_q_block(const int *list, int nlist, int nzcap, const int *wro, const int *wcol,
                                const int *aro, const int *acol, const double *aa, double *Qall,
                                const i64 *qoff) {
  extern __shared__ double sm[];
  const int idx = blockIdx.x;
  if (idx >= nlist) return;
  const int i = list[idx];
  const int b = wro[i], nz = wro[i + 1] - b;
  double *sqv1 = sm, *sqv2 = sm + nzcap;
  double *Qg = Qall + qoff[i];
  double *Q = QSMEM ? (sm + 2 * nzcap) : Qg;
  BlockGroup g;
  const int nq = (int)tri(nz);
  if (QSMEM) {
    for (int t = threadIdx.x; t < nq; t += blockDim.x) Q[t] = Qg[t];
    __syncthreads();
  }
  build_q_coop<!QSMEM>(g, nz, Q, sqv1, sqv2);
  if (QSMEM) {
    for (int t = threadIdx.x; t < nq; t += blockDim.x) Qg[t] = Q[t];
  }
}

// Gram rows of the columns list[0..nlist): one warp per column, the lanes take the entries
// (k, m <= k) side by side.  No barriers and no serial dependence: the searches of all columns
// overlap, which is what hides their latency.
__global__ void __launch_bounds__(256) k_gram_fill(const int *list, int nlist, const int *wro, const int *wcol,
                                                   const int *aro, const int *acol, const double *aa,
                                                   double *Qall, const i64 *qoff) {
  const int w = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (w >= nlist) return;
  const int i = list[w];
  const int b = wro[i], nz = wro[i + 1] - b;
  const int *Qj = wcol + b;
  double *Q = Qall + qoff[i];
  for (int k = 0; k < nz; k++) {
    const int s = Qj[k];
    const int ab = aro[s], an = aro[s + 1] - ab;
    double *qk = Q + tri(k);
    for (int m = lane; m <= k; m += 32) qk[m] = row_at(acol + ab, aa + ab, an, Qj[m]);
  }
}

// Gram rows of columns with a very large support: blockIdx.y = column, the pairs (k, m <= k) are
// spread over gridDim.x blocks
__global__ void __launch_bounds__(256) k_gram_fill_big(const int *list, const int *wro, const int *wcol, const int *aro,
                                                       const int *acol, const double *aa, double *Qall, const i64 *qoff) {
  const int i = list[blockIdx.y];
  const int b = wro[i], nz = wro[i + 1] - b;
  const int *Qj = wcol + b;
  double *Q = Qall + qoff[i];
  for (int k = blockIdx.x * 8 + (threadIdx.x >> 5); k < nz; k += gridDim.x * 8) {
    const int s = Qj[k];
    const int ab = aro[s], an = aro[s + 1] - ab;
    double *qk = Q + tri(k);
    for (int m = threadIdx.x & 31; m <= k; m += 32) qk[m] = row_at(acol + ab, aa + ab, an, Qj[m]);
  }
}

// Very large supports (the reference piles every F row without a coupling into column 0, :2229;
// 1050 rows at 128^3): one column is worked on by a CLUSTER of 8 blocks (2048 threads, 8 SMs'
// worth of L2 bandwidth -- every step re-reads the whole triangle built so far).  Q lives in
// HBM/L2; the vectors of a step are exchanged through L2 and copied into every block's shared
// memory; three cluster barriers per step.  The arithmetic per output is that of build_q_coop.
__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(256)
k_build_q_cluster(const int *list, int nlist, int maxnz, const int *wro, double *Qall, const i64 *qoff, double *scratch) {
  extern __shared__ double sm[];
  cg::cluster_group cl = cg::this_cluster();
  __shared__ double alpha_sh;
  const int cidx = blockIdx.x / 8;
  if (cidx >= nlist) return;                      // whole cluster leaves together
  const int i = list[cidx];
  const int nz = wro[i + 1] - wro[i];
  double *Q = Qall + qoff[i];
  // per column: the two exchange vectors and a second copy Q2 of the triangle in the other
  // packing (element (k, m) at m*nz - m(m-1)/2 + k - m), so that the threads r = 0,1,2.. of the
  // row sums below read neighbouring addresses as well -- rows of the packed Q are tri(r) apart,
  // which costs one L1 wavefront per thread and load
  double *g2 = scratch + (size_t)cidx * (2 * (size_t)maxnz + (size_t)tri(maxnz)), *gy = g2 + maxnz, *Q2 = gy + maxnz;
  double *sqv1 = sm, *sqv2 = sm + maxnz;
  const i64 nzl = nz;
  const int tid = (int)cl.block_rank() * 256 + threadIdx.x, T = 8 * 256;
  for (int k = 0; k < nz; k++) {
    double *qk = Q + tri(k);
    for (int m = threadIdx.x; m <= k; m += 256) sqv1[m] = qk[m];          // Gram row k
    __syncthreads();
    for (int r = tid; r < k; r += T)             // sum over j <= r of Q(r, j) * sqv1[j]
      g2[r] = chain_walk<16>(Q2, (i64)r, 0, r + 1, sqv1, [nzl](int j) { return nzl - j - 1; });
    cl.sync();
    for (int m = threadIdx.x; m < k; m += 256) sqv2[m] = __ldcg(g2 + m);
    __syncthreads();
    for (int r = tid; r < k; r += T) gy[r] = chain_col<16>(Q, r, r, k, sqv2);
    cl.sync();
    for (int m = threadIdx.x; m < k; m += 256) sqv2[m] = __ldcg(gy + m);  // unscaled q_k
    __syncthreads();
    if (threadIdx.x == 0) {                       // one thread per block forms the recurrence
      const double alpha = chain_dot<16, true>(sqv1[k], sqv1, sqv2, k);
      alpha_sh = -1.0 / sqrt(alpha);
    }
    __syncthreads();
    const double alpha = alpha_sh;
    for (int m = tid; m <= k; m += T) {
      const double q = m < k ? sqv2[m] * alpha : -alpha;
      qk[m] = q;
      Q2[(i64)m * nzl - (i64)m * (m - 1) / 2 + (k - m)] = q;
    }
    cl.sync();
  }
}

static void set_smem(const void *fn, size_t bytes) {
  CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}

void build_q_store(QStore &qs, const Csr &Wt, const Csr &At) {
  q_offsets(qs, Wt);
  const int n = Wt.rn;
  if (n == 0) return;
  // bin the columns by support size: <=8 | <=32 | <=64 | <=144 | larger.  The block size and the
  // shared-memory footprint follow the bin, so that small supports do not pay the occupancy of
  // the largest one.
  constexpr int NBIN = 6;
  Buf<int> lists((i64)NBIN * n), cnt(NBIN);
  cnt.zero();
  const int *wro = Wt.ro.p;
  int *lp = lists.p, *cp = cnt.p;
  // several ranks: this rank builds the Q blocks of its column range, the blocks are exchanged
  std::vector<int> cs;
  std::vector<i64> off;
  const bool part = q_partition(qs, n, cs, off);
  const int c0 = part ? cs[(size_t)comm_rank()] : 0, c1 = part ? cs[(size_t)comm_rank() + 1] : n;
  const bool small = test_small_bins();     // test hook: the HBM and cluster builders from 9 rows on
  parallel_for(c1 - c0, [=] DEV(i64 ii) {
    const i64 i = c0 + ii;
    const int nz = wro[i + 1] - wro[i];
    if (nz == 0) return;
    const int bin = nz <= 8 ? 0 : small ? (nz <= 12 ? 4 : 5) : nz <= 32 ? 1 : nz <= 64 ? 2 : nz <= 144 ? 3 : nz <= 256 ? 4 : 5;
    const int p = atomic_add(&cp[bin], 1);
    lp[(i64)bin * n + p] = (int)i;
  });
  std::vector<int> hc = cnt.download();
  Context &c = ctx();
  const int *wcol = Wt.col.p, *aro = At.ro.p, *acol = At.col.p;
  const double *aa = At.a.p;
  static bool attr = false;
  if (!attr) { set_smem((const void *)k_build_q_block<true>, sizeof(double) * (2 * 144 + tri(144))); attr = true; }
  {
    static int logit = -1;
    if (logit < 0) { const char *e = getenv("AMGB_Q_LOG"); logit = (e && *e && *e != '0') ? 1 : 0; }
    if (logit) fprintf(stderr, "build_q: columns %d  maxnz %d  Q entries %lld | bins(<=8,32,64,144,256,more) %d %d %d %d %d %d\n",
                       n, qs.maxnz, (long long)qs.total, hc[0], hc[1], hc[2], hc[3], hc[4], hc[5]);
  }
  // columns with more than 256 rows: a handful (the column-0 pile of the reference) goes to the
  // cluster kernel on a second stream, next to the other bins; many of them are better off as one
  // block each, like the bin below
  const bool big_cluster = hc[5] > 0 && hc[5] <= 32;
  static cudaStream_t aux = nullptr;
  static cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  Buf<double> scratch;
  if (hc[5] && qs.maxnz > 12800) throw Error(-12, "interpolation support of " + std::to_string(qs.maxnz) + " rows exceeds the kernel limit (12800)");
  if (hc[5]) {
    k_gram_fill_big<<<dim3(64, hc[5]), 256, 0, c.stream>>>(lp + 5 * (i64)n, wro, wcol, aro, acol, aa, qs.Q.p, qs.qoff.p);
    c.launches++; post_launch("gram_fill_big");
  }
  if (big_cluster) {
    if (!aux) {
      CUDA_CHECK(cudaStreamCreateWithFlags(&aux, cudaStreamNonBlocking));
      CUDA_CHECK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
      CUDA_CHECK(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
    }
    scratch.alloc((i64)hc[5] * (2 * (i64)qs.maxnz + tri(qs.maxnz)));
    const size_t sm = sizeof(double) * 2 * (size_t)qs.maxnz;
    static size_t sm_set = 48 * 1024;
    if (sm > sm_set) { set_smem((const void *)k_build_q_cluster, sm); sm_set = sm; }
    CUDA_CHECK(cudaEventRecord(ev_fork, c.stream));
    CUDA_CHECK(cudaStreamWaitEvent(aux, ev_fork, 0));
    k_build_q_cluster<<<hc[5] * 8, 256, sm, aux>>>(lp + 5 * (i64)n, hc[5], qs.maxnz, wro, qs.Q.p, qs.qoff.p, scratch.p);
    c.launches++; post_launch("build_q_cluster");
    CUDA_CHECK(cudaEventRecord(ev_join, aux));
  } else if (hc[5]) {
    const size_t sm = sizeof(double) * 2 * (size_t)qs.maxnz;
    static size_t sm_set = 48 * 1024;
    if (sm > sm_set) { set_smem((const void *)k_build_q_block<false>, sm); sm_set = sm; }
    k_build_q_block<false><<<hc[5], 256, sm, c.stream>>>(lp + 5 * (i64)n, hc[5], qs.maxnz, wro, wcol, aro, acol, aa, qs.Q.p, qs.qoff.p);
    c.launches++; post_launch("build_q_global_big");
  }
  for (int b = 0; b < 5; b++) {
    if (!hc[b]) continue;
    k_gram_fill<<<(hc[b] + 7) / 8, 256, 0, c.stream>>>(lp + (i64)b * n, hc[b], wro, wcol, aro, acol, aa, qs.Q.p, qs.qoff.p);
    c.launches++; post_launch("gram_fill");
  }
  if (hc[0]) {
    k_build_q_tile8<8><<<(hc[0] + 31) / 32, 256, 0, c.stream>>>(lp, hc[0], wro, wcol, aro, acol, aa, qs.Q.p, qs.qoff.p);
    c.launches++; post_launch("build_q_tile8");
  }
  const int caps[3] = {32, 64, 144}, threads[3] = {32, 64, 128};
  for (int b = 0; b < 3; b++) {
    if (!hc[b + 1]) continue;
    const size_t sm = sizeof(double) * (2 * caps[b] + tri(caps[b]));
    k_build_q_block<true><<<hc[b + 1], threads[b], sm, c.stream>>>(lp + (i64)(b + 1) * n, hc[b + 1], caps[b], wro, wcol, aro,
                                                                   acol, aa, qs.Q.p, qs.qoff.p);
    c.launches++; post_launch("build_q_block");
  }
  if (hc[4]) {
    const size_t sm = sizeof(double) * (2 * (size_t)256);
    k_build_q_block<false><<<hc[4], 256, sm, c.stream>>>(lp + 4 * (i64)n, hc[4], 256, wro, wcol, aro, acol, aa, qs.Q.p, qs.qoff.p);
    c.launches++; post_launch("build_q_global");
  }
  if (big_cluster) CUDA_CHECK(cudaStreamWaitEvent(c.stream, ev_join, 0));   // join before anything reads Q
  if (part) comm_allgatherv(qs.Q.p, off.data(), "comm.q_blocks");
}

// ---- apply: W row := Q (Q^t (R(B e_i + u_i lambda))) ----
template <int G>
__global__ void __launch_bounds__(256) k_apply_q(int n, int nzcap, const int *wro, const int *wcol, double *wa,
                                                 const int *bro, const int *bcol, const double *ba,
                                                 const double *u, const double *lambda, const double *Qall,
                                                 const i64 *qoff) {
  extern __shared__ double sm[];
  auto tile = cg::tiled_partition<G>(cg::this_thread_block());
  const int slot = threadIdx.x / G, per = blockDim.x / G;
  const int i = blockIdx.x * per + slot;
  if (i >= n) return;
  const int b = wro[i], nz = wro[i + 1] - b;
  if (nz == 0) return;
  const int *Qj = wcol + b;
  const double *Q = Qall + qoff[i];
  double *sqv1 = sm + (size_t)slot * 2 * nzcap, *sqv2 = sqv1 + nzcap;
  const int r0 = tile.thread_rank();
  const int bb = bro[i], bn = bro[i + 1] - bb;
  const double ui = u[i];
  for (int k = r0; k < nz; k += G) {
    const double v = row_at(bcol + bb, ba + bb, bn, Qj[k]);
    sqv1[k] = v + ui * lambda[Qj[k]];
  }
  tile.sync();
  for (int r = r0; r < nz; r += G) {
    double v = 0;
    const double *uu = Q + tri(r);
    for (int j = 0; j <= r; j++) v = v + uu[j] * sqv1[j];
    sqv2[r] = v;
  }
  tile.sync();
  for (int r = r0; r < nz; r += G) {
    double y = 0;
    for (int j = r; j < nz; j++) y = y + Q[tri(j) + r] * sqv2[j];
    wa[b + r] = y;
  }
}

void apply_q(const QStore &qs, Csr &Wt, const Csr &Bt, const double *u, const double *lambda) {
  const int n = Wt.rn;
  if (n == 0) return;
  Context &c = ctx();
  const int nzcap = qs.maxnz > 0 ? qs.maxnz : 1;
  if (qs.maxnz <= 16) {
    const int per = 256 / 8;
    const size_t sm = sizeof(double) * (size_t)per * 2 * nzcap;
    k_apply_q<8><<<(n + per - 1) / per, 256, sm, c.stream>>>(n, nzcap, Wt.ro.p, Wt.col.p, Wt.a.p, Bt.ro.p, Bt.col.p,
                                                             Bt.a.p, u, lambda, qs.Q.p, qs.qoff.p);
  } else {
    int threads = 256;
    size_t sm = sizeof(double) * (size_t)(threads / 32) * 2 * nzcap;
    while (sm > 160 * 1024 && threads > 32) { threads /= 2; sm = sizeof(double) * (size_t)(threads / 32) * 2 * nzcap; }
    if (sm > 48 * 1024) set_smem((const void *)k_apply_q<32>, sm);
    const int per = threads / 32;
    k_apply_q<32><<<(n + per - 1) / per, threads, sm, c.stream>>>(n, nzcap, Wt.ro.p, Wt.col.p, Wt.a.p, Bt.ro.p,
                                                                  Bt.col.p, Bt.a.p, u, lambda, qs.Q.p, qs.qoff.p);
  }
  c.launches++; post_launch("apply_q");
}

// ---- QQ^t: one block per column (one warp for small ones), pairs (m<=j) over threads ----
__global__ void __launch_bounds__(128) k_form_qq(int n, const int *wro, const double *Qall, const i64 *qoff,
                                                 double *QQ, const i64 *qqoff) {
  const int i = blockIdx.x;
  if (i >= n) return;
  const int nz = wro[i + 1] - wro[i];
  if (nz > 192) return;                       // done by k_form_qq_big
  const double *Q = Qall + qoff[i];
  double *out = QQ + qqoff[i];
  for (int idx = threadIdx.x; idx < nz * nz; idx += blockDim.x) {
    const int m = idx / nz, j = idx - m * nz;
    if (m > j) continue;
    double acc = 0;
    for (int k = j; k < nz; k++) acc = acc + Q[tri(k) + m] * Q[tri(k) + j];
    out[(i64)m * nz + j] = acc;
    out[(i64)j * nz + m] = acc;
  }
}
// tiny columns: one thread group of 8 per column
__global__ void __launch_bounds__(256) k_form_qq_small(int n, const int *wro, const double *Qall, const i64 *qoff,
                                                       double *QQ, const i64 *qqoff) {
  const int i = blockIdx.x * 32 + threadIdx.x / 8;
  if (i >= n) return;
  const int nz = wro[i + 1] - wro[i];
  const double *Q = Qall + qoff[i];
  double *out = QQ + qqoff[i];
  for (int idx = threadIdx.x & 7; idx < nz * nz; idx += 8) {
    const int m = idx / nz, j = idx - m * nz;
    if (m > j) continue;
    double acc = 0;
    for (int k = j; k < nz; k++) acc = acc + Q[tri(k) + m] * Q[tri(k) + j];
    out[(i64)m * nz + j] = acc;
    out[(i64)j * nz + m] = acc;
  }
}
// columns with a large support: the pairs (m <= j) of one column are spread over gridDim.x blocks
__global__ void __launch_bounds__(256) k_form_qq_big(const int *list, const int *wro, const double *Qall,
                                                     const i64 *qoff, double *QQ, const i64 *qqoff) {
  const int i = list[blockIdx.y];
  const int nz = wro[i + 1] - wro[i];
  const double *Q = Qall + qoff[i];
  double *out = QQ + qqoff[i];
  const i64 total = (i64)nz * nz;
  for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (i64)gridDim.x * blockDim.x) {
    const int m = (int)(idx / nz), j = (int)(idx - (i64)m * nz);
    if (m > j) continue;
    double acc = 0;
    for (int k = j; k < nz; k++) acc = acc + Q[tri(k) + m] * Q[tri(k) + j];
    out[(i64)m * nz + j] = acc;
    out[(i64)j * nz + m] = acc;
  }
}
#define QQ_BIG 192
void form_qq(QQStore &qq, const QStore &qs, const Csr &Wt) {
  qq_offsets(qq, Wt);
  const int n = Wt.rn;
  if (n == 0) return;
  Context &c = ctx();
  if (qs.maxnz > QQ_BIG) {
    // the few columns with a very large support (the reference piles every F row without a
    // coupling into column 0, :2229) would otherwise be one block each and set the run time
    Buf<int> big(n), nbig(1);
    nbig.zero();
    const int *wro = Wt.ro.p;
    int *bp = big.p, *np_ = nbig.p;
    parallel_for(n, [=] DEV(i64 i) { if (wro[i + 1] - wro[i] > QQ_BIG) bp[atomic_add(np_, 1)] = (int)i; });
    const int nb = nbig.get(0);
    if (nb) {
      k_form_qq_big<<<dim3(96, nb), 256, 0, c.stream>>>(big.p, Wt.ro.p, qs.Q.p, qs.qoff.p, qq.QQ.p, qq.qqoff.p);
      c.launches++; post_launch("form_qq_big");
    }
  }
  if (qs.maxnz <= 12)
    k_form_qq_small<<<(n + 31) / 32, 256, 0, c.stream>>>(n, Wt.ro.p, qs.Q.p, qs.qoff.p, qq.QQ.p, qq.qqoff.p);
  else
    k_form_qq<<<n, 128, 0, c.stream>>>(n, Wt.ro.p, qs.Q.p, qs.qoff.p, qq.QQ.p, qq.qqoff.p);
  c.launches++; post_launch("form_qq");
}

// ---- S accumulation: G threads per row of S; coarse points in ascending order ----
template <int G>
__global__ void __launch_bounds__(256) k_lmop_acc(int nrows, const int *sro, const int *scol, double *sa,
                                                  const int *kro, const int *kcol, const int *tro,
                                                  const int *tcol, const int *tpos, const double *u,
                                                  const double *QQ, const i64 *qqoff) {
  auto tile = cg::tiled_partition<G>(cg::this_thread_block());
  const int j = blockIdx.x * (blockDim.x / G) + threadIdx.x / G;
  if (j >= nrows) return;
  const int r0 = tile.thread_rank();
  const int yb = sro[j], yn = sro[j + 1] - yb;
  for (int q = r0; q < yn; q += G) sa[yb + q] = 0.0;
  if (yn == 0) return;
  tile.sync();
  for (int e = kro[j]; e < kro[j + 1]; e++) {
    const int i = kcol[e];
    const int b = tro[i], nz = tro[i + 1] - b, k = tpos[e] - b;
    const double ui = u[i];
    const double *x = QQ + qqoff[i] + (i64)k * nz;
    for (int kk = r0; kk < nz; kk += G) {
      const int p = pos_of(scol + yb, yn, tcol[b + kk]);
      if (p >= 0) sa[yb + p] = sa[yb + p] + ui * x[kk];
    }
    tile.sync();
  }
}
void lmop_accumulate(Csr &S, const QQStore &qq, const double *u, const Csr &Wskt, const Csr &Wsk,
                     const int *tpos) {
  const int n = S.rn;
  if (n == 0) return;
  Context &c = ctx();
  const double avg = (double)S.nnz / (double)n;
  if (avg <= 24.0) {
    const int per = 256 / 8;
    k_lmop_acc<8><<<(n + per - 1) / per, 256, 0, c.stream>>>(n, S.ro.p, S.col.p, S.a.p, Wsk.ro.p, Wsk.col.p, Wskt.ro.p,
                                                             Wskt.col.p, tpos, u, qq.QQ.p, qq.qqoff.p);
  } else {
    const int per = 256 / 32;
    k_lmop_acc<32><<<(n + per - 1) / per, 256, 0, c.stream>>>(n, S.ro.p, S.col.p, S.a.p, Wsk.ro.p, Wsk.col.p, Wskt.ro.p,
                                                              Wskt.col.p, tpos, u, qq.QQ.p, qq.qqoff.p);
  }
  c.launches++; post_launch("lmop_acc");
}
#endif

}  // namespace amgb

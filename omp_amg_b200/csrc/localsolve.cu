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
      else {
        // the Gram row is sparse (a few dozen of its k entries): a product with an exact zero adds a
        // signed zero, which changes no partial sum, so only the non-zeros enter the dependent chain
        for (int j = 0; j <= r; j++) { const double s = sqv1[j]; if (s != 0.0) v = v + u[j] * s; }
      }
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
    else for (int m = 0; m < k; m++) { const double s = sqv1[m]; if (s != 0.0) alpha = alpha - s * qk[m]; }
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

// Gram rows of columns with a very large support: blockIdx.x = column (the x dimension of a grid
// has no 65535 limit: a 4 M-row Q1 mesh has more such columns than that), the pairs (k, m <= k)
// are spread over gridDim.y blocks
__global__ void __launch_bounds__(256) k_gram_fill_big(const int *list, const int *wro, const int *wcol, const int *aro,
                                                       const int *acol, const double *aa, double *Qall, const i64 *qoff) {
  const int i = list[blockIdx.x];
  const int b = wro[i], nz = wro[i + 1] - b;
  const int *Qj = wcol + b;
  double *Q = Qall + qoff[i];
  for (int k = blockIdx.y * 8 + (threadIdx.x >> 5); k < nz; k += gridDim.y * 8) {
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

// =======================================================================================
// Huge supports: blocked, right-looking A-orthogonalisation over the whole GPU.
//
// The reference piles every F row without a coupling to a C row into column 0 (min_skel :2229):
// one coarse column with a support of 1050 rows at 128^3 Poisson, of 18 568 rows at 64^3
// anisotropic diffusion.  Step k of the reference (:2083-2098) is
//     sqv2[j] = sum_{m<=j} Q_j[m] G[k][m]              (G = Gram rows of A on the support, sparse)
//     q_k[r]  = sum_{j=r}^{k-1} Q_j[r] sqv2[j]          (ascending j, starting from 0)
//     alpha   = G[k][k] - sum_{m<k} G[k][m] q_k[m];  q_k *= -1/sqrt(alpha);  q_k[k] = 1/sqrt(alpha)
// i.e. O(k^2) work with O(k) dependent additions per step, O(nz^3/6) in all.  Two facts make it a
// blocked algorithm without touching a single rounding:
//   * sqv2[j] of step k depends on row j of Q and on G only, so S2[k][j] can be formed as soon as
//     row j is final, for all k > j at once;
//   * q_k[r] receives its addends in ascending j, so after row j is final the update
//     Y[k][r] += Q_j[r] S2[k][j] may be applied to every later row k -- each element still sees its
//     addends in ascending j, one rounded product and one rounded addition each.
// Rows are processed in panels of b: inside a panel the rows are finished one after the other by
// a cooperative kernel (one grid barrier per row: a row applies the pending updates of its own
// panel lazily), then S2 of the panel against all later rows is formed (sparse dots) and the
// trailing rows take the b updates of the panel in one register-tiled pass (k_q_trail: the bulk of
// the work, a GEMM-shaped loop with the accumulation order fixed).  Exact zeros of G are skipped:
// adding a (signed) zero product changes no partial sum.
// =======================================================================================
struct LocalGram { Buf<int> ro, col; Buf<double> a; };   // lower triangle incl. diagonal, local indices

__global__ void __launch_bounds__(256) k_lg_count(int nz, const int *Qj, const int *aro, const int *acol, int *cnt) {
  const int k = blockIdx.x * 256 + threadIdx.x;
  if (k >= nz) return;
  const int s = Qj[k];
  int c = 0;
  for (int e = aro[s]; e < aro[s + 1]; e++) { const int m = pos_of(Qj, k + 1, acol[e]); c += (m >= 0); }
  cnt[k] = c;
}
__global__ void __launch_bounds__(256) k_lg_fill(int nz, const int *Qj, const int *aro, const int *acol, const double *aa,
                                                 const int *gro, int *gcol, double *ga) {
  const int k = blockIdx.x * 256 + threadIdx.x;
  if (k >= nz) return;
  const int s = Qj[k];
  int p = gro[k];
  for (int e = aro[s]; e < aro[s + 1]; e++) {
    const int m = pos_of(Qj, k + 1, acol[e]);
    if (m >= 0) { gcol[p] = m; ga[p] = aa[e]; p++; }
  }
}

#define QP_BMAX 64
// rows [j0, j1) of one column; cooperative launch.  Inside the panel the rows are kept UNSCALED
// with their factor -1/sqrt(alpha) aside (every block holds the factors in shared memory): a read
// multiplies on the fly, which is the same rounded product the reference stores.
// ordered sparse dot by one warp: the lanes fetch their terms side by side (the loads of a sparse
// dot are independent of the running sum; issued one after the other by a single thread each would
// pay its own L2 round trip), lane 0's order of additions is the entry order
__device__ __forceinline__ double warp_sparse_dot(double v0, bool sub, int e0, int e1, const int *gcol, const double *ga,
                                                  const double *q, double f, int diag, int mlimit) {
  const int lane = threadIdx.x & 31;
  double v = v0;
  for (int base = e0; base < e1; base += 32) {
    const int e = base + lane;
    double p = 0.0;
    bool ok = false;
    if (e < e1) {
      const int m = gcol[e];
      if (m <= mlimit) { ok = true; const double qq = (m == diag) ? -f : __ldcg(q + m) * f; p = qq * ga[e]; }
    }
    const unsigned b = __ballot_sync(0xffffffffu, ok);
    const int cnt = __popc(b);                 // the valid entries are a prefix (columns ascending)
    for (int l = 0; l < cnt; l++) { const double pl = __shfl_sync(0xffffffffu, p, l); v = sub ? v - pl : v + pl; }
    if (cnt < 32) break;
  }
  return v;
}

__global__ void __launch_bounds__(256) k_q_panel(int j0, int j1, double *Q, const int *gro, const int *gcol,
                                                 const double *ga) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sfac[QP_BMAX], ss2[QP_BMAX];
  const int t = threadIdx.x, tid = blockIdx.x * 256 + t, T = gridDim.x * 256, w = t >> 5;
  for (int j = j0; j <= j1; j++) {
    // factor of the row finished in the previous step (its elements are complete after the barrier):
    // alpha = G[k][k] - sum_{m<k} G[k][m] y[m]  (the rows of the panel are kept unscaled: f = 1)
    if (j > j0) {
      if (w == 0) {
        const int k = j - 1;
        const int e0 = gro[k], e1 = gro[k + 1];
        double alpha = 0.0;
        if (e1 > e0 && gcol[e1 - 1] == k) alpha = ga[e1 - 1];
        alpha = warp_sparse_dot(alpha, true, e0, e1, gcol, ga, Q + tri(k), 1.0, -1, k - 1);
        if (t == 0) sfac[k - j0] = -1.0 / sqrt(alpha);
      }
      __syncthreads();
    }
    if (j == j1) break;
    const int np = j - j0;
    // S2[j][jp] for the finished rows jp of this panel (every block forms all of them: no exchange)
    for (int jj = w; jj < np; jj += 8) {
      const int jp = j0 + jj;
      const double v = warp_sparse_dot(0.0, false, gro[j], gro[j + 1], gcol, ga, Q + tri(jp), sfac[jj], jp, jp);
      if ((t & 31) == 0) ss2[jj] = v;
    }
    __syncthreads();
    // pending updates of this panel onto row j (earlier panels were applied by their trailing pass);
    // the eight loads of a batch are issued before the first of them is used
    double *qj = Q + tri(j);
    for (int r = tid; r < j; r += T) {
      double y = qj[r];
      for (int jj = r > j0 ? r - j0 : 0; jj < np; jj += 8) {
        double q[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const int jx = jj + u < np ? jj + u : np - 1, jp = j0 + jx;
          q[u] = (r == jp) ? 0.0 : __ldcg(Q + tri(jp) + r);
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
          if (jj + u < np) {
            const double f = sfac[jj + u];
            const double qq = (r == j0 + jj + u) ? -f : q[u] * f;
            y = y + qq * ss2[jj + u];
          }
        }
      }
      qj[r] = y;
    }
    grid.sync();
  }
  // scale the panel rows in place: Q is final from here on
  for (int jj = 0; jj < j1 - j0; jj++) {
    const int k = j0 + jj;
    double *qk = Q + tri(k);
    const double f = sfac[jj];
    for (int m = tid; m <= k; m += T) qk[m] = (m == k) ? -f : qk[m] * f;
  }
}

// S2 of the panel against the later rows: s2p[(k - j1) * bb + jj] = sum_{m <= j0+jj} Q_{j0+jj}[m] G[k][m]
__global__ void __launch_bounds__(256) k_q_s2(int nz, int j0, int j1, const double *Q, const int *gro, const int *gcol,
                                              const double *ga, double *s2p) {
  const int bb = j1 - j0;
  const i64 total = (i64)(nz - j1) * bb;
  for (i64 idx = (i64)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (i64)gridDim.x * 256) {
    const int k = j1 + (int)(idx / bb), jj = (int)(idx % bb), jp = j0 + jj;
    const double *qp = Q + tri(jp);
    double v = 0.0;
    for (int e = gro[k]; e < gro[k + 1]; e++) {
      const int m = gcol[e];
      if (m > jp) break;
      v = v + qp[m] * ga[e];
    }
    s2p[idx] = v;
  }
}

// trailing update: Y[k][r] += sum over the panel rows jp >= r (ascending) of Q_jp[r] * S2[k][jp],
// k in [j1, nz), r in [0, j1).  64 x 64 tiles, 4 x 4 per thread, the panel staged in shared memory.
__global__ void __launch_bounds__(256) k_q_trail(int nz, int j0, int j1, double *Q, const double *s2p) {
  extern __shared__ __align__(16) double trail_sm[];
  double (*sQ)[64 + 4] = reinterpret_cast<double (*)[64 + 4]>(trail_sm);
  double (*sS)[QP_BMAX + 1] = reinterpret_cast<double (*)[QP_BMAX + 1]>(trail_sm + QP_BMAX * (64 + 4));
  const int bb = j1 - j0;
  const int r0 = blockIdx.x * 64, k0 = j1 + blockIdx.y * 64;
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  for (int idx = t; idx < bb * 64; idx += 256) {
    const int jj = idx / 64, c = idx - jj * 64, r = r0 + c, jp = j0 + jj;
    sQ[jj][c] = (r <= jp) ? Q[tri(jp) + r] : 0.0;
  }
  for (int idx = t; idx < 64 * bb; idx += 256) {
    const int kk = idx / bb, jj = idx - kk * bb, k = k0 + kk;
    sS[kk][jj] = (k < nz) ? s2p[(i64)(k - j1) * bb + jj] : 0.0;
  }
  __syncthreads();
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int c = 0; c < 4; c++) {
      const int k = k0 + ty * 4 + a, r = r0 + tx * 4 + c;
      acc[a][c] = (k < nz && r < j1) ? Q[tri(k) + r] : 0.0;
    }
  const bool guard = (r0 + 63 > j0);          // some columns of the tile start inside the panel
  for (int jj = 0; jj < bb; jj++) {
    double q[4], s[4];
#pragma unroll
    for (int c = 0; c < 4; c++) q[c] = sQ[jj][tx * 4 + c];
#pragma unroll
    for (int a = 0; a < 4; a++) s[a] = sS[ty * 4 + a][jj];
    if (!guard) {
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[a][c] = acc[a][c] + q[c] * s[a];
    } else {
      const int jp = j0 + jj;
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int c = 0; c < 4; c++)
          if (r0 + tx * 4 + c <= jp) acc[a][c] = acc[a][c] + q[c] * s[a];
    }
  }
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int c = 0; c < 4; c++) {
      const int k = k0 + ty * 4 + a, r = r0 + tx * 4 + c;
      if (k < nz && r < j1) Q[tri(k) + r] = acc[a][c];
    }
}

// Q of one huge column, in place in the store (which must be zero on entry)
static void build_q_huge(double *Q, int nz, const int *Qj, const Csr &At, int panel) {
  Context &c = ctx();
  LocalGram lg;
  Buf<int> cnt(nz + 1);
  k_lg_count<<<(nz + 255) / 256, 256, 0, c.stream>>>(nz, Qj, At.ro.p, At.col.p, cnt.p);
  c.launches++; post_launch("lg_count");
  lg.ro.alloc(nz + 1);
  const i64 gn = exclusive_scan(cnt.p, lg.ro.p, nz);
  lg.col.alloc(gn); lg.a.alloc(gn);
  k_lg_fill<<<(nz + 255) / 256, 256, 0, c.stream>>>(nz, Qj, At.ro.p, At.col.p, At.a.p, lg.ro.p, lg.col.p, lg.a.p);
  c.launches++; post_launch("lg_fill");
  dev_memset(Q, 0, sizeof(double) * (size_t)tri(nz));
  Buf<double> s2p((i64)nz * panel);
  static int coop_blocks = 0;
  if (!coop_blocks) {
    int nb = 0;
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_q_panel, 256, 0));
    coop_blocks = std::min(c.sm_count * std::max(nb, 1), 64);
  }
  const int *gro = lg.ro.p, *gcol = lg.col.p;
  const double *ga = lg.a.p;
  for (int j0 = 0; j0 < nz; j0 += panel) {
    int j1 = std::min(nz, j0 + panel);
    void *args[] = {(void *)&j0, (void *)&j1, (void *)&Q, (void *)&gro, (void *)&gcol, (void *)&ga};
    CUDA_CHECK(cudaLaunchCooperativeKernel((const void *)k_q_panel, dim3(coop_blocks), dim3(256), args, 0, c.stream));
    c.launches++; post_launch("q_panel");
    if (j1 >= nz) break;
    const i64 total = (i64)(nz - j1) * (j1 - j0);
    k_q_s2<<<(unsigned)std::min<i64>((total + 255) / 256, (i64)c.sm_count * 16), 256, 0, c.stream>>>(nz, j0, j1, Q, gro, gcol, ga, s2p.p);
    c.launches++; post_launch("q_s2");
    constexpr size_t trail_bytes = sizeof(double) * (QP_BMAX * (64 + 4) + 64 * (QP_BMAX + 1));
    static bool trail_attr = false;
    if (!trail_attr) { CUDA_CHECK(cudaFuncSetAttribute((const void *)k_q_trail, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)trail_bytes)); trail_attr = true; }
    k_q_trail<<<dim3((j1 + 63) / 64, (nz - j1 + 63) / 64), 256, trail_bytes, c.stream>>>(nz, j0, j1, Q, s2p.p);
    c.launches++; post_launch("q_trail");
  }
}

// ---- apply for a huge column: y = Q^t x (rows of the triangle), w = Q y (columns) ----
__global__ void __launch_bounds__(256) k_apply_huge_rhs(int nz, const int *Qj, int i, const int *bro, const int *bcol,
                                                        const double *ba, const double *u, const double *lambda,
                                                        double *x) {
  const int k = blockIdx.x * 256 + threadIdx.x;
  if (k >= nz) return;
  const int bb = bro[i], bn = bro[i + 1] - bb;
  const double v = row_at(bcol + bb, ba + bb, bn, Qj[k]);
  x[k] = v + u[i] * lambda[Qj[k]];
}
__global__ void __launch_bounds__(256) k_apply_huge_utt(int nz, const double *Q, const double *x, double *y) {
  // one warp per row r: the lanes stage 32 products at a time, lane 0's order is the row's order
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= nz) return;
  const double *u = Q + tri(r);
  double v = 0;
  for (int base = 0; base <= r; base += 32) {
    const int j = base + lane;
    const double p = (j <= r) ? u[j] * x[j] : 0.0;
    const int m = min(32, r + 1 - base);
    for (int l = 0; l < m; l++) v = v + __shfl_sync(0xffffffffu, p, l);
  }
  if (lane == 0) y[r] = v;
}
__global__ void __launch_bounds__(256) k_apply_huge_ut(int nz, const double *Q, const double *y, double *w) {
  const int r = blockIdx.x * 256 + threadIdx.x;
  if (r >= nz) return;
  double acc = 0;
  for (int j = r; j < nz; j++) acc = acc + Q[tri(j) + r] * y[j];
  w[r] = acc;
}

// ---- Q Q^t of a huge column: out[m][j] = sum_{k >= max(m,j)} q_k[m] q_k[j], k ascending ----
// 64 x 64 tiles of the upper triangle (mirrored on write), 4 x 4 per thread, 32 rows of Q per stage
__global__ void __launch_bounds__(256) k_form_qq_huge(int nz, const double *Q, double *out) {
  __shared__ double sA[32][64 + 4], sB[32][64 + 4];
  const int m0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
  if (m0 > j0) return;                                 // lower tiles come from the mirror
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int c = 0; c < 4; c++) acc[a][c] = 0.0;
  for (int kb = j0; kb < nz; kb += 32) {              // every pair of the tile starts at k >= j >= j0
    __syncthreads();
    for (int idx = t; idx < 32 * 64; idx += 256) {
      const int kk = idx / 64, c = idx - kk * 64, k = kb + kk;
      sA[kk][c] = (k < nz && m0 + c <= k) ? Q[tri(k) + m0 + c] : 0.0;
      sB[kk][c] = (k < nz && j0 + c <= k) ? Q[tri(k) + j0 + c] : 0.0;
    }
    __syncthreads();
    const int kend = min(32, nz - kb);
    const bool guard = (kb < j0 + 64);                // pairs (m, j) start at k = max(m, j) = j here (m <= j in upper tiles... or m > j on the diagonal tile)
    for (int kk = 0; kk < kend; kk++) {
      const int k = kb + kk;
      double a4[4], b4[4];
#pragma unroll
      for (int a = 0; a < 4; a++) a4[a] = sA[kk][ty * 4 + a];
#pragma unroll
      for (int c = 0; c < 4; c++) b4[c] = sB[kk][tx * 4 + c];
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int c = 0; c < 4; c++) {
          if (guard) { const int m = m0 + ty * 4 + a, j = j0 + tx * 4 + c; if (k < (m > j ? m : j)) continue; }
          acc[a][c] = acc[a][c] + a4[a] * b4[c];
        }
    }
  }
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int c = 0; c < 4; c++) {
      const int m = m0 + ty * 4 + a, j = j0 + tx * 4 + c;
      if (m < nz && j < nz && (m0 < j0 || m <= j)) { out[(i64)m * nz + j] = acc[a][c]; out[(i64)j * nz + m] = acc[a][c]; }
    }
}

// which columns are huge, and the largest support among the others (one read-back)
struct HugeRec { int col, nz, wb, pad; long long qoff; };
__global__ void __launch_bounds__(256) k_find_huge(int n, int bignz, const int *wro, const i64 *qoff, HugeRec *rec, int cap, int *meta) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const int nz = wro[i + 1] - wro[i];
  if (nz > bignz) { const int p = atomicAdd(&meta[0], 1); if (p < cap) { rec[p].col = i; rec[p].nz = nz; rec[p].wb = wro[i]; rec[p].qoff = qoff[i]; } }
  else atomicMax(&meta[1], nz);
}
static void find_huge(QStore &qs, const Csr &Wt) {
  const bool force = test_force('p');
  qs.bignz = force ? 12 : 512;
  qs.huge.clear();
  qs.maxnz_small = qs.maxnz;
  if (qs.maxnz <= qs.bignz || Wt.rn == 0) return;
  const int cap = 8192;
  Buf<HugeRec> rec(cap);
  Buf<int> meta(2);
  meta.zero();
  Context &c = ctx();
  k_find_huge<<<(Wt.rn + 255) / 256, 256, 0, c.stream>>>(Wt.rn, qs.bignz, Wt.ro.p, qs.qoff.p, rec.p, cap, meta.p);
  c.launches++; post_launch("find_huge");
  std::vector<int> hm = meta.download();
  // The blocked kernels take one column at a time with the whole GPU: right for the pile (one or
  // two columns), wrong for a coarse level where hundreds of columns have supports of 600 rows
  // (Q1 vertex meshes: 30 s instead of 9 s per setup at 97^3).  With more than 8 candidates only
  // supports above 2048 rows count as huge; the others stay with one block each.
  if (hm[0] > cap) {
    if (force) throw Error(-12, "more than 8192 interpolation supports above " + std::to_string(qs.bignz) + " rows");
    qs.bignz = 0x7fffffff;
    return;
  }
  std::vector<HugeRec> hr((size_t)hm[0]);
  if (hm[0]) d2h(hr.data(), rec.p, sizeof(HugeRec) * (size_t)hm[0]);
  int small_max = hm[1];
  if (!force && hm[0] > 8) {
    qs.bignz = 2048;
    std::vector<HugeRec> keep;
    for (const HugeRec &r : hr) { if (r.nz > qs.bignz) keep.push_back(r); else if (r.nz > small_max) small_max = r.nz; }
    hr.swap(keep);
  }
  std::sort(hr.begin(), hr.end(), [](const HugeRec &a, const HugeRec &b) { return a.col < b.col; });
  for (const HugeRec &r : hr) qs.huge.push_back(QStore::Huge{r.col, r.nz, r.wb, (i64)r.qoff});
  qs.maxnz_small = small_max;
}

static void set_smem(const void *fn, size_t bytes) {
  CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}

void build_q_store(QStore &qs, const Csr &Wt, const Csr &At) {
  q_offsets(qs, Wt);
  const int n = Wt.rn;
  if (n == 0) return;
  // bin the columns by support size.  The block size and the shared-memory footprint follow the
  // bin, so that small supports do not pay the occupancy of the largest one: the builders are
  // chains of dependent additions with five barriers per step, i.e. bound by latency, and what
  // hides latency is the number of columns resident per SM (a triangle of 144 rows is 86 KB: two
  // blocks per SM; of 96 rows 39 KB: five).
  //   0: <=8 (8-thread tiles) | 1..NSM: Q in shared memory, caps below | B_HBM: <=256, Q in HBM/L2 |
  //   B_BIG: larger
  constexpr int NSM = 7, B_HBM = NSM + 1, B_BIG = NSM + 2, NBIN = NSM + 3;
  const int caps[NSM] = {32, 48, 64, 80, 96, 112, 144}, threads[NSM] = {32, 64, 64, 128, 128, 128, 128};
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
  find_huge(qs, Wt);
  const int bignz = qs.bignz;
  parallel_for(c1 - c0, [=] DEV(i64 ii) {
    const i64 i = c0 + ii;
    const int nz = wro[i + 1] - wro[i];
    if (nz == 0 || nz > bignz) return;
    const int bin = nz <= 8 ? 0 : small ? (nz <= 12 ? B_HBM : B_BIG) : nz <= 32 ? 1 : nz <= 48 ? 2 : nz <= 64 ? 3 : nz <= 80 ? 4 :
                    nz <= 96 ? 5 : nz <= 112 ? 6 : nz <= 144 ? 7 : nz <= 256 ? B_HBM : B_BIG;
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
    if (logit) fprintf(stderr, "build_q: columns %d  maxnz %d  Q entries %lld | bins(<=8,32,48,64,80,96,112,144,256,more) %d %d %d %d %d %d %d %d %d %d\n",
                       n, qs.maxnz, (long long)qs.total, hc[0], hc[1], hc[2], hc[3], hc[4], hc[5], hc[6], hc[7], hc[8], hc[9]);
  }
  // columns with more than 256 rows: a handful (the column-0 pile of the reference) goes to the
  // cluster kernel on a second stream, next to the other bins; many of them are better off as one
  // block each, like the bin below
  const bool big_cluster = hc[B_BIG] > 0 && hc[B_BIG] <= 32;
  static cudaStream_t aux = nullptr;
  static cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  Buf<double> scratch;
  if (hc[B_BIG] && qs.maxnz_small > 12800) throw Error(-12, "interpolation support of " + std::to_string(qs.maxnz_small) + " rows exceeds the kernel limit (12800)");
  if (hc[B_BIG]) {
    k_gram_fill_big<<<dim3(hc[B_BIG], 64), 256, 0, c.stream>>>(lp + B_BIG * (i64)n, wro, wcol, aro, acol, aa, qs.Q.p, qs.qoff.p);
    c.launches++; post_launch("gram_fill_big");
  }
  if (big_cluster) {
    if (!aux) {
      CUDA_CHECK(cudaStreamCreateWithFlags(&aux, cudaStreamNonBlocking));
      CUDA_CHECK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
      CUDA_CHECK(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
    }
    scratch.alloc((i64)hc[B_BIG] * (2 * (i64)qs.maxnz_small + tri(qs.maxnz_small)));
    const size_t sm = sizeof(double) * 2 * (size_t)qs.maxnz_small;
    static size_t sm_set = 48 * 1024;
    if (sm > sm_set) { set_smem((const void *)k_build_q_cluster, sm); sm_set = sm; }
    CUDA_CHECK(cudaEventRecord(ev_fork, c.stream));
    CUDA_CHECK(cudaStreamWaitEvent(aux, ev_fork, 0));
    k_build_q_cluster<<<hc[B_BIG] * 8, 256, sm, aux>>>(lp + B_BIG * (i64)n, hc[B_BIG], qs.maxnz_small, wro, qs.Q.p, qs.qoff.p, scratch.p);
    c.launches++; post_launch("build_q_cluster");
    CUDA_CHECK(cudaEventRecord(ev_join, aux));
  } else if (hc[B_BIG]) {
    const size_t sm = sizeof(double) * 2 * (size_t)qs.maxnz_small;
    static size_t sm_set = 48 * 1024;
    if (sm > sm_set) { set_smem((const void *)k_build_q_block<false>, sm); sm_set = sm; }
    k_build_q_block<false><<<hc[B_BIG], 256, sm, c.stream>>>(lp + B_BIG * (i64)n, hc[B_BIG], qs.maxnz_small, wro, wcol, aro, acol, aa, qs.Q.p, qs.qoff.p);
    c.launches++; post_launch("build_q_global_big");
  }
  for (int b = 0; b < B_BIG; b++) {
    if (!hc[b]) continue;
    k_gram_fill<<<(hc[b] + 7) / 8, 256, 0, c.stream>>>(lp + (i64)b * n, hc[b], wro, wcol, aro, acol, aa, qs.Q.p, qs.qoff.p);
    c.launches++; post_launch("gram_fill");
  }
  if (hc[0]) {
    k_build_q_tile8<8><<<(hc[0] + 31) / 32, 256, 0, c.stream>>>(lp, hc[0], wro, wcol, aro, acol, aa, qs.Q.p, qs.qoff.p);
    c.launches++; post_launch("build_q_tile8");
  }
  for (int b = 0; b < NSM; b++) {
    if (!hc[b + 1]) continue;
    const size_t sm = sizeof(double) * (2 * caps[b] + tri(caps[b]));
    k_build_q_block<true><<<hc[b + 1], threads[b], sm, c.stream>>>(lp + (i64)(b + 1) * n, hc[b + 1], caps[b], wro, wcol, aro,
                                                                   acol, aa, qs.Q.p, qs.qoff.p);
    c.launches++; post_launch("build_q_block");
  }
  if (hc[B_HBM]) {
    const size_t sm = sizeof(double) * (2 * (size_t)256);
    k_build_q_block<false><<<hc[B_HBM], 256, sm, c.stream>>>(lp + B_HBM * (i64)n, hc[B_HBM], 256, wro, wcol, aro, acol, aa, qs.Q.p, qs.qoff.p);
    c.launches++; post_launch("build_q_global");
  }
  // huge supports: one column at a time, the whole GPU on each (this rank's columns only)
  for (const QStore::Huge &h : qs.huge) {
    if (h.col < c0 || h.col >= c1) continue;
    build_q_huge(qs.Q.p + h.qoff, h.nz, wcol + h.wb, At, test_force('p') ? 4 : QP_BMAX);
  }
  if (big_cluster) CUDA_CHECK(cudaStreamWaitEvent(c.stream, ev_join, 0));   // join before anything reads Q
  if (part) comm_allgatherv(qs.Q.p, off.data(), "comm.q_blocks");
}

// ---- apply: W row := Q (Q^t (R(B e_i + u_i lambda))) ----
template <int G>
__global__ void __launch_bounds__(256) k_apply_q(int n, int nzcap, const int *wro, const int *wcol, double *wa,
                                                 const int *bro, const int *bcol, const double *ba,
                                                 const double *u, const double *lambda, const double *Qall,
                                                 const i64 *qoff, int bignz) {
  extern __shared__ double sm[];
  auto tile = cg::tiled_partition<G>(cg::this_thread_block());
  const int slot = threadIdx.x / G, per = blockDim.x / G;
  const int i = blockIdx.x * per + slot;
  if (i >= n) return;
  const int b = wro[i], nz = wro[i + 1] - b;
  if (nz == 0 || nz > bignz) return;            // huge supports: k_apply_huge_*
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
  const int nzcap = qs.maxnz_small > 0 ? qs.maxnz_small : 1;
  for (const QStore::Huge &h : qs.huge) {
    Buf<double> xv(h.nz), yv(h.nz);
    k_apply_huge_rhs<<<(h.nz + 255) / 256, 256, 0, c.stream>>>(h.nz, Wt.col.p + h.wb, h.col, Bt.ro.p, Bt.col.p, Bt.a.p, u, lambda, xv.p);
    k_apply_huge_utt<<<(h.nz + 7) / 8, 256, 0, c.stream>>>(h.nz, qs.Q.p + h.qoff, xv.p, yv.p);
    k_apply_huge_ut<<<(h.nz + 255) / 256, 256, 0, c.stream>>>(h.nz, qs.Q.p + h.qoff, yv.p, Wt.a.p + h.wb);
    c.launches += 3; post_launch("apply_huge");
  }
  if (qs.maxnz_small <= 16) {
    const int per = 256 / 8;
    const size_t sm = sizeof(double) * (size_t)per * 2 * nzcap;
    k_apply_q<8><<<(n + per - 1) / per, 256, sm, c.stream>>>(n, nzcap, Wt.ro.p, Wt.col.p, Wt.a.p, Bt.ro.p, Bt.col.p,
                                                             Bt.a.p, u, lambda, qs.Q.p, qs.qoff.p, qs.bignz);
  } else {
    int threads = 256;
    size_t sm = sizeof(double) * (size_t)(threads / 32) * 2 * nzcap;
    while (sm > 160 * 1024 && threads > 32) { threads /= 2; sm = sizeof(double) * (size_t)(threads / 32) * 2 * nzcap; }
    if (sm > 48 * 1024) set_smem((const void *)k_apply_q<32>, sm);
    const int per = threads / 32;
    k_apply_q<32><<<(n + per - 1) / per, threads, sm, c.stream>>>(n, nzcap, Wt.ro.p, Wt.col.p, Wt.a.p, Bt.ro.p,
                                                                  Bt.col.p, Bt.a.p, u, lambda, qs.Q.p, qs.qoff.p, qs.bignz);
  }
  c.launches++; post_launch("apply_q");
}

// sum over k in [k0, nz) of Q[tri(k) + m] * Q[tri(k) + j], k ascending, with the sixteen loads of a
// batch issued before the first product is added (AMGB_QQ_BATCH=1; A/B switch)
__device__ __forceinline__ double qq_pair(const double *Q, int m, int j, int k0, int nz) {
  double acc = 0;
  for (int k = k0; k < nz; k += 8) {
    double a[8], b[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int kk = k + u < nz ? k + u : nz - 1;
      const i64 o = tri(kk);
      a[u] = Q[o + m]; b[u] = Q[o + j];
    }
#pragma unroll
    for (int u = 0; u < 8; u++) if (k + u < nz) acc = acc + a[u] * b[u];
  }
  return acc;
}
// p = tri(j) + m with 0 <= m <= j
__device__ __forceinline__ void tri_index(i64 p, int &j, int &m) {
  i64 jj = (i64)((sqrt(8.0 * (double)p + 1.0) - 1.0) * 0.5);
  while (tri(jj + 1) <= p) jj++;
  while (tri(jj) > p) jj--;
  j = (int)jj; m = (int)(p - tri(jj));
}
// ---- QQ^t: one block per column (one warp for small ones), pairs (m<=j) over threads ----
__global__ void __launch_bounds__(128) k_form_qq(int n, const int *wro, const double *Qall, const i64 *qoff,
                                                 double *QQ, const i64 *qqoff, int bignz, int batch) {
  const int i = blockIdx.x;
  if (i >= n) return;
  const int nz = wro[i + 1] - wro[i];
  if (nz > 192 || nz > bignz) return;         // done by k_form_qq_big / k_form_qq_huge
  const double *Q = Qall + qoff[i];
  double *out = QQ + qqoff[i];
  // the pairs (m <= j) enumerated along the packed triangle, p = tri(j) + m: neighbouring threads
  // share j (same chain length, Q[tri(k)+j] is a broadcast) and read neighbouring Q[tri(k)+m];
  // no thread idles on the m > j half of the square
  const int T = (int)tri(nz);
  for (int p = threadIdx.x; p < T; p += blockDim.x) {
    int j, m;
    tri_index(p, j, m);
    double acc = 0;
    if (batch) acc = qq_pair(Q, m, j, j, nz);
    else for (int k = j; k < nz; k++) acc = acc + Q[tri(k) + m] * Q[tri(k) + j];
    out[(i64)m * nz + j] = acc;
    out[(i64)j * nz + m] = acc;
  }
}
// tiny columns: one thread group of 8 per column
__global__ void __launch_bounds__(256) k_form_qq_small(int n, const int *wro, const double *Qall, const i64 *qoff,
                                                       double *QQ, const i64 *qqoff, int bignz) {
  const int i = blockIdx.x * 32 + threadIdx.x / 8;
  if (i >= n) return;
  const int nz = wro[i + 1] - wro[i];
  if (nz > bignz) return;
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
  const int i = list[blockIdx.x];          // x: no 65535 limit on the number of columns
  const int nz = wro[i + 1] - wro[i];
  const double *Q = Qall + qoff[i];
  double *out = QQ + qqoff[i];
  const i64 total = tri(nz);
  for (i64 idx = (i64)blockIdx.y * blockDim.x + threadIdx.x; idx < total; idx += (i64)gridDim.y * blockDim.x) {
    int j, m;
    tri_index(idx, j, m);
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
  for (const QStore::Huge &h : qs.huge) {
    const i64 qqo = qq.qqoff.get(h.col);
    const int nt = (h.nz + 63) / 64;
    k_form_qq_huge<<<dim3(nt, nt), 256, 0, c.stream>>>(h.nz, qs.Q.p + h.qoff, qq.QQ.p + qqo);
    c.launches++; post_launch("form_qq_huge");
  }
  const int bignz = qs.bignz;
  static int qq_batch = -1;
  if (qq_batch < 0) { const char *e = getenv("AMGB_QQ_BATCH"); qq_batch = (e && *e == '1') ? 1 : 0; }
  if (qs.maxnz_small > QQ_BIG) {
    // the few columns with a very large support (the reference piles every F row without a
    // coupling into column 0, :2229) would otherwise be one block each and set the run time
    Buf<int> big(n), nbig(1);
    nbig.zero();
    const int *wro = Wt.ro.p;
    int *bp = big.p, *np_ = nbig.p;
    parallel_for(n, [=] DEV(i64 i) { const int nz = wro[i + 1] - wro[i]; if (nz > QQ_BIG && nz <= bignz) bp[atomic_add(np_, 1)] = (int)i; });
    const int nb = nbig.get(0);
    if (nb) {
      k_form_qq_big<<<dim3(nb, 96), 256, 0, c.stream>>>(big.p, Wt.ro.p, qs.Q.p, qs.qoff.p, qq.QQ.p, qq.qqoff.p);
      c.launches++; post_launch("form_qq_big");
    }
  }
  if (qs.maxnz_small <= 12)
    k_form_qq_small<<<(n + 31) / 32, 256, 0, c.stream>>>(n, Wt.ro.p, qs.Q.p, qs.qoff.p, qq.QQ.p, qq.qqoff.p, bignz);
  else
    k_form_qq<<<n, 128, 0, c.stream>>>(n, Wt.ro.p, qs.Q.p, qs.qoff.p, qq.QQ.p, qq.qqoff.p, bignz, qq_batch);
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

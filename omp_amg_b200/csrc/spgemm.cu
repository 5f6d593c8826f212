// spgemm.cu -- the Galerkin-product SpGEMM (mxm, amg_setup.c:1894) as a two-phase hash SpGEMM.
//
// Reference semantics that must survive: X[i][c] = sum over k ASCENDING of B[k][c]*A[i][k]
// (separate multiply and add), entries whose sum is exactly 0 are not stored, columns ascending.
//
// Kernel shape: G cooperating threads own one row of X (G = 8, 32 or a block).  They walk the
// row of A sequentially; for each k the G threads take the entries of row k of B side by side --
// those have distinct columns, so no two threads ever touch the same accumulator in one step,
// and the steps are ordered by a group barrier: every accumulator sees its addends in ascending
// k.  Accumulators live in an open-addressing table in shared memory sized by the row's upper
// bound (bins), or in HBM for the rare row that does not fit.  Phase 1 counts the surviving
// entries per row, a scan turns counts into row offsets, phase 2 recomputes, sorts the table
// (bitonic, zeros and empty slots pushed to the end) and writes the row.
#include "sparse.cuh"
#include "comm.cuh"

#ifndef AMGB_EMU
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
#endif

namespace amgb {

Csr spgemm_rowhash(const Csr &A, const Csr &B);
Csr spgemm_partitioned(const Csr &A, const Csr &B, Csr (*local)(const Csr &, const Csr &));

#ifndef AMGB_EMU
namespace {
constexpr int EMPTY = 0x7fffffff;

struct BlockGroup {
  __device__ __forceinline__ int thread_rank() const { return threadIdx.x; }
  __device__ __forceinline__ int size() const { return blockDim.x; }
  __device__ __forceinline__ void sync() const { __syncthreads(); }
};

__device__ __forceinline__ unsigned hash_col(int c) { return (unsigned)c * 2654435761u; }

// Fused mode (phase 3): the row is counted AND written in one pass.  Its place is not known yet
// (the row offsets come from a scan over all counts), so the sorted row goes to an arena at an
// offset taken from an atomic bump pointer and is copied to its final place afterwards.  If the
// arena is too small the row only reports its count and the classic second pass redoes the work.
struct Arena {
  int *cols; double *vals; unsigned long long *top; long long cap; long long *roff; int *ovf;
};
// called by ONE thread of the row's group once the survivor count n is known
__device__ __forceinline__ long long arena_claim(const Arena &ar, int i, int n, int *cnt) {
  cnt[i] = n;
  const unsigned long long off = atomicAdd(ar.top, (unsigned long long)n);
  if ((long long)(off + (unsigned long long)n) > ar.cap) { *ar.ovf = 1; ar.roff[i] = -1; return -1; }
  ar.roff[i] = (long long)off;
  return (long long)off;
}

// accumulate row i of A*B into the table (keys, vals) of HS slots (power of two)
template <class Group>
__device__ __forceinline__ void accumulate_row(const Group &g, int i, const int *aro, const int *acol,
                                               const double *aa, const int *bro, const int *bcol,
                                               const double *ba, int *keys, double *vals, int HS) {
  const int r0 = g.thread_rank(), G = g.size();
  const unsigned mask = (unsigned)(HS - 1);
  for (int h = r0; h < HS; h += G) keys[h] = EMPTY;
  g.sync();
  for (int ja = aro[i]; ja < aro[i + 1]; ja++) {
    const int k = acol[ja];
    const double av = aa[ja];
    const int be = bro[k + 1];
    for (int jb = bro[k] + r0; jb < be; jb += G) {
      const int c = bcol[jb];
      const double p = ba[jb] * av;
      unsigned h = hash_col(c) & mask;
      for (;;) {
        const int old = atomicCAS(&keys[h], EMPTY, c);
        if (old == EMPTY) { double v = 0.0; v = v + p; vals[h] = v; break; }
        if (old == c) { vals[h] = vals[h] + p; break; }
        h = (h + 1) & mask;
      }
    }
    g.sync();
  }
}

// Block version with the loads taken off the critical path (blockDim.x <= 256): the entries of the
// row of A are staged 256 at a time in shared memory together with the bounds of their rows of B,
// and every thread holds its first TWO entries of the next row of B in registers while the current
// one is accumulated.  accumulate_row pays the chain acol -> bro -> bcol/ba of dependent loads
// inside every k step; here only the barrier separates consecutive k.
// OPT (optimistic use): the table may be too small for the row.  New keys are counted; once the
// count passes `limit` no further key is inserted and the row is given up at the next barrier
// (returns false; the caller hands the row to a kernel with a larger table).
struct BlockStage { int b0[256], b1[256]; double av[256]; int fill, full; };
template <bool OPT>
__device__ __forceinline__ bool accumulate_row_block(BlockStage &sg, int i, const int *aro, const int *acol,
                                                     const double *aa, const int *bro, const int *bcol,
                                                     const double *ba, int *keys, double *vals, int HS, int limit) {
  const int t = threadIdx.x, T = blockDim.x;
  const unsigned mask = (unsigned)(HS - 1);
  for (int h = t; h < HS; h += T) keys[h] = EMPTY;
  if (OPT && t == 0) { sg.fill = 0; sg.full = 0; }
  const int a0 = aro[i], a1 = aro[i + 1];
  for (int jc = a0; jc < a1; jc += T) {
    __syncthreads();
    const int my = jc + t;
    if (my < a1) { const int mk = acol[my]; sg.av[t] = aa[my]; sg.b0[t] = bro[mk]; sg.b1[t] = bro[mk + 1]; }
    __syncthreads();
    const int ns = min(T, a1 - jc);
    int b0 = sg.b0[0], b1 = sg.b1[0];
    int pc0 = EMPTY, pc1 = EMPTY;
    double pv0 = 0.0, pv1 = 0.0;
    if (b0 + t < b1) { pc0 = bcol[b0 + t]; pv0 = ba[b0 + t]; }
    if (b0 + T + t < b1) { pc1 = bcol[b0 + T + t]; pv1 = ba[b0 + T + t]; }
    for (int st = 0; st < ns; st++) {
      const int cb0 = b0, cb1 = b1, cc0 = pc0, cc1 = pc1;
      const double cv0 = pv0, cv1 = pv1, av = sg.av[st];
      if (st + 1 < ns) {
        b0 = sg.b0[st + 1]; b1 = sg.b1[st + 1];
        if (b0 + t < b1) { pc0 = bcol[b0 + t]; pv0 = ba[b0 + t]; }
        if (b0 + T + t < b1) { pc1 = bcol[b0 + T + t]; pv1 = ba[b0 + T + t]; }
      }
      int it = 0;
      for (int jb = cb0 + t; jb < cb1; jb += T, it++) {
        const int c = it == 0 ? cc0 : it == 1 ? cc1 : bcol[jb];
        const double p = (it == 0 ? cv0 : it == 1 ? cv1 : ba[jb]) * av;
        unsigned h = hash_col(c) & mask;
        for (;;) {
          if (OPT) {                       // look before claiming a slot: a full table takes no new key
            const int cur = keys[h];
            if (cur == EMPTY) {
              if (atomicAdd(&sg.fill, 1) >= limit) { sg.full = 1; break; }
              const int old = atomicCAS(&keys[h], EMPTY, c);
              if (old == EMPTY) { double v = 0.0; v = v + p; vals[h] = v; break; }
              atomicSub(&sg.fill, 1);      // another thread took the slot (a different column): go on
              h = (h + 1) & mask;
              continue;
            }
            if (cur == c) { vals[h] = vals[h] + p; break; }
            h = (h + 1) & mask;
          } else {
            const int old = atomicCAS(&keys[h], EMPTY, c);
            if (old == EMPTY) { double v = 0.0; v = v + p; vals[h] = v; break; }
            if (old == c) { vals[h] = vals[h] + p; break; }
            h = (h + 1) & mask;
          }
        }
      }
      __syncthreads();
      if (OPT && sg.full) return false;    // uniform: read after the barrier
    }
  }
  __syncthreads();
  return true;
}

// The same accumulation for a tile of G <= 32 threads with the loads taken off the critical path:
// G entries of the row of A (a_ik and the bounds of row k of B) are read at once, one per thread,
// and the first G entries of the NEXT row of B are already in registers while the current one is
// accumulated.  Order of the additions: unchanged (k ascending, one k at a time).
template <int G, class Tile>
__device__ __forceinline__ void accumulate_row_tile(const Tile &g, int i, const int *aro, const int *acol,
                                                    const double *aa, const int *bro, const int *bcol,
                                                    const double *ba, int *keys, double *vals, int HS) {
  const int r0 = g.thread_rank();
  const unsigned mask = (unsigned)(HS - 1);
  for (int h = r0; h < HS; h += G) keys[h] = EMPTY;
  g.sync();
  const int a0 = aro[i], a1 = aro[i + 1];
  for (int jc = a0; jc < a1; jc += G) {
    const int my = jc + r0;
    int mb0 = 0, mb1 = 0;
    double mav = 0.0;
    if (my < a1) { const int mk = acol[my]; mav = aa[my]; mb0 = bro[mk]; mb1 = bro[mk + 1]; }
    const int ns = min(G, a1 - jc);
    int b0 = g.shfl(mb0, 0), b1 = g.shfl(mb1, 0);
    int pc = EMPTY;
    double pv = 0.0;
    if (b0 + r0 < b1) { pc = bcol[b0 + r0]; pv = ba[b0 + r0]; }
    for (int st = 0; st < ns; st++) {
      const int cb0 = b0, cb1 = b1, cc = pc;
      const double cv = pv;
      const double av = g.shfl(mav, st);
      if (st + 1 < ns) {
        b0 = g.shfl(mb0, st + 1); b1 = g.shfl(mb1, st + 1);
        pc = EMPTY;
        if (b0 + r0 < b1) { pc = bcol[b0 + r0]; pv = ba[b0 + r0]; }
      }
      for (int jb = cb0 + r0; jb < cb1; jb += G) {
        const bool first = (jb < cb0 + G);
        const int c = first ? cc : bcol[jb];
        const double p = (first ? cv : ba[jb]) * av;
        unsigned h = hash_col(c) & mask;
        for (;;) {
          const int old = atomicCAS(&keys[h], EMPTY, c);
          if (old == EMPTY) { double v = 0.0; v = v + p; vals[h] = v; break; }
          if (old == c) { vals[h] = vals[h] + p; break; }
          h = (h + 1) & mask;
        }
      }
      g.sync();
    }
  }
}

// exact zeros leave the row (mxm stores y[ib] only if != 0); returns the survivor count
template <class Group>
__device__ __forceinline__ int drop_zeros_count(const Group &g, int *keys, const double *vals, int HS, int *red) {
  const int r0 = g.thread_rank(), G = g.size();
  int c = 0;
  for (int h = r0; h < HS; h += G) {
    if (keys[h] != EMPTY) { if (vals[h] == 0.0) keys[h] = EMPTY; else c++; }
  }
  // group sum through a shared counter
  if (r0 == 0) *red = 0;
  g.sync();
  if (c) atomicAdd(red, c);
  g.sync();
  return *red;
}

template <class Group>
__device__ __forceinline__ void bitonic_sort(const Group &g, int *keys, double *vals, int HS) {
  const int r0 = g.thread_rank(), G = g.size();
  for (int k = 2; k <= HS; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int idx = r0; idx < HS; idx += G) {
        const int ixj = idx ^ j;
        if (ixj > idx) {
          const int ka = keys[idx], kb = keys[ixj];
          const bool up = ((idx & k) == 0);
          if ((ka > kb) == up && ka != kb) {
            keys[idx] = kb; keys[ixj] = ka;
            const double t = vals[idx]; vals[idx] = vals[ixj]; vals[ixj] = t;
          }
        }
      }
      g.sync();
    }
  }
}

template <class Group>
__device__ __forceinline__ void row_phase(const Group &g, int phase, int i, const int *aro, const int *acol,
                                          const double *aa, const int *bro, const int *bcol, const double *ba,
                                          int *keys, double *vals, int HS, int *red, int *cnt, const int *xro,
                                          int *xcol, double *xa, const Arena &ar, long long *sbase) {
  accumulate_row(g, i, aro, acol, aa, bro, bcol, ba, keys, vals, HS);
  const int n = drop_zeros_count(g, keys, vals, HS, red);
  if (phase == 1) { if (g.thread_rank() == 0) cnt[i] = n; return; }
  long long base = xro ? xro[i] : 0;
  if (phase == 3) {
    if (g.thread_rank() == 0) *sbase = arena_claim(ar, i, n, cnt);
    g.sync();
    base = *sbase;
    if (base < 0) return;
    xcol = ar.cols; xa = ar.vals;
  }
  bitonic_sort(g, keys, vals, HS);
  for (int q = g.thread_rank(); q < n; q += g.size()) { xcol[base + q] = keys[q]; xa[base + q] = vals[q]; }
}

// G-thread tiles, table in shared memory: HS slots per tile.  The few survivors (at most CAP) are
// compacted into a list and ranked by column (rank = number of smaller columns), which is much
// cheaper than sorting the whole table.
template <int G, int HS, int CAP>
__global__ void __launch_bounds__(256) k_spgemm_tile(int phase, const int *list, int nlist, const int *aro,
                                                     const int *acol, const double *aa, const int *bro,
                                                     const int *bcol, const double *ba, int *cnt, const int *xro,
                                                     int *xcol, double *xa, Arena ar) {
  constexpr int PER = 256 / G;
  __shared__ int skeys[PER * HS];
  __shared__ double svals[PER * HS];
  __shared__ int lkeys[PER * CAP];
  __shared__ double lvals[PER * CAP];
  __shared__ int sred[PER];
  __shared__ long long sbase[PER];
  auto tile = cg::tiled_partition<G>(cg::this_thread_block());
  const int slot = threadIdx.x / G;
  const int idx = blockIdx.x * PER + slot;
  if (idx >= nlist) return;
  const int i = list[idx];
  int *keys = skeys + slot * HS;
  double *vals = svals + slot * HS;
  accumulate_row_tile<G>(tile, i, aro, acol, aa, bro, bcol, ba, keys, vals, HS);
  const int r0 = tile.thread_rank();
  if (phase == 1) {
    const int n = drop_zeros_count(tile, keys, vals, HS, sred + slot);
    if (r0 == 0) cnt[i] = n;
    return;
  }
  int *lk = lkeys + slot * CAP;
  double *lv = lvals + slot * CAP;
  if (r0 == 0) sred[slot] = 0;
  tile.sync();
  for (int h = r0; h < HS; h += G)
    if (keys[h] != EMPTY && vals[h] != 0.0) { const int p = atomicAdd(&sred[slot], 1); lk[p] = keys[h]; lv[p] = vals[h]; }
  tile.sync();
  const int n = sred[slot];
  long long base = xro ? xro[i] : 0;
  if (phase == 3) {
    if (r0 == 0) sbase[slot] = arena_claim(ar, i, n, cnt);
    tile.sync();
    base = sbase[slot];
    if (base < 0) return;
    xcol = ar.cols; xa = ar.vals;
  }
  for (int e = r0; e < n; e += G) {
    const int key = lk[e];
    int rank = 0;
    for (int f = 0; f < n; f++) rank += (lk[f] < key);
    xcol[base + rank] = key; xa[base + rank] = lv[e];
  }
}

// one block per row, table in HBM (rows whose bound exceeds the shared-memory bins)
__global__ void k_spgemm_global(int phase, const int *list, int nlist, const i64 *toff, int *gkeys, double *gvals,
                                const int *aro, const int *acol, const double *aa, const int *bro,
                                const int *bcol, const double *ba, int *cnt, const int *xro, int *xcol,
                                double *xa, Arena ar) {
  __shared__ int sred;
  __shared__ long long sbase;
  if ((int)blockIdx.x >= nlist) return;
  const i64 base = toff[blockIdx.x];
  const int HS = (int)(toff[blockIdx.x + 1] - base);
  BlockGroup g;
  row_phase(g, phase, list[blockIdx.x], aro, acol, aa, bro, bcol, ba, gkeys + base, gvals + base, HS, &sred, cnt,
            xro, xcol, xa, ar, &sbase);
}

// ---- block-wide exclusive scan of one int per thread (blockDim.x <= 256) ----
__device__ __forceinline__ int block_excl_scan(int v, int *tmp, int *total) {
  const int t = threadIdx.x, T = blockDim.x;
  tmp[t] = v;
  __syncthreads();
  for (int off = 1; off < T; off <<= 1) {
    const int add = (t >= off) ? tmp[t - off] : 0;
    __syncthreads();
    tmp[t] += add;
    __syncthreads();
  }
  const int incl = tmp[t];
  if (total) *total = tmp[T - 1];
  __syncthreads();
  return incl - v;
}

// Dense accumulator: when B has few columns the whole row of X fits in shared memory as a dense
// array; no hashing and no sorting, the columns come out in order.  Untouched and exactly
// cancelled entries are both 0.0 and both dropped, which is what mxm does.
__global__ void __launch_bounds__(256) k_spgemm_dense(int phase, const int *cminv, const int *spanv,
                                                      const int *list, int nlist,
                                                      const int *aro, const int *acol, const double *aa,
                                                      const int *bro, const int *bcol, const double *ba, int *cnt,
                                                      const int *xro, int *xcol, double *xa, Arena ar) {
  extern __shared__ double acc[];
  __shared__ int stmp[256];
  __shared__ int stotal;
  __shared__ long long sbase;
  if ((int)blockIdx.x >= nlist) return;
  const int i = list[blockIdx.x];
  const int cmin = cminv[i], cn = spanv[i];     // the row only touches columns [cmin, cmin+cn)
  const int t = threadIdx.x, T = blockDim.x;
  for (int c = t; c < cn; c += T) acc[c] = 0.0;
  // 256 entries of the row of A (a_ik and the bounds of row k of B) are staged in shared memory
  // at a time; every thread keeps its entry of the next row of B in registers while the current
  // one is added, so only the barrier separates consecutive k.
  __shared__ int sb0[256], sb1[256];
  __shared__ double sav[256];
  const int a0 = aro[i], a1 = aro[i + 1];
  for (int jc = a0; jc < a1; jc += 256) {
    __syncthreads();
    const int my = jc + t;
    if (my < a1) { const int mk = acol[my]; sav[t] = aa[my]; sb0[t] = bro[mk]; sb1[t] = bro[mk + 1]; }
    __syncthreads();
    const int ns = min(256, a1 - jc);
    int b0 = sb0[0], b1 = sb1[0];
    int pc = 0;
    double pv = 0.0;
    if (b0 + t < b1) { pc = bcol[b0 + t]; pv = ba[b0 + t]; }
    for (int st = 0; st < ns; st++) {
      const int cb0 = b0, cb1 = b1, cc = pc;
      const double cv = pv, av = sav[st];
      if (st + 1 < ns) {
        b0 = sb0[st + 1]; b1 = sb1[st + 1];
        if (b0 + t < b1) { pc = bcol[b0 + t]; pv = ba[b0 + t]; }
      }
      for (int jb = cb0 + t; jb < cb1; jb += T) {
        const bool first = (jb < cb0 + T);
        const int c = (first ? cc : bcol[jb]) - cmin;
        acc[c] = acc[c] + (first ? cv : ba[jb]) * av;
      }
      __syncthreads();
    }
  }
  const int seg = (cn + T - 1) / T;
  const int c0 = t * seg, c1 = min(cn, c0 + seg);
  int mine = 0;
  for (int c = c0; c < c1; c++) mine += (acc[c] != 0.0);
  const int off = block_excl_scan(mine, stmp, &stotal);
  if (phase == 1) { if (t == 0) cnt[i] = stotal; return; }
  long long rowbase = xro ? xro[i] : 0;
  if (phase == 3) {
    if (t == 0) sbase = arena_claim(ar, i, stotal, cnt);
    __syncthreads();
    rowbase = sbase;
    if (rowbase < 0) return;
    xcol = ar.cols; xa = ar.vals;
  }
  long long p = rowbase + off;
  for (int c = c0; c < c1; c++) if (acc[c] != 0.0) { xcol[p] = c + cmin; xa[p] = acc[c]; p++; }
}

// Warp-per-row kernel for rows of moderate size: a hash table of HS slots, a bitmap over the
// row's column span and its word prefix, all in the warp's slice of shared memory.  Used in two
// ways: with a table that is certainly large enough (bound <= HS/2), and OPTIMISTICALLY for rows
// whose bound (sum of B row lengths) is large but whose number of distinct columns is usually
// small: the warp counts its insertions and gives the row up once the table is 3/4 full; such
// rows are collected in `overflow` and redone by the dense block kernel.
template <int HS, int WORDS, int WPB>
__global__ void __launch_bounds__(WPB * 32) k_spgemm_warp_bitmap(int phase, const int *cminv, const int *spanv,
                                                            const int *list, int nlist, const int *aro,
                                                            const int *acol, const double *aa, const int *bro,
                                                            const int *bcol, const double *ba, int *cnt,
                                                            const int *xro, int *xcol, double *xa,
                                                            int *overflow, int *noverflow, const int *tier,
                                                            int mytier, Arena ar) {
  extern __shared__ double wsm[];
  // per warp: svals[HS] | skeys[HS] | bits[WORDS] | wpre[WORDS] (ushort)
  constexpr int PER_BYTES = HS * 12 + WORDS * 4 + WORDS * 2;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int idx = blockIdx.x * WPB + w;
  if (idx >= nlist) return;
  const int i = list[idx];
  if (tier && tier[i] != mytier) return;    // phase 2: only the rows this table size completed
  char *basep = (char *)wsm + (size_t)w * PER_BYTES;
  double *svals = (double *)basep;
  int *skeys = (int *)(basep + HS * 8);
  unsigned *bits = (unsigned *)(basep + HS * 12);
  unsigned short *wpre = (unsigned short *)(basep + HS * 12 + WORDS * 4);
  const unsigned mask = (unsigned)(HS - 1);
  for (int h = lane; h < HS; h += 32) skeys[h] = EMPTY;
  __syncwarp();
  int filled = 0;
  bool gaveup = false;
  // The row of A is taken 32 entries at a time (one per lane: k, a_ik and the bounds of row k of
  // B), and the first 32 entries of the NEXT row of B are already in registers while the current
  // one is accumulated, so the chain acol -> bro -> bcol/ba of dependent loads is off the critical
  // path.  The accumulation order (k ascending, one k at a time) is unchanged.
  const int a0 = aro[i], a1 = aro[i + 1];
  for (int jc = a0; jc < a1 && !gaveup; jc += 32) {
    const int my = jc + lane;
    int mb0 = 0, mb1 = 0;
    double mav = 0.0;
    if (my < a1) { const int mk = acol[my]; mav = aa[my]; mb0 = bro[mk]; mb1 = bro[mk + 1]; }
    const int ns = min(32, a1 - jc);
    int b0 = __shfl_sync(0xffffffffu, mb0, 0), b1 = __shfl_sync(0xffffffffu, mb1, 0);
    int pc = EMPTY;
    double pv = 0.0;
    if (b0 + lane < b1) { pc = bcol[b0 + lane]; pv = ba[b0 + lane]; }
    for (int st = 0; st < ns; st++) {
      const int cb0 = b0, cb1 = b1, cc = pc;
      const double cv = pv;
      const double av = __shfl_sync(0xffffffffu, mav, st);
      if (st + 1 < ns) {
        b0 = __shfl_sync(0xffffffffu, mb0, st + 1); b1 = __shfl_sync(0xffffffffu, mb1, st + 1);
        pc = EMPTY;
        if (b0 + lane < b1) { pc = bcol[b0 + lane]; pv = ba[b0 + lane]; }
      }
      if (overflow && filled + (cb1 - cb0) > HS - 64) { gaveup = true; break; }   // never let the table fill up
      int mine = 0;
      for (int jb = cb0 + lane; jb < cb1; jb += 32) {
        const bool first = (jb < cb0 + 32);
        const int c = first ? cc : bcol[jb];
        const double p = (first ? cv : ba[jb]) * av;
        unsigned h = hash_col(c) & mask;
        for (;;) {
          const int old = atomicCAS(&skeys[h], EMPTY, c);
          if (old == EMPTY) { double v = 0.0; v = v + p; svals[h] = v; mine++; break; }
          if (old == c) { svals[h] = svals[h] + p; break; }
          h = (h + 1) & mask;
        }
      }
      filled += __reduce_add_sync(0xffffffffu, mine);
      __syncwarp();
      if (overflow && filled > HS * 3 / 4) { gaveup = true; break; }
    }
  }
  if (gaveup) {
    if (lane == 0) { overflow[atomicAdd(noverflow, 1)] = i; }
    return;
  }
  int n = 0;
  for (int h = lane; h < HS; h += 32)
    if (skeys[h] != EMPTY) { if (svals[h] == 0.0) skeys[h] = EMPTY; else n++; }
  n = __reduce_add_sync(0xffffffffu, n);
  if (phase == 1) { if (lane == 0) cnt[i] = n; return; }
  long long base = xro ? xro[i] : 0;
  if (phase == 3) {
    long long bb = 0;
    if (lane == 0) bb = arena_claim(ar, i, n, cnt);
    base = __shfl_sync(0xffffffffu, bb, 0);
    if (base < 0) return;
    xcol = ar.cols; xa = ar.vals;
  }
  const int cmin = cminv[i];
  const int nw = (spanv[i] + 31) / 32;
  for (int q = lane; q < nw; q += 32) bits[q] = 0u;
  __syncwarp();
  for (int h = lane; h < HS; h += 32) { const int c = skeys[h]; if (c != EMPTY) { const int d = c - cmin; atomicOr(&bits[d >> 5], 1u << (d & 31)); } }
  __syncwarp();
  const int seg = (nw + 31) / 32;
  const int w0 = lane * seg, w1 = min(nw, w0 + seg);
  int mine = 0;
  for (int q = w0; q < w1; q++) mine += __popc(bits[q]);
  int incl = mine;
  for (int off = 1; off < 32; off <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += t; }
  int run = incl - mine;
  for (int q = w0; q < w1; q++) { wpre[q] = (unsigned short)run; run += __popc(bits[q]); }
  __syncwarp();
  for (int h = lane; h < HS; h += 32) {
    const int c = skeys[h];
    if (c == EMPTY) continue;
    const int d = c - cmin;
    const int rank = wpre[d >> 5] + __popc(bits[d >> 5] & ((1u << (d & 31)) - 1u));
    xcol[base + rank] = c; xa[base + rank] = svals[h];
  }
}

// Hash table (shared memory, or HBM for rows that do not fit) plus a bitmap over the row's column
// span in shared memory: the rank of a column is the number of set bits below it, so the row is
// written in column order without sorting.
// sel != null: the block works on list position sel[blockIdx.x] (rows given up by the optimistic
// launch); ovf != null: optimistic launch, rows whose table fills up are appended to ovf.
__global__ void __launch_bounds__(256) k_spgemm_bitmap(int phase, int HS_smem, int maxwords, const int *cminv,
                                                       const int *spanv, const int *list, int nlist,
                                                       const i64 *toff, int *gkeys, double *gvals,
                                                       const int *aro, const int *acol, const double *aa,
                                                       const int *bro, const int *bcol, const double *ba,
                                                       int *cnt, const int *xro, int *xcol, double *xa, Arena ar,
                                                       const int *sel, int *ovf, int *novf, int optlimit) {
  extern __shared__ double dsm[];
  __shared__ int sred;
  __shared__ int stmp[256];
  __shared__ long long sbase;
  if ((int)blockIdx.x >= nlist) return;
  const int q = sel ? sel[blockIdx.x] : (int)blockIdx.x;
  const int i = list[q];
  int HS;
  double *svals;
  int *skeys;
  unsigned *bits;
  if (HS_smem > 0) {
    HS = HS_smem; svals = dsm; skeys = (int *)(dsm + HS); bits = (unsigned *)(skeys + HS);
  } else {
    const i64 base = toff[blockIdx.x];       // tables are laid out in launch order
    HS = (int)(toff[blockIdx.x + 1] - base); svals = gvals + base; skeys = gkeys + base; bits = (unsigned *)dsm;
  }
  int *wpre = (int *)(bits + maxwords);
  BlockGroup g;
  __shared__ BlockStage stage;
  if (ovf) {
    if (!accumulate_row_block<true>(stage, i, aro, acol, aa, bro, bcol, ba, skeys, svals, HS, optlimit)) {
      if (threadIdx.x == 0) ovf[atomicAdd(novf, 1)] = q;
      return;
    }
  } else {
    accumulate_row_block<false>(stage, i, aro, acol, aa, bro, bcol, ba, skeys, svals, HS, 0);
  }
  const int n = drop_zeros_count(g, skeys, svals, HS, &sred);
  if (phase == 1) { if (threadIdx.x == 0) cnt[i] = n; return; }
  long long base = xro ? xro[i] : 0;
  if (phase == 3) {
    if (threadIdx.x == 0) sbase = arena_claim(ar, i, n, cnt);
    __syncthreads();
    base = sbase;
    if (base < 0) return;
    xcol = ar.cols; xa = ar.vals;
  }
  const int cmin = cminv[i];
  const int nw = (spanv[i] + 31) / 32;
  const int t = threadIdx.x, T = blockDim.x;
  for (int w = t; w < nw; w += T) bits[w] = 0u;
  __syncthreads();
  for (int h = t; h < HS; h += T) { const int c = skeys[h]; if (c != EMPTY) { const int d = c - cmin; atomicOr(&bits[d >> 5], 1u << (d & 31)); } }
  __syncthreads();
  const int seg = (nw + T - 1) / T;
  const int w0 = t * seg, w1 = min(nw, w0 + seg);
  int mine = 0;
  for (int w = w0; w < w1; w++) mine += __popc(bits[w]);
  int run = block_excl_scan(mine, stmp, nullptr);
  for (int w = w0; w < w1; w++) { wpre[w] = run; run += __popc(bits[w]); }
  __syncthreads();
  for (int h = t; h < HS; h += T) {
    const int c = skeys[h];
    if (c == EMPTY) continue;
    const int d = c - cmin;
    const int rank = wpre[d >> 5] + __popc(bits[d >> 5] & ((1u << (d & 31)) - 1u));
    xcol[base + rank] = c; xa[base + rank] = svals[h];
  }
}
template <int G>
__global__ void __launch_bounds__(256) k_arena_copy(int rn, const int *xro, const long long *roff, const int *acols,
                                                    const double *avals, int *xcol, double *xa) {
  const int i = blockIdx.x * (256 / G) + threadIdx.x / G;
  if (i >= rn) return;
  const int lane = threadIdx.x % G;
  const int b = xro[i], n = xro[i + 1] - b;
  const long long o = roff[i];
  for (int q = lane; q < n; q += G) { xcol[b + q] = acols[o + q]; xa[b + q] = avals[o + q]; }
}
}  // namespace

struct SpgemmStats { std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev; i64 bytes = 0; i64 calls = 0; };
static SpgemmStats g_stats;
void spgemm_stats_reset() {
  for (auto &e : g_stats.ev) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
  g_stats.ev.clear(); g_stats.bytes = 0; g_stats.calls = 0;
}
// device seconds spent between the recorded event pairs; call after a stream sync
void spgemm_stats_get(double *seconds, i64 *bytes, i64 *calls) {
  double s = 0;
  for (auto &e : g_stats.ev) { float ms = 0; if (cudaEventElapsedTime(&ms, e.first, e.second) == cudaSuccess) s += ms * 1e-3; }
  *seconds = s; *bytes = g_stats.bytes; *calls = g_stats.calls;
}

static int g_spgemm_impl = -1;

static Csr spgemm_core_local(const Csr &A, const Csr &B);
// one GPU: the kernels below; several ranks: the rows of X are partitioned (sparse.cu)
static Csr spgemm_core(const Csr &A, const Csr &B) {
  if (comm_active() && A.rn >= comm_size() && A.nnz + B.nnz >= comm_min_work())
    return spgemm_partitioned(A, B, spgemm_core_local);
  return spgemm_core_local(A, B);
}

// X = A*B through the transposed product when that has the better shape.  Entry by entry,
// (B'A')[c][i] = sum over k ascending of A'[k][i]*B'[c][k] has the same addends in the same
// order as (AB)[i][c] = sum over k ascending of B[k][c]*A[i][k] (a product of two doubles does
// not depend on the order of its factors), so X = (B'A')' bit for bit, exact zeros included.
// Row-wise SpGEMM does one ordered step per entry of the row of A; long rows of A against short
// rows of B (Af*W on the coarse levels: 500 against 50) are hundreds of nearly empty steps, while
// the transposed product takes few steps that each fill the whole thread block.
static Csr g_At_cache;
static unsigned long long g_At_key = 0;     // Csr::uid of the cached operand (0 = empty)
void spgemm_cache_reset() { g_At_cache = Csr(); g_At_key = 0; }

Csr spgemm(const Csr &A, const Csr &B) {
  if (g_spgemm_impl < 0) { const char *e = getenv("AMGB_SPGEMM"); g_spgemm_impl = (e && !strcmp(e, "rowhash")) ? 0 : 1; }
  if (g_spgemm_impl == 0) return spgemm_rowhash(A, B);
  if (A.cn != B.rn) throw Error(-4, "spgemm: dimension mismatch");
  static int tr_on = -1;
  if (tr_on < 0) { const char *e = getenv("AMGB_SPGEMM_TRANSPOSED"); tr_on = (e && *e == '0') ? 0 : 1; }
  const bool force_t = test_force('t');
  if ((tr_on || force_t) && A.rn > 0 && B.rn > 0 && (A.nnz > (1 << 20) || (force_t && A.nnz > 0)) && B.nnz > 0) {
    const double la = (double)A.nnz / A.rn, lb = (double)B.nnz / B.rn;
    if (force_t || (la > 64.0 && la > 4.0 * lb)) {
      StageTimer st_("prim.spgemm(transposed)");
      if (A.uid == 0 || g_At_key != A.uid) {   // A = Af recurs within a level
        g_At_cache = transpose(A);
        g_At_key = A.uid;
      }
      Csr Bt = transpose(B);
      Csr Xt = spgemm_core(Bt, g_At_cache);
      return transpose(Xt);
    }
  }
  return spgemm_core(A, B);
}

static Csr spgemm_core_local(const Csr &A, const Csr &B) {
  StageTimer st_("prim.spgemm");
  Context &c = ctx();
  const int rn = A.rn;
  if (rn == 0) { Csr X(0, B.cn, 0); X.ro.zero(); return X; }
  const int *aro = A.ro.p, *acol = A.col.p, *bro = B.ro.p, *bcol = B.col.p;
  const double *aa = A.a.p, *ba = B.a.p;
  // Per row: need = min(sum of B row lengths, columns of B) bounds the distinct columns;
  // [cmin, cmin+span) is the column range the row can touch (B rows are sorted, so it comes from
  // their first and last entries).  Bins:
  //   0  need <= 24                  8-thread tiles, 64-slot hash tables in shared memory, bitonic
  //   1  need <= 96                  warps, 256-slot tables, bitonic
  //   2  need <= 256, span <= 32512  warps, 512-slot tables + bitmap over the span
  //   3  span <= 24576               warps, 1024-slot tables + bitmap, optimistic; rows whose distinct
  //                                  columns do not fit are redone by a block with a dense accumulator
  //   6  need <= 768                 block, 2048-slot table + bitmap over the span
  //   7  need <= 3072                block, 8192-slot table + bitmap over the span
  //   8  span <= 800k                block, table in HBM + bitmap over the span in shared memory
  //   9  anything else               block, table in HBM, bitonic sort
  constexpr int NB = 10;
  constexpr int BM6_SPAN = 700000, BM7_SPAN = 400000, BM8_SPAN = 800000;   // bitmap = span/4 bytes
  const int bcn = B.cn;
  Buf<int> lists((i64)NB * rn), bcnt(NB), need(rn), cminv(rn), spanv(rn), maxspan(NB);
  Buf<unsigned long long> needsum(1);      // sum of the per-row bounds: an upper bound of nnz(X)
  bcnt.zero(); maxspan.zero(); needsum.zero();
  unsigned long long *nsum = needsum.p;
  int *lp = lists.p, *bc = bcnt.p, *nd = need.p, *cmv = cminv.p, *spv = spanv.p, *mxs = maxspan.p;
  const bool small = test_small_bins();
  parallel_for(rn, [=] DEV(i64 i) {
    i64 ub = 0;
    int lo = 0x7fffffff, hi = -1;
    for (int ja = aro[i]; ja < aro[i + 1]; ja++) {
      const int k = acol[ja], b0 = bro[k], b1 = bro[k + 1];
      ub += b1 - b0;
      if (b1 > b0) { const int f = bcol[b0], l = bcol[b1 - 1]; if (f < lo) lo = f; if (l > hi) hi = l; }
    }
    if (ub > bcn) ub = bcn;
    const int span = hi >= lo ? hi - lo + 1 : 0;
    if (ub > span) ub = span;
    nd[i] = (int)ub; cmv[i] = hi >= lo ? lo : 0; spv[i] = span;
    int bin;
    if (small && ub > 24 && span <= BM8_SPAN) bin = 8;      // test hook: block kernel + overflow path
    else if (ub <= 24) bin = 0;
    else if (ub <= 96) bin = 1;
    else if (ub <= 256 && span <= 32512) bin = 2;
    else if (span <= 24576) bin = 3;          // optimistic warp kernel first, dense fallback (bins 4,5 unused)
    else if (ub <= 768 && span <= BM6_SPAN) bin = 6;
    else if (ub <= 3072 && span <= BM7_SPAN) bin = 7;
    else if (span <= BM8_SPAN) bin = 8;
    else bin = 9;
    const int p = atomic_add(&bc[bin], 1);
    lp[(i64)bin * rn + p] = (int)i;
    atomic_max_i32(&mxs[bin], span);
    if (ub) atomic_add(nsum, (unsigned long long)ub);
  });
  std::vector<int> hc = bcnt.download();
  const i64 need_total = (i64)needsum.get(0);
  std::vector<int> hms = maxspan.download();
  // rows of bins 8 and 9 get tables in HBM
  Buf<i64> tsz5, toff5, tsz6, toff6;
  Buf<int> gkeys5, gkeys6;
  Buf<double> gvals5, gvals6;
  Buf<int> ovf8, novf8;
  int n_ovf8 = 0;
  bool sel8 = false;       // the HBM launch of bin 8 works on the positions listed in ovf8
  for (int bin = 9; bin <= 9; bin++) {     // bin 8 gets its tables after the optimistic launch
    if (!hc[bin]) continue;
    Buf<i64> &tsz = bin == 8 ? tsz5 : tsz6, &toff = bin == 8 ? toff5 : toff6;
    tsz.alloc(hc[bin] + 1); toff.alloc(hc[bin] + 1);
    i64 *ts = tsz.p;
    const int *lb = lp + bin * (i64)rn;
    parallel_for(hc[bin], [=] DEV(i64 q) { i64 s = 256; while (s < 2 * (i64)nd[lb[q]]) s <<= 1; ts[q] = s; });
    const i64 total = exclusive_scan64(tsz.p, toff.p, hc[bin]);
    (bin == 8 ? gkeys5 : gkeys6).alloc(total);
    (bin == 8 ? gvals5 : gvals6).alloc(total);
  }
  Buf<int> cnt(rn + 1), xro(rn + 1);
  cudaEvent_t e0, e1;
  CUDA_CHECK(cudaEventCreate(&e0)); CUDA_CHECK(cudaEventCreate(&e1));
  static bool attr = false;
  if (!attr) {
    CUDA_CHECK(cudaFuncSetAttribute((const void *)k_spgemm_dense, cudaFuncAttributeMaxDynamicSharedMemorySize, 24576 * 8));
    CUDA_CHECK(cudaFuncSetAttribute((const void *)k_spgemm_bitmap, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
    CUDA_CHECK(cudaFuncSetAttribute((const void *)k_spgemm_warp_bitmap<512, 1016, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * (512 * 12 + 1016 * 6)));
    CUDA_CHECK(cudaFuncSetAttribute((const void *)k_spgemm_warp_bitmap<1024, 768, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * (1024 * 12 + 768 * 6)));
    CUDA_CHECK(cudaFuncSetAttribute((const void *)k_spgemm_warp_bitmap<2048, 768, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * (2048 * 12 + 768 * 6)));
    CUDA_CHECK(cudaFuncSetAttribute((const void *)k_spgemm_warp_bitmap<4096, 768, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 12 + 768 * 6));
    attr = true;
  }
  auto words = [](int span) { return (span + 31) / 32; };
  auto L = [&](int bin) { return lp + bin * (i64)rn; };
  Buf<int> ovf[3], novf, tierv;
  int n_ovf[3] = {0, 0, 0};
  Csr X;
  CUDA_CHECK(cudaEventRecord(e0, c.stream));
  // Fused mode: one pass counts and writes every row into an arena, a copy puts the rows in
  // place.  Only if the arena was too small does the classic second pass recompute the rows.
  static int fused_on = -1;
  if (fused_on < 0) { const char *e = getenv("AMGB_SPGEMM_FUSED"); fused_on = (e && *e == '0') ? 0 : 1; }
  Arena ar{nullptr, nullptr, nullptr, 0, nullptr, nullptr};
  Buf<int> acols, aflag;
  Buf<double> avals;
  Buf<unsigned long long> atop;
  Buf<i64> aroff;
  if (fused_on) {
    // the arena holds every row of X: sum of the per-row bounds when that is affordable (the
    // pattern products W_skel*W_skel' expand 15x), else a multiple of the operands -- a too small
    // arena costs a second pass over all rows
    i64 cap = 6 * (A.nnz + B.nnz) + rn;
    if (cap < (1 << 20)) cap = 1 << 20;
    const i64 cap_hi = (i64)400 << 20;             // 400 Mi entries = 4.8 GB
    if (need_total > cap) cap = need_total < cap_hi ? need_total : (cap > cap_hi ? cap : cap_hi);
    if (test_force('o')) cap = 48;                 // test hook: (nearly) every row overflows the arena
    acols.alloc(cap); avals.alloc(cap); atop.alloc(1); aflag.alloc(1); aroff.alloc(rn);
    atop.zero(); aflag.zero();
    ar = Arena{acols.p, avals.p, atop.p, cap, aroff.p, aflag.p};
  }
  bool done = false;
  for (int pass = 0; pass < 2 && !done; pass++) {
    const int phase = fused_on ? (pass == 0 ? 3 : 2) : pass + 1;
    const bool first = (pass == 0);
    int *xcol = phase == 2 ? X.col.p : nullptr;
    double *xa = phase == 2 ? X.a.p : nullptr;
    if (hc[0]) {
      k_spgemm_tile<8, 64, 24><<<(hc[0] + 31) / 32, 256, 0, c.stream>>>(phase, L(0), hc[0], aro, acol, aa, bro, bcol, ba, cnt.p, xro.p, xcol, xa, ar);
      c.launches++; post_launch("spgemm_tile8");
    }
    if (hc[1]) {
      k_spgemm_tile<32, 256, 96><<<(hc[1] + 7) / 8, 256, 0, c.stream>>>(phase, L(1), hc[1], aro, acol, aa, bro, bcol, ba, cnt.p, xro.p, xcol, xa, ar);
      c.launches++; post_launch("spgemm_tile32");
    }
    if (hc[2]) {
      k_spgemm_warp_bitmap<512, 1016, 4><<<(hc[2] + 3) / 4, 128, 4 * (512 * 12 + 1016 * 6), c.stream>>>(
          phase, cmv, spv, L(2), hc[2], aro, acol, aa, bro, bcol, ba, cnt.p, xro.p, xcol, xa, nullptr, nullptr, nullptr, 0, ar);
      c.launches++; post_launch("spgemm_warp_bitmap");
    }
    if (hc[3]) {
      // optimistic ladder: 1024-, 2048-, 4096-slot tables, then the dense block kernel
      if (first) {
        for (int t = 0; t < 3; t++) { ovf[t].alloc(hc[3]); }
        novf.alloc(3); tierv.alloc(rn);
        novf.zero(); tierv.zero();
      }
      static int ntiers = -1;
      if (ntiers < 0) { const char *e = getenv("AMGB_SPGEMM_TIERS"); ntiers = e ? atoi(e) : 1; if (ntiers < 0 || ntiers > 3) ntiers = 1; }
      const int *lists3[4] = {L(3), ovf[0].p, ovf[1].p, ovf[2].p};
      int counts3[4] = {hc[3], n_ovf[0], n_ovf[1], n_ovf[2]};
      if (first && ntiers < 3) {
        // rows skip the tiers that are switched off: everything left goes to the dense kernel
        const int last = ntiers;    // first disabled tier
        if (last == 0) { n_ovf[0] = n_ovf[1] = n_ovf[2] = hc[3]; d2d(ovf[2].p, L(3), sizeof(int) * (size_t)hc[3]);
                         int *fl = tierv.p; const int *ol = L(3); parallel_for(hc[3], [=] DEV(i64 q) { fl[ol[q]] = 3; }); }
      }
      for (int t = 0; t < 3; t++) {
        if (t >= ntiers) {
          if (first && t > 0 && t == ntiers && n_ovf[t - 1]) {   // hand the overflow of the last enabled tier to dense
            n_ovf[2] = n_ovf[t - 1];
            if (t - 1 != 2) d2d(ovf[2].p, ovf[t - 1].p, sizeof(int) * (size_t)n_ovf[t - 1]);
            int *fl = tierv.p; const int *ol = ovf[2].p; parallel_for(n_ovf[2], [=] DEV(i64 q) { fl[ol[q]] = 3; });
          }
          continue;
        }
        if (!counts3[t]) continue;
        int *ol = first ? ovf[t].p : nullptr, *on = first ? novf.p + t : nullptr;
        const int *tv = first ? nullptr : tierv.p;
        if (t == 0)
          k_spgemm_warp_bitmap<1024, 768, 4><<<(counts3[t] + 3) / 4, 128, 4 * (1024 * 12 + 768 * 6), c.stream>>>(
              phase, cmv, spv, lists3[t], counts3[t], aro, acol, aa, bro, bcol, ba, cnt.p, xro.p, xcol, xa, ol, on, tv, t, ar);
        else if (t == 1)
          k_spgemm_warp_bitmap<2048, 768, 2><<<(counts3[t] + 1) / 2, 64, 2 * (2048 * 12 + 768 * 6), c.stream>>>(
              phase, cmv, spv, lists3[t], counts3[t], aro, acol, aa, bro, bcol, ba, cnt.p, xro.p, xcol, xa, ol, on, tv, t, ar);
        else
          k_spgemm_warp_bitmap<4096, 768, 1><<<counts3[t], 32, 4096 * 12 + 768 * 6, c.stream>>>(
              phase, cmv, spv, lists3[t], counts3[t], aro, acol, aa, bro, bcol, ba, cnt.p, xro.p, xcol, xa, ol, on, tv, t, ar);
        c.launches++; post_launch("spgemm_warp_optimistic");
        if (first) {
          std::vector<int> hn = novf.download();
          n_ovf[t] = hn[(size_t)t];
          counts3[t + 1] = n_ovf[t];
          if (n_ovf[t]) { int *fl = tierv.p; const int *olist = ovf[t].p; const int tt = t + 1; parallel_for(n_ovf[t], [=] DEV(i64 q) { fl[olist[q]] = tt; }); }
        }
      }
      if (n_ovf[2]) {
        k_spgemm_dense<<<n_ovf[2], 256, (size_t)(hms[3] > 0 ? hms[3] : 1) * 8, c.stream>>>(phase, cmv, spv, ovf[2].p, n_ovf[2], aro, acol, aa, bro, bcol, ba, cnt.p, xro.p, xcol, xa, ar);
        c.launches++; post_launch("spgemm_dense");
      }
    }
    if (hc[6]) {
      const int mw = words(hms[6]);
      k_spgemm_bitmap<<<hc[6], 128, 2048 * 12 + (size_t)mw * 8, c.stream>>>(phase, 2048, mw, cmv, spv, L(6), hc[6], nullptr, nullptr, nullptr, aro, acol, aa, bro, bcol, ba, cnt.p, xro.p, xcol, xa, ar, nullptr, nullptr, nullptr, 0);
      c.launches++; post_launch("spgemm_bitmap2k");
    }
    if (hc[7]) {
      const int mw = words(hms[7]);
      k_spgemm_bitmap<<<hc[7], 256, 8192 * 12 + (size_t)mw * 8, c.stream>>>(phase, 8192, mw, cmv, spv, L(7), hc[7], nullptr, nullptr, nullptr, aro, acol, aa, bro, bcol, ba, cnt.p, xro.p, xcol, xa, ar, nullptr, nullptr, nullptr, 0);
      c.launches++; post_launch("spgemm_bitmap8k");
    }
    if (hc[8]) {
      // The bound of these rows (sum of the B row lengths) exceeds every shared-memory table, but
      // the number of DISTINCT columns usually does not (Af*W on the coarse levels compresses 15
      // products into one entry): first an optimistic launch with the 8192-slot table in shared
      // memory; rows that fill it are collected and redone with a table in HBM sized by the bound.
      const int mw = words(hms[8]);
      const size_t opt_sm = 8192 * 12 + (size_t)mw * 8;
      const char *o8 = getenv("AMGB_SPGEMM_OPT8");      // =0: straight to the HBM tables (A/B checks)
      const bool optimistic = opt_sm <= 200 * 1024 && !(o8 && *o8 == '0');
      if (first) {
        n_ovf8 = hc[8];
        if (optimistic) {
          ovf8.alloc(hc[8]); novf8.alloc(1); novf8.zero();
          k_spgemm_bitmap<<<hc[8], 256, opt_sm, c.stream>>>(phase, 8192, mw, cmv, spv, L(8), hc[8], nullptr, nullptr, nullptr, aro, acol, aa, bro, bcol, ba, cnt.p, xro.p, xcol, xa, ar, nullptr, ovf8.p, novf8.p, small ? 40 : 8192 / 4 * 3);
          c.launches++; post_launch("spgemm_bitmap_optimistic");
          n_ovf8 = novf8.get(0);
          sel8 = true;
        }
        if (n_ovf8) {                      // HBM tables for the rows that are left, in launch order
          tsz5.alloc(n_ovf8 + 1); toff5.alloc(n_ovf8 + 1);
          i64 *ts = tsz5.p;
          const int *lb = L(8), *sl = optimistic ? ovf8.p : nullptr;
          parallel_for(n_ovf8, [=] DEV(i64 q) { i64 sz = 256; const int row = lb[sl ? sl[q] : (int)q]; while (sz < 2 * (i64)nd[row]) sz <<= 1; ts[q] = sz; });
          const i64 total = exclusive_scan64(tsz5.p, toff5.p, n_ovf8);
          gkeys5.alloc(total); gvals5.alloc(total);
        }
      } else if (optimistic && (n_ovf8 < hc[8] || sel8)) {
        // second pass (the arena was too small, rare): every row through the HBM kernel; which rows
        // the optimistic launch gives up can differ by a few between two runs (the fill count is
        // transiently high while threads race for a slot), so its list is not reused
        n_ovf8 = hc[8]; sel8 = false;
        tsz5.alloc(n_ovf8 + 1); toff5.alloc(n_ovf8 + 1);
        i64 *ts = tsz5.p;
        const int *lb = L(8);
        parallel_for(n_ovf8, [=] DEV(i64 q) { i64 sz = 256; while (sz < 2 * (i64)nd[lb[q]]) sz <<= 1; ts[q] = sz; });
        const i64 total = exclusive_scan64(tsz5.p, toff5.p, n_ovf8);
        gkeys5.alloc(total); gvals5.alloc(total);
      }
      if (n_ovf8) {
        k_spgemm_bitmap<<<n_ovf8, 256, (size_t)mw * 8 + 16, c.stream>>>(phase, 0, mw, cmv, spv, L(8), n_ovf8, toff5.p, gkeys5.p, gvals5.p, aro, acol, aa, bro, bcol, ba, cnt.p, xro.p, xcol, xa, ar, sel8 ? ovf8.p : nullptr, nullptr, nullptr, 0);
        c.launches++; post_launch("spgemm_bitmap_hbm");
      }
    }
    if (hc[9]) {
      k_spgemm_global<<<hc[9], 256, 0, c.stream>>>(phase, L(9), hc[9], toff6.p, gkeys6.p, gvals6.p, aro, acol, aa, bro, bcol, ba, cnt.p, xro.p, xcol, xa, ar);
      c.launches++; post_launch("spgemm_global");
    }
    if (first) {
      CUDA_CHECK(cudaEventRecord(e1, c.stream));
      g_stats.ev.emplace_back(e0, e1);
      const i64 nnz = exclusive_scan(cnt.p, xro.p, rn);
      X = Csr(rn, B.cn, nnz);
      d2d(X.ro.p, xro.p, sizeof(int) * (size_t)(rn + 1));
      CUDA_CHECK(cudaEventCreate(&e0)); CUDA_CHECK(cudaEventCreate(&e1));
      CUDA_CHECK(cudaEventRecord(e0, c.stream));
      if (fused_on && aflag.get(0) == 0) {       // every row is in the arena: copy into place
        if (nnz > 0) {
          if ((double)nnz / (double)rn <= 12.0)
            k_arena_copy<8><<<(rn + 31) / 32, 256, 0, c.stream>>>(rn, xro.p, aroff.p, acols.p, avals.p, X.col.p, X.a.p);
          else
            k_arena_copy<32><<<(rn + 7) / 8, 256, 0, c.stream>>>(rn, xro.p, aroff.p, acols.p, avals.p, X.col.p, X.a.p);
          c.launches++; post_launch("spgemm_arena_copy");
        }
        done = true;
      }
    }
  }
  CUDA_CHECK(cudaEventRecord(e1, c.stream));
  g_stats.ev.emplace_back(e0, e1);
  {
    static int logit = -1;
    if (logit < 0) { const char *e = getenv("AMGB_SPGEMM_LOG"); logit = (e && *e && *e != '0') ? 1 : 0; }
    if (logit) {
      stream_sync();
      float m1 = 0, m2 = 0;
      const size_t ne = g_stats.ev.size();
      cudaEventElapsedTime(&m1, g_stats.ev[ne - 2].first, g_stats.ev[ne - 2].second);
      cudaEventElapsedTime(&m2, g_stats.ev[ne - 1].first, g_stats.ev[ne - 1].second);
      fprintf(stderr, "spgemm A %dx%d nnz %lld  B %dx%d nnz %lld  X nnz %lld | bins %d %d %d %d %d %d %d %d %d %d | phase1 %.3f ms phase2 %.3f ms | %.1f GB/s\n",
              A.rn, A.cn, (long long)A.nnz, B.rn, B.cn, (long long)B.nnz, (long long)X.nnz, hc[0], hc[1], hc[2], hc[3], hc[4],
              n_ovf[0] * 10000 + n_ovf[2], hc[6], hc[7], hc[8], hc[9], m1, m2,
              (12.0 * (A.nnz + B.nnz + X.nnz)) / ((m1 + m2) * 1e-3) / 1e9);
    }
  }
  // algorithmic bytes: A, B and X each moved once (12 B per entry, 4 B per row offset)
  g_stats.bytes += 12 * (A.nnz + B.nnz + X.nnz) + 4 * ((i64)A.rn + B.rn + X.rn + 3);
  g_stats.calls++;
  return X;
}
#else
void spgemm_stats_reset() {}
void spgemm_cache_reset() {}
void spgemm_stats_get(double *seconds, i64 *bytes, i64 *calls) { *seconds = 0; *bytes = 0; *calls = 0; }
Csr spgemm(const Csr &A, const Csr &B) {
  if (comm_active() && A.rn >= comm_size() && A.nnz + B.nnz >= comm_min_work())
    return spgemm_partitioned(A, B, spgemm_rowhash);
  return spgemm_rowhash(A, B);
}
#endif

}  // namespace amgb

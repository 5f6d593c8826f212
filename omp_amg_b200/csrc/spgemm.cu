// spgemm.cu -- the Galerkin-product SpGEMM (mxm, amg_setup.c:1894) as a two-phase hash SpGEMM.
//
// Reference semantics that must survive: X[i][c] = sum over k ASCENDING of B[k][c]*A[i][k]
// (separate multiply and add), entries whose sum is exactly 0 are not stored, columns ascending.
//
// Kernel shape: G cooperating threads own one row of X (G = 8, 32 or a block).  They walk the
// row of A sequentially; for each k the G threads take the entries of row k of B side by side --
// those have distinct columns, so no two threads ever touch the same accumulator in one step,
// and the steps are ordered by a group barrier: every accumulator sees its addends in ascending
// k.  Accumulators live in an open-addressing table in shared memory; rows are placed on a ladder
// of tiers (8-thread tiles, warps with 256/512/2048 slots, blocks with 4096/8192 slots, tables in
// HBM) by a lower and an upper bound of their distinct columns and run OPTIMISTICALLY: a row whose
// table fills up is handed to the next tier on the device (no host round trip).  The rows of B are
// staged asynchronously (cp.async rings per warp, bulk async copies on mbarriers per block) so
// that a k-step never waits for HBM.  One fused pass counts and writes every row, sorted, into an
// arena; a scan gives the row offsets and a copy puts the rows in place.
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

// ---------------------------------------------------------------------------------------
// Asynchronous staging of the rows of B (sm_100a): per-lane cp.async (LDGSTS) rings for the
// warp kernel, bulk async copies (cp.async.bulk -> UBLKCP, completion on an mbarrier) for the
// block kernel.  The k-steps of a row are serialised by the ordering rule, so what a step may
// never do is wait for HBM: the entries of the next rows of B are already in shared memory.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async4(void *dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void *dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// One addend for column c.  Most accesses find their key (a row of W'(Af W) adds 10-40 products
// per entry), so the probe is a plain load; only an empty slot is claimed with a CAS (two threads
// of a step never hold the same column, but they may want the same empty slot).
__device__ __forceinline__ bool table_add(int *keys, double *vals, unsigned mask, int sh, int c, double p) {
  unsigned h = hash_col(c) >> sh;              // the high bits of the multiplicative hash
  for (;;) {
    int cur = ((volatile int *)keys)[h];
    if (cur == EMPTY) {
      cur = atomicCAS(&keys[h], EMPTY, c);
      if (cur == EMPTY) { double v = 0.0; v = v + p; vals[h] = v; return true; }
    }
    if (cur == c) { vals[h] = vals[h] + p; return false; }
    h = (h + 1) & mask;
  }
}

// Which kernel finished a row is recorded (tierdone) so that a second pass -- only needed when the
// arena was too small -- sends every row straight to the kernel that can hold it.
// Rows a kernel cannot hold are appended to the list of the next tier: route[] gives, per row, the
// destination after the warp tiers (block 4096 / block 8192 / HBM table / global), decided by the
// binning kernel from the row's column span.
struct Tiers {
  int *list[11];       // per tier: rows to process (bin members first, rows handed down appended)
  int *count;          // count[t]: entries of list[t] (device side)
  int *cursor;         // cursor[t]: next position a warp/block takes
  signed char *done;   // tier that completed the row
  const unsigned char *route;   // first block tier able to hold the row's span (4..7)
};
__device__ __forceinline__ void hand_down(const Tiers &tr, int from, int i) {
  int to = from + 1;
  if (to >= 4 && tr.route[i] > to) to = tr.route[i];
  tr.list[to][atomicAdd(&tr.count[to], 1)] = i;
}

// ---------------------------------------------------------------------------------------
// Warp per row.  Table of HS slots in the warp's slice of shared memory, filled OPTIMISTICALLY:
// the bound of a row (sum of the B row lengths) says little about its number of distinct columns
// (W'(Af W) compresses 10-40 products into one entry), so a row runs with a small table and is
// handed down if that fills up.  The count of new keys is exact (warp reduction per step), so
// which rows are handed down is deterministic.  The next D rows of B travel through a per-lane
// cp.async ring (each lane copies and later reads its own entries: no barrier besides
// wait_group).  The survivors are compacted, sorted by column as packed (column, slot) words with
// a bitonic network over the next power of two, and written in order.
// ---------------------------------------------------------------------------------------
constexpr int RING_D = 4, RING_C = 64;
// shared memory is addressed by 32-bit offsets into the shared window (ld/st/atom.shared): the
// pointers handed around as generic addresses made the compiler rebuild the window base in every
// k-step (profiles/r2_ncu_spgemm_warp_source_before.txt: 110 instructions of overhead per step)
__device__ __forceinline__ int lds_i32(unsigned a) { int v; asm volatile("ld.shared.s32 %0, [%1];\n" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ double lds_f64(unsigned a) { double v; asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ unsigned lds_u16(unsigned a) { unsigned short v; asm volatile("ld.shared.u16 %0, [%1];\n" : "=h"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts_i32(unsigned a, int v) { asm volatile("st.shared.s32 [%0], %1;\n" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_f64(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;\n" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void sts_u16(unsigned a, unsigned v) { asm volatile("st.shared.u16 [%0], %1;\n" ::"r"(a), "h"((unsigned short)v) : "memory"); }
__device__ __forceinline__ int atoms_cas(unsigned a, int cmp, int val) {
  int old;
  asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;\n" : "=r"(old) : "r"(a), "r"(cmp), "r"(val) : "memory");
  return old;
}
__device__ __forceinline__ void cp_async4s(unsigned dst, const void *src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async8s(unsigned dst, const void *src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(src) : "memory"); }

// per warp: values[HS] | keys[HS] | union { ring values[D*C], ring columns[D*C] ; slots to sort [HS] (u16) }
template <int HS>
struct WarpSm {
  static constexpr int U = (RING_D * RING_C * 12 > HS * 2) ? RING_D * RING_C * 12 : HS * 2;
  static constexpr int PER = HS * 12 + U;
  static constexpr int PER_DENSE = HS * 8 + RING_D * RING_C * 12;
  static constexpr int BITS = HS == 128 ? 7 : HS == 256 ? 8 : HS == 512 ? 9 : HS == 1024 ? 10 : HS == 2048 ? 11 : 12;
};
// DENSE: the table is replaced by one accumulator per column of the row's span (at most HS
// columns: the coarse levels, where a row touches most of the few thousand columns there are); no
// probing, no sorting -- the columns come out in order.
template <int HS, int WPB, bool DENSE>
__global__ void __launch_bounds__(WPB * 32) k_spgemm_warp(int phase, int mytier, Tiers tr, const int *aro, const int *acol,
                                                          const double *aa, const int *bro, const int *bcol,
                                                          const double *ba, int *cnt, const int *xro, int *xcol,
                                                          double *xa, int optimistic, const int *cminv,
                                                          const int *spanv, Arena ar) {
  extern __shared__ __align__(16) unsigned char wsm[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned sv = smem_u32(wsm) + (unsigned)w * (DENSE ? WarpSm<HS>::PER_DENSE : WarpSm<HS>::PER);   // values / accumulators
  asm volatile("mov.u32 %0, %0;\n" : "+r"(sv));      // opaque: keep the shared-window offset in a register
                                                      // (rematerialised, it is an S2R + LEA per use)
  const unsigned sk = sv + HS * 8;                                        // keys
  const unsigned su = sv + (DENSE ? HS * 8 : HS * 12);                    // ring / sort buffer
  const unsigned rv = su, rc = su + RING_D * RING_C * 8;
  constexpr int SH = 32 - WarpSm<HS>::BITS;
  const int *list = tr.list[mytier];
  const int nl = tr.count[mytier];
  if (phase != 2) { xcol = ar.cols; xa = ar.vals; }
  for (;;) {
    int idx = 0;
    if (lane == 0) idx = atomicAdd(&tr.cursor[mytier], 1);
    idx = __shfl_sync(0xffffffffu, idx, 0);
    if (idx >= nl) break;
    const int i = list[idx];
    if (phase == 2 && tr.done[i] != mytier) continue;
    __syncwarp();
    int cmin = 0, span = 0;
    if (DENSE) {
      cmin = cminv[i]; span = spanv[i];
      for (int h = lane; h < span; h += 32) sts_f64(sv + 8 * h, 0.0);
    } else {
      for (int h = lane; h < HS; h += 32) sts_i32(sk + 4 * h, EMPTY);
    }
    __syncwarp();
    int filled = 0;
    bool gaveup = false;
    const int a0 = aro[i], a1 = aro[i + 1];
    // 32 entries of the row of A per batch (k, a_ik and the bounds of row k of B, one per lane);
    // the batch after the current one is already on its way
    int nb0 = 0, nb1 = 0;
    double nav = 0.0;
    if (a0 + lane < a1) { const int mk = acol[a0 + lane]; nav = aa[a0 + lane]; nb0 = bro[mk]; nb1 = bro[mk + 1]; }
    for (int jc = a0; jc < a1 && !gaveup; jc += 32) {
      const int mb0 = nb0, mb1 = nb1;
      const double mav = nav;
      nb0 = 0; nb1 = 0; nav = 0.0;
      { const int nx = jc + 32 + lane; if (nx < a1) { const int mk = acol[nx]; nav = aa[nx]; nb0 = bro[mk]; nb1 = bro[mk + 1]; } }
      const int ns = min(32, a1 - jc);
      // stage p of the ring: the first RING_C entries of row p of this batch, two per lane
#define RING_ISSUE(P)                                                                       \
  {                                                                                         \
    const int pb0 = __shfl_sync(0xffffffffu, mb0, (P)) + lane, pn = __shfl_sync(0xffffffffu, mb1, (P)) - pb0; \
    const unsigned so = (unsigned)(((P) % RING_D) * RING_C + lane);                         \
    if (pn > 0) { cp_async4s(rc + 4 * so, bcol + pb0); cp_async8s(rv + 8 * so, ba + pb0); } \
    if (pn > 32) { cp_async4s(rc + 4 * so + 128, bcol + pb0 + 32); cp_async8s(rv + 8 * so + 256, ba + pb0 + 32); } \
  }
#pragma unroll
      for (int p = 0; p < RING_D - 1; p++) { if (p < ns) RING_ISSUE(p); cp_async_commit(); }
      for (int st = 0; st < ns; st++) {
        if (st + RING_D - 1 < ns) RING_ISSUE(st + RING_D - 1);
        cp_async_commit();
        cp_async_wait<RING_D - 1>();
        const int b0 = __shfl_sync(0xffffffffu, mb0, st), len = __shfl_sync(0xffffffffu, mb1, st) - b0;
        const double av = __shfl_sync(0xffffffffu, mav, st);
        if (optimistic && filled + len > optimistic) { gaveup = true; break; }     // the table can never fill up
        const unsigned so = (unsigned)((st % RING_D) * RING_C);
        int mine = 0;
        for (int q = lane; q < len; q += 32) {
          int c;
          double v;
          if (q < RING_C) { c = lds_i32(rc + 4 * (so + q)); v = lds_f64(rv + 8 * (so + q)); }
          else { c = bcol[b0 + q]; v = ba[b0 + q]; }
          const double p = v * av;
          if (DENSE) { const unsigned a = sv + 8 * (unsigned)(c - cmin); sts_f64(a, lds_f64(a) + p); continue; }
          unsigned h = ((unsigned)c * 2654435761u) >> SH;
          for (;;) {
            int cur = lds_i32(sk + 4 * h);
            if (cur == EMPTY) {
              cur = atoms_cas(sk + 4 * h, EMPTY, c);
              if (cur == EMPTY) { double z = 0.0; z = z + p; sts_f64(sv + 8 * h, z); mine++; break; }
            }
            if (cur == c) { sts_f64(sv + 8 * h, lds_f64(sv + 8 * h) + p); break; }
            h = (h + 1) & (HS - 1);
          }
        }
        if (!DENSE && optimistic) filled += __reduce_add_sync(0xffffffffu, mine);
        __syncwarp();
      }
#undef RING_ISSUE
    }
    cp_async_wait<0>();
    if (gaveup) {
      if (lane == 0) hand_down(tr, mytier, i);
      continue;
    }
    if (DENSE) {
      // count, claim the row's place, write the non-zeros in column order
      int n = 0;
      for (int h = lane; h < span; h += 32) n += (lds_f64(sv + 8 * h) != 0.0);
      n = __reduce_add_sync(0xffffffffu, n);
      if (lane == 0) tr.done[i] = (signed char)mytier;
      if (phase == 1) { if (lane == 0) cnt[i] = n; continue; }
      long long rowbase = 0;
      if (phase == 3) {
        if (lane == 0) rowbase = arena_claim(ar, i, n, cnt);
        rowbase = __shfl_sync(0xffffffffu, rowbase, 0);
        if (rowbase < 0) continue;
      } else rowbase = xro[i];
      int run = 0;
      for (int h0 = 0; h0 < span; h0 += 32) {
        const int h = h0 + lane;
        const double v = h < span ? lds_f64(sv + 8 * h) : 0.0;
        const unsigned b = __ballot_sync(0xffffffffu, v != 0.0);
        if (v != 0.0) { const long long pos = rowbase + run + __popc(b & ((1u << lane) - 1u)); xcol[pos] = h + cmin; xa[pos] = v; }
        run += __popc(b);
      }
      continue;
    }
    // survivors (exact zeros leave the row): their slots, to be sorted by column
    int n = 0;
    for (int h = lane; h < HS; h += 32) {
      const bool live = (lds_i32(sk + 4 * h) != EMPTY) && (lds_f64(sv + 8 * h) != 0.0);
      const unsigned b = __ballot_sync(0xffffffffu, live);
      if (live) sts_u16(su + 2 * (n + __popc(b & ((1u << lane) - 1u))), (unsigned)h);
      n += __popc(b);
    }
    if (lane == 0) tr.done[i] = (signed char)mytier;
    if (phase == 1) { if (lane == 0) cnt[i] = n; continue; }
    long long rowbase = 0;
    if (phase == 3) {
      if (lane == 0) rowbase = arena_claim(ar, i, n, cnt);
      rowbase = __shfl_sync(0xffffffffu, rowbase, 0);
      if (rowbase < 0) continue;
    } else rowbase = xro[i];
    int P = 32;
    while (P < n) P <<= 1;
    for (int e = n + lane; e < P; e += 32) sts_u16(su + 2 * e, 0xffffu);      // padding sorts last
    __syncwarp();
    // bitonic network over the slots, compared by their columns (all different)
    for (int k = 2; k <= P; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int t = lane; t < (P >> 1); t += 32) {
          const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1)), hi = lo | j;
          const unsigned x = lds_u16(su + 2 * lo), y = lds_u16(su + 2 * hi);
          const int kx = x == 0xffffu ? EMPTY : lds_i32(sk + 4 * x), ky = y == 0xffffu ? EMPTY : lds_i32(sk + 4 * y);
          const bool up = ((lo & k) == 0);
          if ((kx > ky) == up && kx != ky) { sts_u16(su + 2 * lo, y); sts_u16(su + 2 * hi, x); }
        }
        __syncwarp();
      }
    for (int e = lane; e < n; e += 32) {
      const unsigned slot = lds_u16(su + 2 * e);
      xcol[rowbase + e] = lds_i32(sk + 4 * slot);
      xa[rowbase + e] = lds_f64(sv + 8 * slot);
    }
  }
}

// ---------------------------------------------------------------------------------------
// Block per row, for rows with thousands of distinct columns (Af*W on the coarse levels: 150
// k-steps of 500 entries each, 2000 distinct columns).  Table in shared memory (4096 or 8192
// slots, optimistic with an exact fill count) or in HBM (sized by the bound), bitmap over the
// row's column span for the ranking.  The rows of B arrive through a ring of BR_S stages filled
// by bulk async copies (one elected thread arms the stage's mbarrier with the byte count and
// issues the two copies of a row: values and columns, as 16-byte aligned windows); all threads
// wait on the stage's barrier, accumulate, meet at the step's __syncthreads (which the ordering
// rule needs anyway), and the elected thread refills the stage just freed.
// ---------------------------------------------------------------------------------------
// Ring geometry: BR_BYTES of shared memory cut into stages of `cap` entries (cap = the longest row
// of B, at most 768), up to 16 stages: short rows of B need many rows in flight to cover the
// latency of a bulk copy, long rows need few.
constexpr int BR_SMAX = 16, BR_CAPMAX = 768;
constexpr int BR_BYTES = 4 * ((BR_CAPMAX + 2) * 8 + (BR_CAPMAX + 8) * 4);
struct RingGeom { int stages, cap; };
static RingGeom ring_geom(int maxlb) {
  int cap = maxlb < BR_CAPMAX ? maxlb : BR_CAPMAX;
  cap = (cap + 7) / 8 * 8;
  if (cap < 32) cap = 32;
  int st = BR_BYTES / ((cap + 2) * 8 + (cap + 8) * 4);
  if (st > BR_SMAX) st = BR_SMAX;
  return RingGeom{st, cap};
}
// fill[st % 3] collects the new keys of step st: a step reads the (complete) count of the step
// before it and zeroes the counter of the step after it, so every thread sees the same total
struct BlockStage { int b0[256], b1[256]; double av[256]; int fill[3]; };

__device__ __forceinline__ void ring_issue(double *rv, int *rc, unsigned long long *bar, int s, int cap, int b0, int b1,
                                           const int *bcol, const double *ba) {
  const int L = min(b1 - b0, cap);
  if (L <= 0) return;
  const int sv = b0 & ~1, ev = (b0 + L + 1) & ~1, sc = b0 & ~3, ec = (b0 + L + 3) & ~3;
  const unsigned bv = (unsigned)(ev - sv) * 8u, bc = (unsigned)(ec - sc) * 4u;
  mbar_expect_tx(&bar[s], bv + bc);
  bulk_g2s(rv + s * (cap + 2), ba + sv, bv, &bar[s]);
  bulk_g2s(rc + s * (cap + 8), bcol + sc, bc, &bar[s]);
}

// DENSE: instead of a table, one fp64 accumulator per column of the row's span (coarse levels:
// a few thousand columns in all, rows that touch most of them).  No probing and no ranking: the
// columns come out in order; untouched and exactly cancelled entries are both 0.0 and both
// dropped, which is what mxm does.
template <bool DENSE>
__global__ void __launch_bounds__(256) k_spgemm_block(int phase, int mytier, Tiers tr, int HS_smem, int maxwords,
                                                      const int *cminv, const int *spanv, const i64 *toff,
                                                      int *gkeys, double *gvals, const int *aro, const int *acol,
                                                      const double *aa, const int *bro, const int *bcol,
                                                      const double *ba, int *cnt, const int *xro, int *xcol,
                                                      double *xa, int optimistic, RingGeom rg, Arena ar) {
  extern __shared__ __align__(16) unsigned char dsm[];
  __shared__ int sred, srow, stotal;
  __shared__ int stmp[256];
  __shared__ long long sbase;
  __shared__ BlockStage sg;
  __shared__ __align__(8) unsigned long long bar[BR_SMAX];
  const int t = threadIdx.x, T = blockDim.x;
  // dynamic shared memory: ring values | ring columns | [table values | table keys] | bitmap | prefix
  //                                              or  | dense accumulator over the span
  const int BR_S = rg.stages, BR_CAP = rg.cap;
  double *rv = (double *)dsm;
  int *rc = (int *)(dsm + (size_t)BR_S * (BR_CAP + 2) * 8);
  unsigned char *after = dsm + BR_BYTES;
  if (t == 0) {
    for (int s = 0; s < BR_S; s++) mbar_init(&bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  unsigned phasebits = 0;
  const int *list = tr.list[mytier];
  const int nl = tr.count[mytier];
  if (phase != 2) { xcol = ar.cols; xa = ar.vals; }
  BlockGroup g;
  for (;;) {
    __syncthreads();
    if (t == 0) srow = atomicAdd(&tr.cursor[mytier], 1);
    __syncthreads();
    const int q = srow;
    if (q >= nl) break;
    const int i = list[q];
    if (phase == 2 && tr.done[i] != mytier) continue;
    int HS = 0;
    double *svals = (double *)after;
    int *skeys = nullptr;
    unsigned *bits = nullptr;
    const int cmin = cminv[i], span = spanv[i];
    if (DENSE) {
      for (int c = t; c < span; c += T) svals[c] = 0.0;
    } else {
      if (HS_smem > 0) {
        HS = HS_smem; skeys = (int *)(after + (size_t)HS * 8); bits = (unsigned *)(after + (size_t)HS * 12);
      } else {
        const i64 tb = toff[q];                  // HBM tables are laid out by list position
        HS = (int)(toff[q + 1] - tb); svals = gvals + tb; skeys = gkeys + tb; bits = (unsigned *)after;
      }
      for (int h = t; h < HS; h += T) skeys[h] = EMPTY;
    }
    const unsigned mask = (unsigned)(HS - 1);
    const int sh = __clz(HS) + 1;                // HS is a power of two: 32 - log2(HS)
    if (t < 3) sg.fill[t] = 0;
    int filled = 0;                            // new keys of all completed steps (same in every thread)
    bool gaveup = false;
    const int a0 = aro[i], a1 = aro[i + 1];
    int gst = 0;                               // step number within the row
    for (int jc = a0; jc < a1 && !gaveup; jc += 256) {
      __syncthreads();
      const int my = jc + t;
      if (my < a1) { const int mk = acol[my]; sg.av[t] = aa[my]; sg.b0[t] = bro[mk]; sg.b1[t] = bro[mk + 1]; }
      __syncthreads();
      const int ns = min(256, a1 - jc);
      if (t == 0)
        for (int p = 0; p < BR_S && p < ns; p++) ring_issue(rv, rc, bar, p, BR_CAP, sg.b0[p], sg.b1[p], bcol, ba);
      for (int st = 0; st < ns; st++, gst++) {
        const int b0 = sg.b0[st], b1 = sg.b1[st], len = b1 - b0, s = st % BR_S;
        const double av = sg.av[st];
        if (!DENSE && optimistic) {
          if (gst > 0) filled += sg.fill[(gst + 2) % 3];           // the step before this one, complete
          if (t == 0) sg.fill[(gst + 1) % 3] = 0;
          if (filled + len > optimistic) {
            // give the row up; the copies already in flight must land before the ring is reused
            for (int p = st; p < ns && p < st + BR_S; p++)
              if (sg.b1[p] > sg.b0[p]) { const int ps = p % BR_S; mbar_wait(&bar[ps], (phasebits >> ps) & 1u); phasebits ^= (1u << ps); }
            gaveup = true;
            break;
          }
        }
        if (len > 0) { mbar_wait(&bar[s], (phasebits >> s) & 1u); phasebits ^= (1u << s); }
        const double *sv = rv + s * (BR_CAP + 2) + (b0 & 1);
        const int *sc = rc + s * (BR_CAP + 8) + (b0 & 3);
        int mine = 0;
        for (int e = t; e < len; e += T) {
          int c;
          double v;
          if (e < BR_CAP) { c = sc[e]; v = sv[e]; }
          else { c = bcol[b0 + e]; v = ba[b0 + e]; }
          if (DENSE) svals[c - cmin] = svals[c - cmin] + v * av;
          else mine += table_add(skeys, svals, mask, sh, c, v * av) ? 1 : 0;
        }
        if (!DENSE && optimistic && mine) atomicAdd(&sg.fill[gst % 3], mine);
        __syncthreads();
        if (t == 0 && st + BR_S < ns) ring_issue(rv, rc, bar, s, BR_CAP, sg.b0[st + BR_S], sg.b1[st + BR_S], bcol, ba);
      }
    }
    __syncthreads();
    if (gaveup) {
      if (t == 0) hand_down(tr, mytier, i);
      continue;
    }
    if (DENSE) {
      // contiguous column segments per thread: count, scan, write in order
      const int seg = (span + T - 1) / T;
      const int c0 = t * seg, c1 = min(span, c0 + seg);
      int mine = 0;
      for (int c = c0; c < c1; c++) mine += (svals[c] != 0.0);
      const int off = block_excl_scan(mine, stmp, &stotal);
      const int n = stotal;
      if (t == 0) tr.done[i] = (signed char)mytier;
      if (phase == 1) { if (t == 0) cnt[i] = n; continue; }
      long long rowbase = 0;
      if (phase == 3) {
        if (t == 0) sbase = arena_claim(ar, i, n, cnt);
        __syncthreads();
        rowbase = sbase;
        if (rowbase < 0) continue;
      } else rowbase = xro[i];
      long long p = rowbase + off;
      for (int c = c0; c < c1; c++) if (svals[c] != 0.0) { xcol[p] = c + cmin; xa[p] = svals[c]; p++; }
      continue;
    }
    int *wpre = (int *)(bits + maxwords);
    const int n = drop_zeros_count(g, skeys, svals, HS, &sred);
    if (t == 0) tr.done[i] = (signed char)mytier;
    if (phase == 1) { if (t == 0) cnt[i] = n; continue; }
    long long rowbase = 0;
    if (phase == 3) {
      if (t == 0) sbase = arena_claim(ar, i, n, cnt);
      __syncthreads();
      rowbase = sbase;
      if (rowbase < 0) continue;
    } else rowbase = xro[i];
    const int nw = (span + 31) / 32;
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
      xcol[rowbase + rank] = c; xa[rowbase + rank] = svals[h];
    }
  }
}

// last resort: table in HBM sized by the bound, bitonic sort of the whole table (rows whose span
// exceeds every bitmap)
__global__ void __launch_bounds__(256) k_spgemm_global(int phase, int mytier, Tiers tr, const i64 *toff, int *gkeys,
                                                       double *gvals, const int *aro, const int *acol,
                                                       const double *aa, const int *bro, const int *bcol,
                                                       const double *ba, int *cnt, const int *xro, int *xcol,
                                                       double *xa, Arena ar) {
  __shared__ int sred;
  __shared__ long long sbase;
  const int nl = tr.count[mytier];
  BlockGroup g;
  for (int q = blockIdx.x; q < nl; q += gridDim.x) {
    __syncthreads();
    const i64 base = toff[q];
    const int HS = (int)(toff[q + 1] - base);
    const int i = tr.list[mytier][q];
    if (threadIdx.x == 0) tr.done[i] = (signed char)mytier;
    row_phase(g, phase, i, aro, acol, aa, bro, bcol, ba, gkeys + base, gvals + base, HS, &sred, cnt, xro, xcol, xa, ar, &sbase);
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
// per-row bounds and the entry tier, G lanes per row of A
//   need  = min(sum of the B row lengths, columns of B, span): upper bound of the distinct columns
//   maxb  = longest B row: a lower bound (the entries of one row of B have distinct columns)
//   [cmin, cmin+span): the column range the row can touch (B rows are sorted: first/last entries)
constexpr int SPAN4 = 90000, SPAN8 = 330000, SPAN_HBM = 720000;   // bitmap = span/4 bytes of shared memory
constexpr int DSPAN8 = 8192, DSPAN = 22000;                       // dense accumulators: 8 bytes per column of the span
constexpr int T_TILE8 = 0, T_TILE32 = 1, T_W512 = 2, T_W2048 = 3, T_B4096 = 4, T_B8192 = 5, T_HBM = 6, T_GLOBAL = 7,
              T_DENSE8 = 8, T_DENSE = 9, T_WDENSE = 10, NTIER = 11;
constexpr int WDSPAN = 4608;                                      // warp-level dense accumulator: 36 KB per warp
template <int G>
__global__ void __launch_bounds__(256) k_spgemm_bin(int rn, const int *aro, const int *acol, const int *bro,
                                                    const int *bcol, int bcn, int small, Tiers tr, int *need,
                                                    int *cminv, int *spanv, unsigned char *route, int *stats,
                                                    unsigned long long *needsum, const signed char *hint) {
  const int i = blockIdx.x * (256 / G) + threadIdx.x / G;
  if (i >= rn) return;
  const int lane = threadIdx.x % G;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
  long long ub = 0;
  int lo = 0x7fffffff, hi = -1, lb = 0;
  for (int ja = aro[i] + lane; ja < aro[i + 1]; ja += G) {
    const int k = acol[ja], b0 = bro[k], b1 = bro[k + 1];
    ub += b1 - b0;
    lb = max(lb, b1 - b0);
    if (b1 > b0) { lo = min(lo, bcol[b0]); hi = max(hi, bcol[b1 - 1]); }
  }
  for (int off = G / 2; off >= 1; off >>= 1) {
    ub += __shfl_down_sync(gmask, ub, off, G);
    lb = max(lb, __shfl_down_sync(gmask, lb, off, G));
    lo = min(lo, __shfl_down_sync(gmask, lo, off, G));
    hi = max(hi, __shfl_down_sync(gmask, hi, off, G));
  }
  if (lane) return;
  const long long products = ub;
  if (ub > bcn) ub = bcn;
  const int span = hi >= lo ? hi - lo + 1 : 0;
  if (ub > span) ub = span;
  need[i] = (int)ub; cminv[i] = hi >= lo ? lo : 0; spanv[i] = span;
  const int r = span <= SPAN4 ? T_B4096 : span <= SPAN8 ? T_B8192 : span <= SPAN_HBM ? T_HBM : T_GLOBAL;
  route[i] = (unsigned char)r;
  int bin;
  if (ub <= 24) bin = T_TILE8;
  else if (ub <= 96) bin = T_TILE32;
  // touches most of its span with long steps (a block per row pays a barrier per k: the rows of B
  // must be long enough to keep 256 threads busy)
  else if (!small && span <= DSPAN && products * 2 >= span && products >= 128LL * (aro[i + 1] - aro[i]))
    bin = span <= DSPAN8 ? T_DENSE8 : T_DENSE;
  // the same with short steps: a warp per row
  else if (!small && span <= WDSPAN && products * 2 >= span) bin = T_WDENSE;
  else if (small && span <= DSPAN && (i & 3) == 3) bin = span <= WDSPAN ? T_WDENSE : (i & 4) ? T_DENSE8 : T_DENSE;   // test hook: a quarter of the rows
  else if (small || lb <= 192) bin = T_W512;     // test hook: every larger row walks down the whole ladder
  else if (lb <= 768) bin = T_W2048;
  else bin = r;
  // the tier that finished this row in the last product with the same left operand (Af*W0, Af*W,
  // Af*W of the next skeleton: the patterns grow slowly): start there instead of at the bottom of
  // the ladder.  A stale hint costs time, never correctness -- the row is still handed down.
  if (hint && !small && bin >= T_W512 && bin <= T_GLOBAL) {
    int ht = hint[i];
    if (ht > bin && ht <= T_GLOBAL) { if (ht >= T_B4096 && r > ht) ht = r; bin = ht; }
  }
  tr.list[bin][atomicAdd(&tr.count[bin], 1)] = i;
  if (bin == T_WDENSE) {}
  else if (bin >= T_DENSE8) atomicMax(&stats[bin == T_DENSE8 ? 2 : 3], span);     // stats[2]/[3]: largest dense span
  else if (bin >= T_W512) {
    atomicMax(&stats[r], span);                                  // stats[4..6]: largest span per route class
    atomicMax(&stats[r >= T_HBM ? 1 : 0], (int)ub);              // stats[0]/[1]: largest bound (route < / >= HBM)
  }
  if (ub) atomicAdd(needsum, (unsigned long long)ub);
  atomicMax(&stats[8], lb);                                      // longest row of B that is used
}

__global__ void k_table_sizes(const int *list, const int *count, const int *need, i64 *ts) {
  const int n = *count;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
    i64 s = 256;
    while (s < 2 * (i64)need[list[q]]) s <<= 1;
    ts[q] = s;
  }
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
// a pair of timing events that is either handed to the statistics or destroyed (error paths)
struct EventPair {
  cudaEvent_t a = nullptr, b = nullptr;
  EventPair() { CUDA_CHECK(cudaEventCreate(&a)); cudaError_t e = cudaEventCreate(&b); if (e != cudaSuccess) { cudaEventDestroy(a); a = nullptr; CUDA_CHECK(e); } }
  ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
  void keep() { g_stats.ev.emplace_back(a, b); a = b = nullptr; }
};

static int g_spgemm_impl = -1;
// entry-tier hints: which tier completed every row of the last product with left operand uid
// (a few entries: between Af*W0 and the next round's Af*W0 the pattern product W_skel*W_skel' runs)
struct TierHint { unsigned long long uid = 0; int rn = 0; unsigned long long stamp = 0; Buf<signed char> done; };
static TierHint g_hints[4];
static unsigned long long g_hint_clock = 0;
// diagnostics (amgb_debug_spgemm): rows per tier of the last product, entry counts and final counts
static bool g_collect_tiers = false;
static int g_last_tiers[2 * 11];
void spgemm_debug_collect(bool on) { g_collect_tiers = on; }
void spgemm_debug_tiers(int out[22]) { memcpy(out, g_last_tiers, sizeof g_last_tiers); }

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
// With the staged kernels the row-wise product handles long rows of A against short rows of B
// well enough that the two extra transposes no longer pay; the route is kept for
// AMGB_SPGEMM_TRANSPOSED=1 and the test hook.
static Csr g_At_cache;
static unsigned long long g_At_key = 0;     // Csr::uid of the cached operand (0 = empty)
void spgemm_cache_reset() { g_At_cache = Csr(); g_At_key = 0; for (auto &h : g_hints) h = TierHint(); }

Csr spgemm(const Csr &A, const Csr &B) {
  if (g_spgemm_impl < 0) { const char *e = getenv("AMGB_SPGEMM"); g_spgemm_impl = (e && !strcmp(e, "rowhash")) ? 0 : 1; }
  if (g_spgemm_impl == 0) return spgemm_rowhash(A, B);
  if (A.cn != B.rn) throw Error(-4, "spgemm: dimension mismatch");
  static int tr_on = -1;
  if (tr_on < 0) { const char *e = getenv("AMGB_SPGEMM_TRANSPOSED"); tr_on = (e && *e == '1') ? 1 : 0; }
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

template <class K>
static int blocks_per_sm(K kernel, int threads, size_t smem) {
  int nb = 0;
  CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, smem));
  return nb > 0 ? nb : 1;
}

static Csr spgemm_core_local(const Csr &A, const Csr &B) {
  StageTimer st_("prim.spgemm");
  Context &c = ctx();
  const int rn = A.rn;
  if (rn == 0) { Csr X(0, B.cn, 0); X.ro.zero(); return X; }
  const int *aro = A.ro.p, *acol = A.col.p, *bro = B.ro.p, *bcol = B.col.p;
  const double *aa = A.a.p, *ba = B.a.p;
  const bool small = test_small_bins();
  // ---- binning: entry tier of every row, bounds, spans
  Buf<int> lists((i64)NTIER * rn), meta(32), need(rn), cminv(rn), spanv(rn);
  Buf<unsigned char> route(rn);
  Buf<signed char> done(rn);
  Buf<unsigned long long> needsum(1);
  meta.zero(); needsum.zero(); done.zero();
  Tiers tr;
  for (int t = 0; t < NTIER; t++) tr.list[t] = lists.p + (i64)t * rn;
  tr.count = meta.p; tr.cursor = meta.p + NTIER; tr.done = done.p; tr.route = route.p;
  int *stats = meta.p + 2 * NTIER;
  static int hints_on = -1;
  if (hints_on < 0) { const char *e = getenv("AMGB_SPGEMM_HINTS"); hints_on = (e && *e == '0') ? 0 : 1; }
  const signed char *hint = nullptr;
  if (hints_on && A.uid != 0)
    for (auto &h : g_hints) if (h.uid == A.uid && h.rn == rn) hint = h.done.p;
  if ((double)A.nnz / rn > 12.0) k_spgemm_bin<8><<<(rn + 31) / 32, 256, 0, c.stream>>>(rn, aro, acol, bro, bcol, B.cn, small ? 1 : 0, tr, need.p, cminv.p, spanv.p, route.p, stats, needsum.p, hint);
  else k_spgemm_bin<1><<<(rn + 255) / 256, 256, 0, c.stream>>>(rn, aro, acol, bro, bcol, B.cn, small ? 1 : 0, tr, need.p, cminv.p, spanv.p, route.p, stats, needsum.p, hint);
  c.launches++; post_launch("spgemm_bin");
  int hm[32];
  unsigned long long need_total_u = 0;
  d2h_async(hm, meta.p, sizeof hm);
  d2h_async(&need_total_u, needsum.p, sizeof need_total_u);
  stream_sync();
  const int *hc = hm;                       // entry counts per tier
  const i64 need_total = (i64)need_total_u;
  const int *hs = hm + 2 * NTIER;
  const int maxneed_lo = hs[0], maxneed_hi = hs[1], dspan8 = hs[2], dspan = hs[3];
  const RingGeom rg = ring_geom(hs[8]);
  const int span4 = hs[T_B4096], span5 = std::max(span4, hs[T_B8192]), span6 = std::max(span5, hs[T_HBM]);
  auto words = [](int span) { return (span + 31) / 32; };

  static bool attr = false;
  if (!attr) {
    CUDA_CHECK(cudaFuncSetAttribute((const void *)k_spgemm_block<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    CUDA_CHECK(cudaFuncSetAttribute((const void *)k_spgemm_block<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr = true;
  }
  Buf<int> cnt(rn + 1), xro(rn + 1);
  // ---- arena of the fused pass: every row is written, already sorted, at an offset taken from a
  // bump pointer, and copied into place once the row offsets are known
  static int fused_on = -1;
  if (fused_on < 0) { const char *e = getenv("AMGB_SPGEMM_FUSED"); fused_on = (e && *e == '0') ? 0 : 1; }
  Arena ar{nullptr, nullptr, nullptr, 0, nullptr, nullptr};
  Buf<int> acols, aflag;
  Buf<double> avals;
  Buf<unsigned long long> atop;
  Buf<i64> aroff;
  if (fused_on) {
    // sized by the sum of the per-row bounds when that is affordable (the pattern products
    // W_skel*W_skel' expand 15x), else a multiple of the operands; a too small arena costs a
    // second pass over all rows
    i64 cap = 6 * (A.nnz + B.nnz) + rn;
    if (cap < (1 << 20)) cap = 1 << 20;
    const i64 cap_hi = (i64)400 << 20;             // 400 Mi entries = 4.8 GB
    if (need_total > cap) cap = need_total < cap_hi ? need_total : (cap > cap_hi ? cap : cap_hi);
    if (need_total < cap) cap = need_total > 0 ? need_total : 1;
    if (test_force('o')) cap = 48;                 // test hook: (nearly) every row overflows the arena
    acols.alloc(cap); avals.alloc(cap); atop.alloc(1); aflag.alloc(1); aroff.alloc(rn);
    atop.zero(); aflag.zero();
    ar = Arena{acols.p, avals.p, atop.p, cap, aroff.p, aflag.p};
  }
  // optimistic limits (test hook: tiny, so that small problems walk through every tier)
  // (5/8 of the slots: linear probing degrades quickly above that)
  const int lim512 = small ? 20 : 512 * 5 / 8, lim2048 = small ? 40 : 2048 * 5 / 8;
  const int lim4096 = small ? 60 : 4096 * 5 / 8, lim8192 = small ? 80 : 8192 * 5 / 8;
  // can any row reach the HBM tiers?  (only then their lists are read back to size the tables)
  const bool hbm_possible = small || hc[T_HBM] || hc[T_GLOBAL] || maxneed_lo > lim8192 || maxneed_hi > lim2048;
  Buf<i64> tsz6, toff6, tsz7, toff7;
  Buf<int> gkeys6, gkeys7;
  Buf<double> gvals6, gvals7;
  int n6 = 0, n7 = 0;
  Csr X;
  EventPair ev1, ev2;
  CUDA_CHECK(cudaEventRecord(ev1.a, c.stream));
  bool done_all = false;
  for (int pass = 0; pass < 2 && !done_all; pass++) {
    const int phase = fused_on ? (pass == 0 ? 3 : 2) : pass + 1;
    const bool first = (pass == 0);
    int *xcol = phase == 2 ? X.col.p : nullptr;
    double *xa = phase == 2 ? X.a.p : nullptr;
    if (!first) dev_memset(tr.cursor, 0, NTIER * sizeof(int));
    if (hc[T_TILE8]) {
      k_spgemm_tile<8, 64, 24><<<(hc[T_TILE8] + 31) / 32, 256, 0, c.stream>>>(phase, tr.list[T_TILE8], hc[T_TILE8], aro, acol, aa, bro, bcol, ba, cnt.p, xro.p, xcol, xa, ar);
      c.launches++; post_launch("spgemm_tile8");
    }
    if (hc[T_TILE32]) {
      k_spgemm_tile<32, 256, 96><<<(hc[T_TILE32] + 7) / 8, 256, 0, c.stream>>>(phase, tr.list[T_TILE32], hc[T_TILE32], aro, acol, aa, bro, bcol, ba, cnt.p, xro.p, xcol, xa, ar);
      c.launches++; post_launch("spgemm_tile32");
    }
    for (int t = T_DENSE8; t <= T_DENSE; t++) {
      if (!hc[t]) continue;
      const size_t sm = BR_BYTES + (size_t)(t == T_DENSE8 ? dspan8 : dspan) * 8;
      const int bps = blocks_per_sm(k_spgemm_block<true>, 256, sm);
      k_spgemm_block<true><<<std::min(hc[t], c.sm_count * bps), 256, sm, c.stream>>>(phase, t, tr, 0, 0, cminv.p, spanv.p, nullptr, nullptr, nullptr, aro, acol, aa, bro, bcol, ba, cnt.p, xro.p, xcol, xa, 0, rg, ar);
      c.launches++; post_launch("spgemm_block_dense");
    }
    if (hc[T_WDENSE]) {
      constexpr size_t sm = WarpSm<WDSPAN>::PER_DENSE;
      static int bps = 0;
      if (!bps) {
        CUDA_CHECK(cudaFuncSetAttribute((const void *)k_spgemm_warp<WDSPAN, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        bps = blocks_per_sm(k_spgemm_warp<WDSPAN, 1, true>, 32, sm);
      }
      k_spgemm_warp<WDSPAN, 1, true><<<std::min(hc[T_WDENSE], c.sm_count * bps), 32, sm, c.stream>>>(phase, T_WDENSE, tr, aro, acol, aa, bro, bcol, ba, cnt.p, xro.p, xcol, xa, 0, cminv.p, spanv.p, ar);
      c.launches++; post_launch("spgemm_warp_dense");
    }
    // the ladder: every tier also takes the rows handed down by the tiers above it (first pass)
    const bool lower = hc[T_W512] || hc[T_W2048];
    if (hc[T_W512]) {
      constexpr size_t sm = 4 * WarpSm<512>::PER;
      static int bps = 0;
      if (!bps) bps = blocks_per_sm(k_spgemm_warp<512, 4, false>, 128, sm);
      const int grid = std::min((hc[T_W512] + 3) / 4, c.sm_count * bps);
      k_spgemm_warp<512, 4, false><<<grid, 128, sm, c.stream>>>(phase, T_W512, tr, aro, acol, aa, bro, bcol, ba, cnt.p, xro.p, xcol, xa, first ? lim512 : 0, cminv.p, spanv.p, ar);
      c.launches++; post_launch("spgemm_warp512");
    }
    if (lower) {
      constexpr size_t sm = WarpSm<2048>::PER;
      static int bps = 0;
      if (!bps) bps = blocks_per_sm(k_spgemm_warp<2048, 1, false>, 32, sm);
      k_spgemm_warp<2048, 1, false><<<c.sm_count * bps, 32, sm, c.stream>>>(phase, T_W2048, tr, aro, acol, aa, bro, bcol, ba, cnt.p, xro.p, xcol, xa, first ? lim2048 : 0, cminv.p, spanv.p, ar);
      c.launches++; post_launch("spgemm_warp2048");
    }
    const bool to4 = hc[T_B4096] || (lower && (small || maxneed_lo > lim2048) && span4 > 0);
    if (to4) {
      const int mw = words(span4);
      const size_t sm = BR_BYTES + (size_t)4096 * 12 + (size_t)mw * 8;
      const int bps = blocks_per_sm(k_spgemm_block<false>, 256, sm);
      k_spgemm_block<false><<<c.sm_count * bps, 256, sm, c.stream>>>(phase, T_B4096, tr, 4096, mw, cminv.p, spanv.p, nullptr, nullptr, nullptr, aro, acol, aa, bro, bcol, ba, cnt.p, xro.p, xcol, xa, first ? lim4096 : 0, rg, ar);
      c.launches++; post_launch("spgemm_block4096");
    }
    const bool to5 = hc[T_B8192] || to4 || (lower && (small || maxneed_lo > lim2048) && span5 > 0);
    if (to5) {
      const int mw = words(span5);
      const size_t sm = BR_BYTES + (size_t)8192 * 12 + (size_t)mw * 8;
      k_spgemm_block<false><<<c.sm_count, 256, sm, c.stream>>>(phase, T_B8192, tr, 8192, mw, cminv.p, spanv.p, nullptr, nullptr, nullptr, aro, acol, aa, bro, bcol, ba, cnt.p, xro.p, xcol, xa, first ? lim8192 : 0, rg, ar);
      c.launches++; post_launch("spgemm_block8192");
    }
    if (hbm_possible) {
      if (first) {
        int h2[2];
        d2h(h2, tr.count + T_HBM, sizeof h2);        // rows that reached the HBM tiers
        n6 = h2[0]; n7 = h2[1];
        for (int t = T_HBM; t <= T_GLOBAL; t++) {
          const int nt = t == T_HBM ? n6 : n7;
          if (!nt) continue;
          Buf<i64> &tsz = t == T_HBM ? tsz6 : tsz7, &toff = t == T_HBM ? toff6 : toff7;
          tsz.alloc(nt + 1); toff.alloc(nt + 1);
          k_table_sizes<<<(nt + 255) / 256, 256, 0, c.stream>>>(tr.list[t], tr.count + t, need.p, tsz.p);
          c.launches++; post_launch("spgemm_table_sizes");
          const i64 total = exclusive_scan64(tsz.p, toff.p, nt);
          (t == T_HBM ? gkeys6 : gkeys7).alloc(total);
          (t == T_HBM ? gvals6 : gvals7).alloc(total);
        }
      }
      if (n6) {
        const int mw = words(span6);
        const size_t sm = BR_BYTES + (size_t)mw * 8;
        const int bps = blocks_per_sm(k_spgemm_block<false>, 256, sm);
        k_spgemm_block<false><<<std::min(n6, c.sm_count * bps), 256, sm, c.stream>>>(phase, T_HBM, tr, 0, mw, cminv.p, spanv.p, toff6.p, gkeys6.p, gvals6.p, aro, acol, aa, bro, bcol, ba, cnt.p, xro.p, xcol, xa, 0, rg, ar);
        c.launches++; post_launch("spgemm_block_hbm");
      }
      if (n7) {
        k_spgemm_global<<<std::min(n7, c.sm_count * 8), 256, 0, c.stream>>>(phase, T_GLOBAL, tr, toff7.p, gkeys7.p, gvals7.p, aro, acol, aa, bro, bcol, ba, cnt.p, xro.p, xcol, xa, ar);
        c.launches++; post_launch("spgemm_global");
      }
    }
    if (first) {
      CUDA_CHECK(cudaEventRecord(ev1.b, c.stream));
      exclusive_scan_dev(cnt.p, xro.p, rn);
      int tot = 0, ovf = 0;
      d2h_async(&tot, xro.p + rn, sizeof(int));
      if (fused_on) d2h_async(&ovf, aflag.p, sizeof(int));
      stream_sync();
      const i64 nnz = tot;
      X = Csr(rn, B.cn, nnz);
      d2d(X.ro.p, xro.p, sizeof(int) * (size_t)(rn + 1));
      CUDA_CHECK(cudaEventRecord(ev2.a, c.stream));
      if (fused_on && ovf == 0) {       // every row is in the arena: copy into place
        if (nnz > 0) {
          if ((double)nnz / (double)rn <= 12.0)
            k_arena_copy<8><<<(rn + 31) / 32, 256, 0, c.stream>>>(rn, xro.p, aroff.p, acols.p, avals.p, X.col.p, X.a.p);
          else
            k_arena_copy<32><<<(rn + 7) / 8, 256, 0, c.stream>>>(rn, xro.p, aroff.p, acols.p, avals.p, X.col.p, X.a.p);
          c.launches++; post_launch("spgemm_arena_copy");
        }
        done_all = true;
      }
    }
  }
  CUDA_CHECK(cudaEventRecord(ev2.b, c.stream));      // the copy, or the second pass
  ev1.keep(); ev2.keep();
  {
    static int logit = -1;
    if (logit < 0) { const char *e = getenv("AMGB_SPGEMM_LOG"); logit = (e && *e && *e != '0') ? 1 : 0; }
    if (logit) {
      stream_sync();
      float m1 = 0, m2 = 0;
      const size_t ne = g_stats.ev.size();
      cudaEventElapsedTime(&m1, g_stats.ev[ne - 2].first, g_stats.ev[ne - 2].second);
      cudaEventElapsedTime(&m2, g_stats.ev[ne - 1].first, g_stats.ev[ne - 1].second);
      m1 += m2;
      int hcnt[NTIER];
      d2h(hcnt, tr.count, sizeof hcnt);
      fprintf(stderr, "spgemm A %dx%d nnz %lld  B %dx%d nnz %lld  X nnz %lld | tiers in %d %d %d %d %d %d %d %d d %d %d | total %d %d %d %d %d %d %d %d | %.3f ms | %.1f GB/s\n",
              A.rn, A.cn, (long long)A.nnz, B.rn, B.cn, (long long)B.nnz, (long long)X.nnz, hc[0], hc[1], hc[2], hc[3], hc[4],
              hc[5], hc[6], hc[7], hc[8], hc[9] * 1000000 + hc[10], hcnt[0], hcnt[1], hcnt[2], hcnt[3], hcnt[4], hcnt[5], hcnt[6], hcnt[7], m1,
              (12.0 * (A.nnz + B.nnz + X.nnz)) / (m1 * 1e-3) / 1e9);
    }
  }
  if (hints_on && A.uid != 0) {          // replace this operand's entry, else the oldest one
    TierHint *slot = &g_hints[0];
    for (auto &h : g_hints) { if (h.uid == A.uid) { slot = &h; break; } if (h.stamp < slot->stamp) slot = &h; }
    slot->uid = A.uid; slot->rn = rn; slot->stamp = ++g_hint_clock; slot->done = std::move(done);
  }
  if (g_collect_tiers) {
    int hcnt[NTIER];
    d2h(hcnt, tr.count, sizeof hcnt);
    for (int t = 0; t < NTIER; t++) { g_last_tiers[t] = hc[t]; g_last_tiers[NTIER + t] = hcnt[t]; }
  }
  // algorithmic bytes: A, B and X each moved once (12 B per entry, 4 B per row offset)
  g_stats.bytes += 12 * (A.nnz + B.nnz + X.nnz) + 4 * ((i64)A.rn + B.rn + X.rn + 3);
  g_stats.calls++;
  return X;
}
#else
void spgemm_stats_reset() {}
void spgemm_debug_collect(bool) {}
void spgemm_debug_tiers(int out[22]) { memset(out, 0, 22 * sizeof(int)); }
void spgemm_cache_reset() {}
void spgemm_stats_get(double *seconds, i64 *bytes, i64 *calls) { *seconds = 0; *bytes = 0; *calls = 0; }
Csr spgemm(const Csr &A, const Csr &B) {
  if (comm_active() && A.rn >= comm_size() && A.nnz + B.nnz >= comm_min_work())
    return spgemm_partitioned(A, B, spgemm_rowhash);
  return spgemm_rowhash(A, B);
}
#endif

}  // namespace amgb

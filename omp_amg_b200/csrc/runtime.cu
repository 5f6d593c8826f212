// runtime.cu -- context, stream-ordered memory, prefix scans, deterministic reductions.
#include "common.cuh"
#include "comm.cuh"
#include <algorithm>
#include <chrono>
#include <map>

namespace amgb {

static Context g_ctx;
Context &ctx() { return g_ctx; }

#ifndef AMGB_EMU
// =======================================================================================
// CUDA build
// =======================================================================================
static bool g_inited = false;
static int g_debug_sync = -1;
void post_launch(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && g_debug_sync > 0) e = cudaStreamSynchronize(g_ctx.stream);
  if (e != cudaSuccess)
    throw Error(-100, std::string("CUDA kernel '") + what + "' failed: " + cudaGetErrorString(e));
}
void ctx_init(int device) {
  if (g_inited) return;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    throw Error(-101, "omp_amg_b200: no CUDA device available (this library has no CPU path)");
  if (g_debug_sync < 0) { const char *e = getenv("AMGB_DEBUG_SYNC"); g_debug_sync = (e && *e && *e != '0') ? 1 : 0; }
  { const char *e = getenv("AMGB_REDUCE"); if (e && !strcmp(e, "tree")) g_ctx.reduce_seq = 0; else if (e && !strcmp(e, "seq")) g_ctx.reduce_seq = 1; }
  if (device >= 0) CUDA_CHECK(cudaSetDevice(device));
  int dev = 0;
  CUDA_CHECK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
  g_ctx.sm_count = prop.multiProcessorCount;
  CUDA_CHECK(cudaStreamCreateWithFlags(&g_ctx.stream, cudaStreamNonBlocking));
  cudaMemPool_t pool;
  CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, dev));
  unsigned long long thr = ~0ULL;   // keep freed blocks cached in the pool
  CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
  g_inited = true;
}
// Device memory: a caching allocator over cudaMalloc.  A setup makes ~10^5 allocations whose sizes
// repeat from level to level and from setup to setup; the stream-ordered pool of the driver
// re-maps physical memory when it fragments, which showed up as 10x run-to-run swings in kernels
// that allocate several large temporaries (transpose).  Blocks are rounded up to 1/8-octave size
// classes and kept in per-class free lists; all work is on one stream, so a freed block may be
// handed out again immediately.  AMGB_ALLOC=async selects the driver pool instead.
static std::multimap<size_t, void *> g_free_blocks;
static std::map<void *, size_t> g_block_size;
static size_t g_cached_bytes = 0, g_live_bytes = 0, g_peak_bytes = 0;
static int g_alloc_async = -1;
static size_t size_class(size_t b) {
  if (b < 512) return 512;
  size_t p = 512;
  while (p * 2 <= b) p *= 2;               // p <= b < 2p
  const size_t step = p / 8;
  return ((b + step - 1) / step) * step;
}
void dev_release_cache() {
  for (auto &kv : g_free_blocks) { cudaFree(kv.second); g_block_size.erase(kv.second); }
  g_free_blocks.clear();
  g_cached_bytes = 0;
}
void *dev_alloc(size_t bytes) {
  if (g_alloc_async < 0) { const char *e = getenv("AMGB_ALLOC"); g_alloc_async = (e && !strcmp(e, "async")) ? 1 : 0; }
  void *p = nullptr;
  if (g_alloc_async) {
    CUDA_CHECK(cudaMallocAsync(&p, bytes ? bytes : 8, g_ctx.stream));
    return p;
  }
  const size_t sz = size_class(bytes ? bytes : 8);
  auto it = g_free_blocks.find(sz);
  if (it != g_free_blocks.end()) {
    p = it->second;
    g_free_blocks.erase(it);
    g_cached_bytes -= sz;
  } else {
    cudaError_t e = cudaMalloc(&p, sz);
    if (e != cudaSuccess) {                 // give the cache back to the driver and retry once
      cudaGetLastError();
      CUDA_CHECK(cudaStreamSynchronize(g_ctx.stream));
      dev_release_cache();
      e = cudaMalloc(&p, sz);
      if (e != cudaSuccess)
        throw Error(-102, std::string("out of device memory allocating ") + std::to_string(sz) + " bytes");
    }
    g_block_size[p] = sz;
  }
  g_live_bytes += sz;
  if (g_live_bytes > g_peak_bytes) g_peak_bytes = g_live_bytes;
  return p;
}
void dev_free(void *p) {
  if (!p) return;
  if (g_alloc_async) { cudaFreeAsync(p, g_ctx.stream); return; }
  auto it = g_block_size.find(p);
  if (it == g_block_size.end()) { cudaFree(p); return; }
  g_free_blocks.emplace(it->second, p);
  g_cached_bytes += it->second;
  g_live_bytes -= it->second;
}
size_t dev_peak_bytes() { return g_peak_bytes; }
void dev_memset(void *p, int v, size_t bytes) { if (bytes) CUDA_CHECK(cudaMemsetAsync(p, v, bytes, g_ctx.stream)); }
void h2d(void *dst, const void *src, size_t bytes) {
  if (bytes) CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, g_ctx.stream));
}
void d2h(void *dst, const void *src, size_t bytes) {
  if (bytes) CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g_ctx.stream));
  stream_sync();
}
void d2h_async(void *dst, const void *src, size_t bytes) {       // the caller synchronises
  if (bytes) CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g_ctx.stream));
}
void d2d(void *dst, const void *src, size_t bytes) {
  if (bytes) CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, g_ctx.stream));
}
void stream_sync() { CUDA_CHECK(cudaStreamSynchronize(g_ctx.stream)); g_ctx.syncs++; }

// ---- exclusive scan: 1024 items per block, block sums scanned recursively ----
template <class T>
__global__ void __launch_bounds__(256) k_scan_block(const T *in, T *out, T *bsum, i64 n) {
  __shared__ T warp_tot[8];
  const i64 base = (i64)blockIdx.x * 1024 + threadIdx.x * 4;
  T v[4], s = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) { v[k] = (base + k < n) ? in[base + k] : (T)0; s += v[k]; }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  T incl = s;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    T t = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += t;
  }
  if (lane == 31) warp_tot[w] = incl;
  __syncthreads();
  T woff = 0;
  for (int k = 0; k < w; k++) woff += warp_tot[k];
  T excl = woff + incl - s;
#pragma unroll
  for (int k = 0; k < 4; k++) { if (base + k < n) out[base + k] = excl; excl += v[k]; }
  if (threadIdx.x == 255 && bsum) bsum[blockIdx.x] = woff + incl;
}
template <class T>
__global__ void __launch_bounds__(256) k_scan_add(T *out, const T *boff, i64 n) {
  i64 i = (i64)blockIdx.x * 1024 + threadIdx.x;
  T o = boff[blockIdx.x];
#pragma unroll
  for (int k = 0; k < 4; k++, i += 256) if (i < n) out[i] += o;
}
template <class T>
static void scan_rec(const T *in, T *out, i64 n) {     // out[0..n) exclusive; out[n] untouched
  if (n <= 0) return;
  i64 nb = (n + 1023) / 1024;
  if (nb == 1) {
    k_scan_block<T><<<1, 256, 0, g_ctx.stream>>>(in, out, (T *)nullptr, n);
    g_ctx.launches++; post_launch(__func__);
    return;
  }
  Buf<T> bsum(nb), boff(nb);
  k_scan_block<T><<<(unsigned)nb, 256, 0, g_ctx.stream>>>(in, out, bsum.p, n);
  g_ctx.launches++; post_launch(__func__);
  scan_rec<T>(bsum.p, boff.p, nb);
  k_scan_add<T><<<(unsigned)nb, 256, 0, g_ctx.stream>>>(out, boff.p, n);
  g_ctx.launches++; post_launch(__func__);
}
template <class T>
__global__ void k_scan_last(const T *in, T *out, i64 n) { out[n] = out[n - 1] + in[n - 1]; }
template <class T>
static T scan_total(const T *in, T *out, i64 n) {
  if (n <= 0) { T z = 0; h2d(out, &z, sizeof(T)); stream_sync(); return 0; }
  scan_rec<T>(in, out, n);
  k_scan_last<T><<<1, 1, 0, g_ctx.stream>>>(in, out, n);
  g_ctx.launches++; post_launch(__func__);
  T tot;
  d2h(&tot, out + n, sizeof(T));
  return tot;
}
i64 exclusive_scan(const int *in, int *out, i64 n) { return (i64)scan_total<int>(in, out, n); }
void exclusive_scan_dev(const int *in, int *out, i64 n) {      // total left in out[n], no read-back
  if (n <= 0) { dev_memset(out, 0, sizeof(int)); return; }
  scan_rec<int>(in, out, n);
  k_scan_last<int><<<1, 1, 0, g_ctx.stream>>>(in, out, n);
  g_ctx.launches++; post_launch(__func__);
}
i64 exclusive_scan64(const i64 *in, i64 *out, i64 n) { return scan_total<i64>(in, out, n); }

// ---- deterministic tree sum ----
__device__ __forceinline__ double chunk_tree(double x) {
  __shared__ double wsum[8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) x = __dadd_rn(x, __shfl_down_sync(0xffffffffu, x, off));
  if (lane == 0) wsum[w] = x;
  __syncthreads();
  double y = 0.0;
  if (w == 0) {
    y = (lane < 8) ? wsum[lane] : 0.0;
#pragma unroll
    for (int off = 4; off >= 1; off >>= 1) y = __dadd_rn(y, __shfl_down_sync(0xffffffffu, y, off));
  }
  return y;   // valid in thread 0
}
__global__ void __launch_bounds__(256) k_tree_sum(const double *v, i64 n, double *out) {
  const i64 base = (i64)blockIdx.x * 1024 + threadIdx.x;
  double x = (base < n) ? v[base] : 0.0;
  x = __dadd_rn(x, (base + 256 < n) ? v[base + 256] : 0.0);
  x = __dadd_rn(x, (base + 512 < n) ? v[base + 512] : 0.0);
  x = __dadd_rn(x, (base + 768 < n) ? v[base + 768] : 0.0);
  double r = chunk_tree(x);
  if (threadIdx.x == 0) out[blockIdx.x] = r;
}
__global__ void __launch_bounds__(256) k_tree_dot(const double *a, const double *b, i64 n, double *out) {
  const i64 base = (i64)blockIdx.x * 1024 + threadIdx.x;
  double x = (base < n) ? __dmul_rn(a[base], b[base]) : 0.0;
  x = __dadd_rn(x, (base + 256 < n) ? __dmul_rn(a[base + 256], b[base + 256]) : 0.0);
  x = __dadd_rn(x, (base + 512 < n) ? __dmul_rn(a[base + 512], b[base + 512]) : 0.0);
  x = __dadd_rn(x, (base + 768 < n) ? __dmul_rn(a[base + 768], b[base + 768]) : 0.0);
  double r = chunk_tree(x);
  if (threadIdx.x == 0) out[blockIdx.x] = r;
}
static void tree_finish(Buf<double> &part, i64 nc, double *out) {
  while (nc > 1) {
    i64 nc2 = (nc + 1023) / 1024;
    Buf<double> nxt(nc2);
    k_tree_sum<<<(unsigned)nc2, 256, 0, g_ctx.stream>>>(part.p, nc, nxt.p);
    g_ctx.launches++; post_launch(__func__);
    part = std::move(nxt);
    nc = nc2;
  }
  d2d(out, part.p, sizeof(double));
}
// result in the device scalar *out; b == nullptr: plain sum
void tree_dot_dev(double *out, const double *a, const double *b, i64 n) {
  if (n <= 0) { dev_memset(out, 0, sizeof(double)); return; }
  i64 nc = (n + 1023) / 1024;
  Buf<double> part(nc);
  if (b) k_tree_dot<<<(unsigned)nc, 256, 0, g_ctx.stream>>>(a, b, n, part.p);
  else k_tree_sum<<<(unsigned)nc, 256, 0, g_ctx.stream>>>(a, n, part.p);
  g_ctx.launches++; post_launch(__func__);
  tree_finish(part, nc, out);
}
double tree_sum(const double *v, i64 n) {
  if (n <= 0) return 0.0;
  Buf<double> out(1);
  tree_dot_dev(out.p, v, nullptr, n);
  return out.get(0);
}
double tree_dot(const double *a, const double *b, i64 n) {
  if (n <= 0) return 0.0;
  Buf<double> out(1);
  tree_dot_dev(out.p, a, b, n);
  return out.get(0);
}

// ---- left-to-right sum: thread 0 carries the running sum (one rounding per element, in index
// order, exactly as the reference's loops); the other 31 warps stage the next chunk of products
// in shared memory so that the chain never waits for HBM ----
#define SEQ_CHUNK 2048
__global__ void __launch_bounds__(1024) k_seq_dot(const double *a, const double *b, i64 n, double *out) {
  __shared__ double buf[2][SEQ_CHUNK];
  const int t = threadIdx.x;
  const i64 nchunks = (n + SEQ_CHUNK - 1) / SEQ_CHUNK;
  for (int j = t; j < SEQ_CHUNK && j < n; j += 1024) buf[0][j] = b ? __dmul_rn(a[j], b[j]) : a[j];
  double r = 0;
  for (i64 c = 0; c < nchunks; c++) {
    __syncthreads();
    const int cur = (int)(c & 1);
    if (t >= 32) {
      const i64 base = (c + 1) * SEQ_CHUNK;
      for (i64 j = base + (t - 32); j < base + SEQ_CHUNK && j < n; j += 992)
        buf[cur ^ 1][j - base] = b ? __dmul_rn(a[j], b[j]) : a[j];
    } else if (t == 0) {
      const i64 base = c * SEQ_CHUNK;
      const int len = (int)((n - base < SEQ_CHUNK) ? (n - base) : SEQ_CHUNK);
      const double *p = buf[cur];
      int j = 0;
      if (len >= 8) {
        // software pipeline: the next eight values are in registers before the chain needs them
        double a0 = p[0], a1 = p[1], a2 = p[2], a3 = p[3], a4 = p[4], a5 = p[5], a6 = p[6], a7 = p[7];
        for (; j + 16 <= len; j += 8) {
          const double b0 = p[j + 8], b1 = p[j + 9], b2 = p[j + 10], b3 = p[j + 11], b4 = p[j + 12],
                       b5 = p[j + 13], b6 = p[j + 14], b7 = p[j + 15];
          r = __dadd_rn(r, a0); r = __dadd_rn(r, a1); r = __dadd_rn(r, a2); r = __dadd_rn(r, a3);
          r = __dadd_rn(r, a4); r = __dadd_rn(r, a5); r = __dadd_rn(r, a6); r = __dadd_rn(r, a7);
          a0 = b0; a1 = b1; a2 = b2; a3 = b3; a4 = b4; a5 = b5; a6 = b6; a7 = b7;
        }
        r = __dadd_rn(r, a0); r = __dadd_rn(r, a1); r = __dadd_rn(r, a2); r = __dadd_rn(r, a3);
        r = __dadd_rn(r, a4); r = __dadd_rn(r, a5); r = __dadd_rn(r, a6); r = __dadd_rn(r, a7);
        j += 8;
      }
      for (; j < len; j++) r = __dadd_rn(r, p[j]);
    }
  }
  if (t == 0) *out = r;
}
// ---------------------------------------------------------------------------------------
// Exact left-to-right sum in parallel.
//
// The reference adds the terms one by one, rounding after every addition.  While the running
// sum s stays inside one binade (same sign, same exponent e, ulp u = 2^(e-52)) every step is
//     S' = S + r(x),  S = |s|/u (a 53-bit integer),  x = (+-p)/u,
// where r(x) is x rounded to the nearest integer -- independent of S -- except for exact ties
// (fractional part 1/2), which go to the even result and therefore depend on the PARITY of S only.
// So a chunk of terms is reduced by (1) per-term integer/fraction split of x, (2) a scan over the
// two-state parity automaton (a tie resets the parity to even, any other term xors it with its
// increment), (3) an integer prefix sum.  The first term at which the running sum could leave the
// binade (or is not finite/normal) ends the chunk: everything before it is exact, that term and a
// short adaptive burst after it are added the plain way by one thread, and the next chunk starts
// in the new binade.  For the sums this library needs (norms, r'z, p'Ap: essentially positive
// terms) such events are O(log n), so the cost is ~n/8192 block-wide scans instead of n dependent
// additions; in the worst case (a sum hovering around zero) it degrades to the plain chain.
// The result is bit-identical to the sequential loop for every input (tests/test_gpu_parity.py).
// ---------------------------------------------------------------------------------------
#define EPS_T 1024
#define EPS_E 8
#define EPS_BURST_MAX 2048
struct ParFn { unsigned char reset, x; };   // parity_out = reset ? x : parity_in ^ x
__device__ __forceinline__ ParFn par_compose(ParFn f, ParFn g) {   // apply f, then g
  ParFn r;
  r.reset = f.reset | g.reset;
  r.x = g.reset ? g.x : (unsigned char)(f.x ^ g.x);
  return r;
}
// split x = sg * |p| / 2^se: fl = floor(x), up = 1 if frac > 1/2, tie = 1 if frac == 1/2, bad = cannot
__device__ __forceinline__ void eps_split(double p, int ssign, int se, long long &fl, int &up, int &tie, int &bad) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(p);
  const int ex = (int)((bits >> 52) & 0x7ff);
  unsigned long long pm = bits & 0xfffffffffffffULL;
  fl = 0; up = 0; tie = 0; bad = 0;
  if (ex == 0x7ff) { bad = 1; return; }
  int pe;
  if (ex == 0) { pe = -1074; } else { pm |= (1ULL << 52); pe = ex - 1075; }
  if (pm == 0) return;
  const int sg = ((bits >> 63) ? -1 : 1) * ssign;
  const int sh = se - pe;
  if (sh <= 0) {
    if (-sh > 9) { bad = 1; return; }
    const long long I = (long long)(pm << (-sh));
    fl = sg > 0 ? I : -I;
    return;
  }
  if (sh > 63) { if (sg < 0) { fl = -1; up = 1; } return; }
  const unsigned long long I = (sh >= 64) ? 0ULL : (pm >> sh);
  const unsigned long long F = pm & ((1ULL << sh) - 1ULL), half = 1ULL << (sh - 1);
  if (sg > 0) {
    fl = (long long)I;
    if (F > half) up = 1; else if (F == half) tie = 1;
  } else if (F == 0) {
    fl = -(long long)I;
  } else {
    fl = -(long long)I - 1;
    const unsigned long long G = (1ULL << sh) - F;
    if (G > half) up = 1; else if (G == half) tie = 1;
  }
}
struct EpsShared {
  double s_sh;
  i64 pos_sh;
  int burst_sh, need_serial_sh, vmin_sh;
  long long Sv_sh;
  double prod[EPS_BURST_MAX];
  ParFn pf[EPS_T];
  long long ls[EPS_T];
};
// exact left-to-right accumulation of the terms [begin, n) onto s0 by one block; result in sh.s_sh
__device__ void eps_run(const double *a, const double *b, i64 begin, i64 n, double s0, EpsShared &sh) {
  double &s_sh = sh.s_sh;
  i64 &pos_sh = sh.pos_sh;
  int &burst_sh = sh.burst_sh, &need_serial_sh = sh.need_serial_sh, &vmin_sh = sh.vmin_sh;
  long long &Sv_sh = sh.Sv_sh;
  double *prod = sh.prod;
  ParFn *pf = sh.pf;
  long long *ls = sh.ls;
  const int t = threadIdx.x;
  __syncthreads();
  if (t == 0) { s_sh = s0; pos_sh = begin; burst_sh = 16; need_serial_sh = 0; }
  __syncthreads();
  while (true) {
    const i64 pos = pos_sh;
    if (pos >= n) break;
    const double s = s_sh;
    const unsigned long long sbits = (unsigned long long)__double_as_longlong(s);
    const int sex = (int)((sbits >> 52) & 0x7ff);
    const bool s_ok = (sex != 0 && sex != 0x7ff);           // normal, non-zero, finite
    if (sbits == 0ULL && !need_serial_sh) {
      // running sum is +0: zeros (of either sign) leave it +0 and the first non-zero term p gives
      // 0 + p = p exactly, so leading zeros are skipped with a parallel search
      if (t == 0) vmin_sh = EPS_T * EPS_E;
      __syncthreads();
      const i64 base0 = pos + (i64)t * EPS_E;
      int first = -1;
      for (int k = 0; k < EPS_E && first < 0; k++) {
        const i64 j = base0 + k;
        if (j >= n) { first = k; break; }
        const double p = b ? __dmul_rn(a[j], b[j]) : a[j];
        if (p != 0.0) first = k;              // NaN counts as non-zero
      }
      if (first >= 0) atomicMin(&vmin_sh, t * EPS_E + first);
      __syncthreads();
      if (t == 0) {
        const i64 j = pos + vmin_sh;
        if (vmin_sh == EPS_T * EPS_E || j >= n) pos_sh = (j < n) ? j : n;
        else {
          s_sh = __dadd_rn(0.0, b ? __dmul_rn(a[j], b[j]) : a[j]); pos_sh = j + 1;
          // a sum that starts from zero doubles every few terms at first: the next 256 terms go
          // through the plain chain (1 us) instead of one parallel pass per binade crossed
          need_serial_sh = 1; burst_sh = 256;
        }
      }
      __syncthreads();
      continue;
    }
    if (need_serial_sh || !s_ok) {
      // plain chain for a short burst: products staged by all threads, added by thread 0
      if (!s_ok && !need_serial_sh && t == 0) burst_sh = min(burst_sh * 2, EPS_BURST_MAX);   // subnormal / non-finite sums
      __syncthreads();
      const int cntb = (int)((n - pos < burst_sh) ? (n - pos) : burst_sh);
      for (int j = t; j < cntb; j += EPS_T) prod[j] = b ? __dmul_rn(a[pos + j], b[pos + j]) : a[pos + j];
      __syncthreads();
      if (t == 0) {
        double r = s;
        for (int j = 0; j < cntb; j++) r = __dadd_rn(r, prod[j]);
        s_sh = r; pos_sh = pos + cntb; need_serial_sh = 0;
      }
      __syncthreads();
      continue;
    }
    const int ssign = (sbits >> 63) ? -1 : 1;
    const int se = sex - 1075;
    const long long S0 = (long long)((sbits & 0xfffffffffffffULL) | (1ULL << 52));
    // ---- per-term split ----
    long long fl[EPS_E];
    int up[EPS_E], tie[EPS_E], bad[EPS_E];
    const i64 base = pos + (i64)t * EPS_E;
    ParFn mine; mine.reset = 0; mine.x = 0;
#pragma unroll
    for (int k = 0; k < EPS_E; k++) {
      const i64 j = base + k;
      if (j < n) {
        const double p = b ? __dmul_rn(a[j], b[j]) : a[j];
        eps_split(p, ssign, se, fl[k], up[k], tie[k], bad[k]);
      } else { fl[k] = 0; up[k] = 0; tie[k] = 0; bad[k] = 0; }
      ParFn e;
      e.reset = (unsigned char)tie[k];
      e.x = tie[k] ? 0 : (unsigned char)((fl[k] + up[k]) & 1);
      mine = par_compose(mine, e);
    }
    // ---- inclusive scan of the parity functions over threads: shuffles inside a warp, the 32
    // warp totals scanned by warp 0 ----
    {
      const int lane = t & 31, wid = t >> 5;
      unsigned fr = mine.reset, fx = mine.x;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const unsigned pr = __shfl_up_sync(0xffffffffu, fr, off), px = __shfl_up_sync(0xffffffffu, fx, off);
        if (lane >= off) { fx = fr ? fx : (px ^ fx); fr = fr | pr; }      // compose(prev, cur)
      }
      if (lane == 31) { pf[wid].reset = (unsigned char)fr; pf[wid].x = (unsigned char)fx; }
      __syncthreads();
      if (wid == 0) {
        unsigned wr = pf[lane].reset, wx = pf[lane].x;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const unsigned pr = __shfl_up_sync(0xffffffffu, wr, off), px = __shfl_up_sync(0xffffffffu, wx, off);
          if (lane >= off) { wx = wr ? wx : (px ^ wx); wr = wr | pr; }
        }
        pf[32 + lane].reset = (unsigned char)wr; pf[32 + lane].x = (unsigned char)wx;
      }
      __syncthreads();
      if (wid > 0) {                        // prepend everything before this warp
        const unsigned pr = pf[32 + wid - 1].reset, px = pf[32 + wid - 1].x;
        fx = fr ? fx : (px ^ fx); fr = fr | pr;
      }
      // exclusive value for this thread = inclusive of the previous thread
      const unsigned er = __shfl_up_sync(0xffffffffu, fr, 1), ex = __shfl_up_sync(0xffffffffu, fx, 1);
      unsigned qr, qx;
      if (lane > 0) { qr = er; qx = ex; }
      else if (wid > 0) { qr = pf[32 + wid - 1].reset; qx = pf[32 + wid - 1].x; }
      else { qr = 0; qx = 0; }
      mine.reset = (unsigned char)qr; mine.x = (unsigned char)qx;   // now: function of all earlier threads
    }
    int par = (int)(S0 & 1);
    par = mine.reset ? mine.x : (par ^ mine.x);
    // ---- increments with the tie decision, local sums ----
    long long c[EPS_E], lsum = 0;
#pragma unroll
    for (int k = 0; k < EPS_E; k++) {
      long long ck = fl[k] + up[k];
      if (tie[k]) { ck += ((par + fl[k]) & 1); par = 0; }
      else par ^= (int)(ck & 1);
      if (bad[k]) ck = 0;
      c[k] = ck; lsum += ck;
    }
    long long Sexcl;
    {
      const int lane = t & 31, wid = t >> 5;
      long long incl = lsum;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const long long pv = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += pv;
      }
      __syncthreads();                      // ls is free again
      if (lane == 31) ls[wid] = incl;
      __syncthreads();
      if (wid == 0) {
        long long wv = ls[lane];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const long long pv = __shfl_up_sync(0xffffffffu, wv, off);
          if (lane >= off) wv += pv;
        }
        ls[32 + lane] = wv;
      }
      __syncthreads();
      Sexcl = incl - lsum + (wid > 0 ? ls[32 + wid - 1] : 0);
    }
    long long S = S0 + Sexcl;
    // ---- first term that may leave the binade ----
    if (t == 0) vmin_sh = EPS_T * EPS_E;
    __syncthreads();
    long long Sbefore[EPS_E + 1];
    int myv = -1;
#pragma unroll
    for (int k = 0; k < EPS_E; k++) {
      Sbefore[k] = S;
      const i64 j = base + k;
      if (myv < 0 && (j >= n || bad[k] || S + fl[k] < (1LL << 52) || S + fl[k] + 1 >= (1LL << 53))) myv = k;
      S += c[k];
    }
    Sbefore[EPS_E] = S;
    if (myv >= 0) atomicMin(&vmin_sh, t * EPS_E + myv);
    __syncthreads();
    const int v = vmin_sh;                                   // accepted terms: [pos, pos+v)
    if (v == EPS_T * EPS_E) { if (t == EPS_T - 1) Sv_sh = Sbefore[EPS_E]; }
    else if (v / EPS_E == t) Sv_sh = Sbefore[v % EPS_E];
    __syncthreads();
    if (t == 0) {
      const unsigned long long m = (unsigned long long)Sv_sh;
      const unsigned long long nb = ((unsigned long long)(ssign < 0) << 63) | ((unsigned long long)(se + 1075) << 52) |
                                    (m & 0xfffffffffffffULL);
      s_sh = __longlong_as_double((long long)nb);
      pos_sh = pos + v;
      const bool hit = (pos + v < n) && (v < EPS_T * EPS_E);
      need_serial_sh = hit ? 1 : 0;
      if (hit) burst_sh = (v < 256) ? min(burst_sh * 2, EPS_BURST_MAX) : max(16, burst_sh / 2);
    }
    __syncthreads();
  }
}
__global__ void __launch_bounds__(EPS_T) k_eps_dot(const double *a, const double *b, i64 n, double *out) {
  __shared__ EpsShared sh;
  eps_run(a, b, 0, n, 0.0, sh);
  if (threadIdx.x == 0) *out = sh.s_sh;
}

// ---- many-block version for long vectors ----
// The vector is cut into SEGMENTS of at most 8192 terms.  Every segment is reduced speculatively
// by its own block, assuming the running sum enters it with the sign and exponent of an
// approximate (order-free) prefix sum and stays in that binade.  A segment record holds the
// integer total for an even incoming mantissa, what changes for an odd one (only the first tie of
// the segment sees the incoming parity: delta), and the extreme values the running mantissa takes
// before/after that tie, so that one thread can afterwards walk the segments in order, verify the
// assumption for the ACTUAL incoming sum (same sign, same exponent, mantissa stays inside
// [2^52, 2^53)) and either accept the segment in O(1) or hand it to eps_run.
//
// A sum of n similar terms crosses a binade about log2(n) times, and a segment with a crossing
// inside can never be accepted.  So the plan is made from approximate prefix sums at two
// granularities (k_eps_plan): a chunk of 8192 terms whose approximate running sum keeps its
// exponent is one segment; a chunk where it changes is cut at its 256-term sub-chunks into runs of
// constant exponent (speculative segments, each with its own hypothesis) and the sub-chunks where
// the exponent changes, which the walker adds the plain way (256 dependent additions from shared
// memory).  Whatever the plan says, every accept is verified against the actual running sum and
// every rejected segment is redone exactly, so a wrong guess costs time, never a bit.
struct EpsChunk {
  long long total0, lo_pre, hi_pre, lo_post, hi_post;
  i64 begin, end;
  int delta, bad, gsign, gsex;
  int kind, pad_;                   // 0 speculative, 1 plain chain, 2 eps_run (no hypothesis)
};
#define EPS_C (EPS_T * EPS_E)
#define EPS_SUB 256                  // terms per sub-chunk
#define EPS_NSUB (EPS_C / EPS_SUB)   // 32 sub-chunks per chunk
// sub[c*32 + w]: order-free sum of the terms [c*8192 + w*256, +256) (one warp each)
__global__ void __launch_bounds__(256) k_eps_chunk_sums(const double *a, const double *b, i64 n, double *sub) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const i64 cbase = (i64)blockIdx.x * EPS_C;
  for (int w = wid; w < EPS_NSUB; w += 8) {
    const i64 base = cbase + (i64)w * EPS_SUB;
    double x = 0.0;
#pragma unroll
    for (int k = 0; k < EPS_SUB / 32; k++) { const i64 j = base + k * 32 + lane; if (j < n) x += b ? a[j] * b[j] : a[j]; }
    for (int off = 16; off >= 1; off >>= 1) x += __shfl_down_sync(0xffffffffu, x, off);
    if (lane == 0) sub[(i64)blockIdx.x * EPS_NSUB + w] = x;
  }
}
__device__ __forceinline__ int eps_expo(double x) { return (int)(((unsigned long long)__double_as_longlong(x) >> 52) & 0x7ff); }
__device__ __forceinline__ int eps_sgn(double x) { return ((unsigned long long)__double_as_longlong(x) >> 63) ? -1 : 1; }
__device__ __forceinline__ bool eps_same_binade(double x, double y) {
  const int ex = eps_expo(x);
  return ex == eps_expo(y) && eps_sgn(x) == eps_sgn(y) && ex != 0 && ex != 0x7ff;
}
// The plan, by one block: warp w looks at chunk c = w, w+32, ... -- lane l sees the approximate
// running sum before and after sub-chunk l (a warp scan of the 32 sub-sums on top of the chunk's
// approximate prefix) -- and the chunk becomes
//   * one speculative segment, if the exponent never changes,
//   * runs of sub-chunks with one exponent (speculative) and single sub-chunks where it changes
//     (plain chain), if it changes in at most EPS_MAXCUT sub-chunks,
//   * one eps_run segment otherwise (sums around zero, leading zeros: eps_run adapts by itself).
// The chunk prefixes and the segment offsets are two short serial scans by thread 0 over shared
// memory; chunks are handled in tiles of EPS_PLAN_TILE.
#define EPS_MAXCUT 8
#define EPS_SEG_PER_CHUNK (2 * EPS_MAXCUT + 1)
#define EPS_PLAN_TILE 1024
__device__ __forceinline__ int eps_plan_chunk(const double *sub, int c, i64 n, double pre, EpsChunk *out) {
  // returns the number of segments of chunk c; writes them to out[0..) unless out == nullptr
  const int lane = threadIdx.x & 31;
  const i64 cb = (i64)c * EPS_C, ce = (cb + EPS_C < n) ? cb + EPS_C : n;
  double incl = sub[(i64)c * EPS_NSUB + lane];
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) { const double pv = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += pv; }
  const double after = pre + incl;                                   // approximate running sum after my sub-chunk
  double before = __shfl_up_sync(0xffffffffu, after, 1);
  if (lane == 0) before = pre;
  const i64 sb = cb + (i64)lane * EPS_SUB;
  const bool live = sb < ce;                                         // sub-chunk holds terms
  const bool steady = !live || eps_same_binade(before, after);       // exponent kept across my sub-chunk
  const unsigned unsteady = __ballot_sync(0xffffffffu, !steady);
  if (unsteady == 0u) {
    if (out && lane == 0) { EpsChunk &R = out[0]; R.begin = cb; R.end = ce; R.kind = 0; R.bad = 0; R.gsign = eps_sgn(pre); R.gsex = eps_expo(pre); }
    return 1;
  }
  if (__popc(unsteady) > EPS_MAXCUT) {
    if (out && lane == 0) { EpsChunk &R = out[0]; R.begin = cb; R.end = ce; R.kind = 2; R.bad = 1; R.gsign = 1; R.gsex = 0; }
    return 1;
  }
  // a steady sub-chunk starts a run when it is the first of the chunk or follows an unsteady one
  // (consecutive steady sub-chunks share their exponent: after of one is before of the next)
  const bool prev_unsteady = lane > 0 && ((unsteady >> (lane - 1)) & 1u);
  const bool head = live && (!steady || lane == 0 || prev_unsteady);
  const unsigned heads = __ballot_sync(0xffffffffu, head);
  if (out && head) {
    const int slot = __popc(heads & ((1u << lane) - 1u));
    const unsigned later = (lane == 31) ? 0u : (heads >> (lane + 1));
    const int nextw = later ? lane + 1 + (__ffs(later) - 1) : EPS_NSUB;
    EpsChunk &R = out[slot];
    i64 e = cb + (i64)nextw * EPS_SUB;
    if (e > ce) e = ce;
    R.begin = sb; R.end = e;
    R.kind = steady ? 0 : 1;
    R.bad = steady ? 0 : 1;
    R.gsign = eps_sgn(before); R.gsex = eps_expo(before);
  }
  return __popc(heads);
}
__global__ void __launch_bounds__(1024) k_eps_plan(const double *sub, int nchunks, i64 n, EpsChunk *rec, int *nseg_out) {
  __shared__ double pre_sh[EPS_PLAN_TILE];
  __shared__ int off_sh[EPS_PLAN_TILE];
  __shared__ double carry_sh;
  __shared__ int ns_sh;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  if (t == 0) { carry_sh = 0.0; ns_sh = 0; }
  for (int c0 = 0; c0 < nchunks; c0 += EPS_PLAN_TILE) {
    const int nt = min(EPS_PLAN_TILE, nchunks - c0);
    __syncthreads();
    for (int q = wid; q < nt; q += 32) {                               // approximate chunk totals
      double x = sub[(i64)(c0 + q) * EPS_NSUB + lane];
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) x += __shfl_down_sync(0xffffffffu, x, off);
      if (lane == 0) pre_sh[q] = x;
    }
    __syncthreads();
    if (t == 0) {                                                      // exclusive prefix: running sum before each chunk
      double p = carry_sh;
      for (int q = 0; q < nt; q++) { const double x = pre_sh[q]; pre_sh[q] = p; p += x; }
      carry_sh = p;
    }
    __syncthreads();
    for (int q = wid; q < nt; q += 32) {
      const int k = eps_plan_chunk(sub, c0 + q, n, pre_sh[q], nullptr);
      if (lane == 0) off_sh[q] = k;
    }
    __syncthreads();
    if (t == 0) {
      int o = ns_sh;
      for (int q = 0; q < nt; q++) { const int k = off_sh[q]; off_sh[q] = o; o += k; }
      ns_sh = o;
    }
    __syncthreads();
    for (int q = wid; q < nt; q += 32) eps_plan_chunk(sub, c0 + q, n, pre_sh[q], rec + off_sh[q]);
  }
  __syncthreads();
  if (t == 0) *nseg_out = ns_sh;
}
__global__ void __launch_bounds__(EPS_T) k_eps_chunk_stats(const double *a, const double *b, EpsChunk *rec, const int *nseg) {
  __shared__ ParFn pf[64 + 32];
  __shared__ long long ls[64 + 32];
  __shared__ int first_tie, bad_sh, d0_sh;
  __shared__ long long lo_pre, hi_pre, lo_post, hi_post;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int total = *nseg;
  for (int seg = blockIdx.x; seg < total; seg += gridDim.x) {
  __syncthreads();
  EpsChunk &R = rec[seg];
  if (R.kind != 0) continue;
  const i64 n = R.end;
  const int sex = R.gsex, ssign = R.gsign;
  // an unusable hypothesis is a property of the record, the same for every thread (bad_sh is
  // written by whichever thread meets a bad term later on and is only read after the barriers)
  if (sex == 0 || sex == 0x7ff) { if (t == 0) R.bad = 1; continue; }
  if (t == 0) {
    first_tie = EPS_C; bad_sh = 0; d0_sh = 0;
    lo_pre = lo_post = (1LL << 62); hi_pre = hi_post = -(1LL << 62);
  }
  __syncthreads();
  const int se = sex - 1075;
  const i64 base = R.begin + (i64)t * EPS_E;
  long long fl[EPS_E];
  int up[EPS_E], tie[EPS_E];
  int anybad = 0;
  unsigned fr = 0, fx = 0;
#pragma unroll
  for (int k = 0; k < EPS_E; k++) {
    const i64 j = base + k;
    int bd = 0;
    if (j < n) {
      const double p = b ? __dmul_rn(a[j], b[j]) : a[j];
      eps_split(p, ssign, se, fl[k], up[k], tie[k], bd);
    } else { fl[k] = 0; up[k] = 0; tie[k] = 0; }
    anybad |= bd;
    const unsigned er = (unsigned)tie[k], ex = tie[k] ? 0u : (unsigned)((fl[k] + up[k]) & 1);
    fx = er ? ex : (fx ^ ex); fr = fr | er;
  }
  if (anybad) bad_sh = 1;
  // scan of the parity functions (incoming parity 0)
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const unsigned pr = __shfl_up_sync(0xffffffffu, fr, off), px = __shfl_up_sync(0xffffffffu, fx, off);
    if (lane >= off) { fx = fr ? fx : (px ^ fx); fr = fr | pr; }
  }
  if (lane == 31) { pf[wid].reset = (unsigned char)fr; pf[wid].x = (unsigned char)fx; }
  __syncthreads();
  if (wid == 0) {
    unsigned wr = pf[lane].reset, wx = pf[lane].x;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned pr = __shfl_up_sync(0xffffffffu, wr, off), px = __shfl_up_sync(0xffffffffu, wx, off);
      if (lane >= off) { wx = wr ? wx : (px ^ wx); wr = wr | pr; }
    }
    pf[32 + lane].reset = (unsigned char)wr; pf[32 + lane].x = (unsigned char)wx;
  }
  __syncthreads();
  if (wid > 0) { const unsigned pr = pf[32 + wid - 1].reset, px = pf[32 + wid - 1].x; fx = fr ? fx : (px ^ fx); fr = fr | pr; }
  const unsigned er = __shfl_up_sync(0xffffffffu, fr, 1), ex = __shfl_up_sync(0xffffffffu, fx, 1);
  unsigned qr, qx;
  if (lane > 0) { qr = er; qx = ex; }
  else if (wid > 0) { qr = pf[32 + wid - 1].reset; qx = pf[32 + wid - 1].x; }
  else { qr = 0; qx = 0; }
  int par = qr ? (int)qx : (int)(0 ^ qx);          // parity before this thread's terms if the incoming one is even
  // increments, first tie
  long long c[EPS_E], lsum = 0;
  int myfirst = -1, myd0 = 0;
#pragma unroll
  for (int k = 0; k < EPS_E; k++) {
    long long ck = fl[k] + up[k];
    if (tie[k]) {
      const int d = (int)((par + fl[k]) & 1);
      if (myfirst < 0) { myfirst = k; myd0 = d; }
      ck += d; par = 0;
    } else par ^= (int)(ck & 1);
    c[k] = ck; lsum += ck;
  }
  if (myfirst >= 0) atomicMin(&first_tie, t * EPS_E + myfirst);
  // prefix sums of the increments
  long long incl = lsum;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) { const long long pv = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += pv; }
  if (lane == 31) ls[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    long long wv = ls[lane];
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { const long long pv = __shfl_up_sync(0xffffffffu, wv, off); if (lane >= off) wv += pv; }
    ls[32 + lane] = wv;
  }
  __syncthreads();
  long long P = incl - lsum + (wid > 0 ? ls[32 + wid - 1] : 0);      // prefix before this thread's terms
  const int ft = first_tie;
  if (ft / EPS_E == t && ft < EPS_C) d0_sh = myd0;
  long long mlo_pre = (1LL << 62), mhi_pre = -(1LL << 62), mlo_post = (1LL << 62), mhi_post = -(1LL << 62);
#pragma unroll
  for (int k = 0; k < EPS_E; k++) {
    const i64 j = base + k;
    if (j < n) {
      const long long v = P + fl[k];
      if (t * EPS_E + k <= ft) { mlo_pre = min(mlo_pre, v); mhi_pre = max(mhi_pre, v + 1); }
      else { mlo_post = min(mlo_post, v); mhi_post = max(mhi_post, v + 1); }
    }
    P += c[k];
  }
  for (int off = 16; off >= 1; off >>= 1) {
    mlo_pre = min(mlo_pre, __shfl_down_sync(0xffffffffu, mlo_pre, off)); mhi_pre = max(mhi_pre, __shfl_down_sync(0xffffffffu, mhi_pre, off));
    mlo_post = min(mlo_post, __shfl_down_sync(0xffffffffu, mlo_post, off)); mhi_post = max(mhi_post, __shfl_down_sync(0xffffffffu, mhi_post, off));
  }
  if (lane == 0) { atomicMin(&lo_pre, mlo_pre); atomicMax(&hi_pre, mhi_pre); atomicMin(&lo_post, mlo_post); atomicMax(&hi_post, mhi_post); }
  __syncthreads();
  if (t == EPS_T - 1) R.total0 = P;                                   // total of the segment (incoming parity even)
  if (t == 0) {
    R.lo_pre = lo_pre; R.hi_pre = hi_pre; R.lo_post = lo_post; R.hi_post = hi_post;
    R.delta = (ft < EPS_C) ? (1 - 2 * d0_sh) : 0;
    R.bad = bad_sh;
  }
  }
}
#define EPS_SREC 192                 // segment records cached in shared memory by the walker
__global__ void __launch_bounds__(EPS_T) k_eps_combine(const double *a, const double *b, const EpsChunk *rec,
                                                       const int *nseg_p, double *out, int *nfallback) {
  __shared__ EpsShared sh;
  __shared__ int next_sh;
  __shared__ double cur_sh;
  __shared__ EpsChunk srec[EPS_SREC];
  const int t = threadIdx.x;
  const int nseg = *nseg_p;
  if (t == 0) { cur_sh = 0.0; next_sh = 0; }
  for (int q = t; q < nseg && q < EPS_SREC; q += EPS_T) srec[q] = rec[q];
  __syncthreads();
  int fb = 0;
  while (true) {
    if (t == 0) {
      // accept as many consecutive segments as the assumptions allow
      double s = cur_sh;
      int c = next_sh;
      for (; c < nseg; c++) {
        const EpsChunk &R = (c < EPS_SREC) ? srec[c] : rec[c];
        const unsigned long long bits = (unsigned long long)__double_as_longlong(s);
        const int sex = (int)((bits >> 52) & 0x7ff), sg = (bits >> 63) ? -1 : 1;
        if (R.kind != 0 || R.bad || sex != R.gsex || sg != R.gsign || sex == 0 || sex == 0x7ff) break;
        const long long S = (long long)((bits & 0xfffffffffffffULL) | (1ULL << 52));
        const int odd = (int)(S & 1);
        const long long d = odd ? R.delta : 0;
        const long long lo = min(R.lo_pre, R.lo_post + d), hi = max(R.hi_pre, R.hi_post + d);
        if (S + lo < (1LL << 52) || S + hi >= (1LL << 53)) break;
        const long long So = S + R.total0 + d;
        const unsigned long long nb = (bits & 0xfff0000000000000ULL) | ((unsigned long long)So & 0xfffffffffffffULL);
        s = __longlong_as_double((long long)nb);
      }
      cur_sh = s; next_sh = c;
    }
    __syncthreads();
    const int c = next_sh;
    if (c >= nseg) break;
    const EpsChunk &R = (c < EPS_SREC) ? srec[c] : rec[c];
    const i64 b0 = R.begin, b1 = R.end;
    const int kind = R.kind;
    __syncthreads();
    if (kind == 1 && b1 - b0 <= EPS_BURST_MAX) {
      // the plan expects a binade crossing in these few terms: plain chain, products staged by all
      const int cnt = (int)(b1 - b0);
      for (int j = t; j < cnt; j += EPS_T) sh.prod[j] = b ? __dmul_rn(a[b0 + j], b[b0 + j]) : a[b0 + j];
      __syncthreads();
      if (t == 0) {
        double r = cur_sh;
        for (int j = 0; j < cnt; j++) r = __dadd_rn(r, sh.prod[j]);
        cur_sh = r; next_sh = c + 1;
      }
    } else {
      eps_run(a, b, b0, b1, cur_sh, sh);                              // the whole block redoes this segment exactly
      fb++;
      __syncthreads();
      if (t == 0) { cur_sh = sh.s_sh; next_sh = c + 1; }
    }
    __syncthreads();
  }
  if (t == 0) { *out = cur_sh; if (nfallback) *nfallback = fb; }
}
static int g_eps = -1;

void seq_dot_dev(double *outp, const double *a, const double *b, i64 n) {
  if (n <= 0) { dev_memset(outp, 0, sizeof(double)); return; }
  if (g_eps < 0) { const char *e = getenv("AMGB_SEQDOT"); g_eps = (e && !strcmp(e, "chain")) ? 0 : (e && !strcmp(e, "block")) ? 1 : 2; }
  if (g_eps == 2 && (n >= 4 * (i64)EPS_C || (test_force('b') && n > EPS_C))) {
    const int nch = (int)((n + EPS_C - 1) / EPS_C);
    const int cap = EPS_SEG_PER_CHUNK * nch;
    const int grid = 2 * nch + 64 < cap ? 2 * nch + 64 : cap;       // a plan rarely holds more segments than this
    Buf<double> sub((i64)nch * EPS_NSUB);
    Buf<EpsChunk> rec(cap);
    Buf<int> nseg(1);
    k_eps_chunk_sums<<<nch, 256, 0, g_ctx.stream>>>(a, b, n, sub.p);
    k_eps_plan<<<1, 1024, 0, g_ctx.stream>>>(sub.p, nch, n, rec.p, nseg.p);
    k_eps_chunk_stats<<<grid, EPS_T, 0, g_ctx.stream>>>(a, b, rec.p, nseg.p);
    k_eps_combine<<<1, EPS_T, 0, g_ctx.stream>>>(a, b, rec.p, nseg.p, outp, nullptr);
    g_ctx.launches += 3;
  } else if (g_eps) k_eps_dot<<<1, EPS_T, 0, g_ctx.stream>>>(a, b, n, outp);
  else k_seq_dot<<<1, 1024, 0, g_ctx.stream>>>(a, b, n, outp);
  g_ctx.launches++; post_launch(__func__);
}
double seq_dot(const double *a, const double *b, i64 n) {
  if (n <= 0) return 0.0;
  StageTimer st_("prim.seq_dot");
  Buf<double> out(1);
  seq_dot_dev(out.p, a, b, n);
  return out.get(0);
}
double seq_sum(const double *v, i64 n) { return seq_dot(v, nullptr, n); }

// ---- max with first index ----
struct MaxIdx { double v; i64 i; };
__device__ __forceinline__ MaxIdx better(MaxIdx a, MaxIdx b) {
  if (b.i < 0) return a;
  if (a.i < 0) return b;
  if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}
__global__ void __launch_bounds__(256) k_max_first(const double *v, const i64 *idx_in, i64 n, double *ov, i64 *oi) {
  __shared__ double sv[8];
  __shared__ i64 si[8];
  MaxIdx m{0.0, -1};
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
    MaxIdx c{v[i], idx_in ? idx_in[i] : i};
    m = better(m, c);
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    MaxIdx o{__shfl_down_sync(0xffffffffu, m.v, off), __shfl_down_sync(0xffffffffu, m.i, off)};
    m = better(m, o);
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { sv[w] = m.v; si[w] = m.i; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; k++) m = better(m, MaxIdx{sv[k], si[k]});
    ov[blockIdx.x] = m.v; oi[blockIdx.x] = m.i;
  }
}
void max_first(const double *v, i64 n, double *val, i64 *idx) {
  if (n <= 0) throw Error(-3, "max_first on an empty vector");
  StageTimer st_("prim.max_first");
  i64 nb = (n + 255) / 256;
  if (nb > 1024) nb = 1024;
  Buf<double> pv(nb), fv(1);
  Buf<i64> pi(nb), fi(1);
  k_max_first<<<(unsigned)nb, 256, 0, g_ctx.stream>>>(v, nullptr, n, pv.p, pi.p);
  k_max_first<<<1, 256, 0, g_ctx.stream>>>(pv.p, pi.p, nb, fv.p, fi.p);
  g_ctx.launches += 2; post_launch(__func__);
  struct { double v; i64 i; } out;
  d2h(&out.v, fv.p, sizeof(double));
  d2h(&out.i, fi.p, sizeof(i64));
  *val = out.v;
  if (idx) *idx = out.i;
}

__global__ void __launch_bounds__(256) k_count_nonzero(const double *v, i64 n, unsigned long long *out) {
  unsigned long long c = 0;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
    c += (v[i] != 0.0);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) c += __shfl_down_sync(0xffffffffu, c, off);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}
__global__ void __launch_bounds__(256) k_max_first2_final(const double *pv, const i64 *pi, const double *qv, i64 nb,
                                                          double *out) {
  __shared__ double sv[8], sq[8];
  __shared__ i64 si[8];
  MaxIdx m{0.0, -1};
  double q = -DBL_MAX;
  for (i64 i = threadIdx.x; i < nb; i += blockDim.x) { m = better(m, MaxIdx{pv[i], pi[i]}); q = fmax(q, qv[i]); }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    MaxIdx o{__shfl_down_sync(0xffffffffu, m.v, off), __shfl_down_sync(0xffffffffu, m.i, off)};
    m = better(m, o);
    q = fmax(q, __shfl_down_sync(0xffffffffu, q, off));
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { sv[w] = m.v; si[w] = m.i; sq[w] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; k++) { m = better(m, MaxIdx{sv[k], si[k]}); q = fmax(q, sq[k]); }
    out[0] = m.v; out[1] = (double)m.i; out[2] = q;
  }
}
__global__ void __launch_bounds__(256) k_max_first2(const double *a, const double *b, i64 n, double *ov, i64 *oi, double *oq) {
  __shared__ double sv[8], sq[8];
  __shared__ i64 si[8];
  MaxIdx m{0.0, -1};
  double q = -DBL_MAX;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
    m = better(m, MaxIdx{a[i], i});
    q = fmax(q, b[i]);
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    MaxIdx o{__shfl_down_sync(0xffffffffu, m.v, off), __shfl_down_sync(0xffffffffu, m.i, off)};
    m = better(m, o);
    q = fmax(q, __shfl_down_sync(0xffffffffu, q, off));
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { sv[w] = m.v; si[w] = m.i; sq[w] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; k++) { m = better(m, MaxIdx{sv[k], si[k]}); q = fmax(q, sq[k]); }
    ov[blockIdx.x] = m.v; oi[blockIdx.x] = m.i; oq[blockIdx.x] = q;
  }
}
void max_first2(const double *a, const double *b, i64 n, double *amax, i64 *aidx, double *bmax) {
  if (n <= 0) throw Error(-3, "max_first2 on an empty vector");
  i64 nb = (n + 255) / 256;
  if (nb > 592) nb = 592;
  Buf<double> pv(nb), qv(nb), out(3);
  Buf<i64> pi(nb);
  k_max_first2<<<(unsigned)nb, 256, 0, g_ctx.stream>>>(a, b, n, pv.p, pi.p, qv.p);
  k_max_first2_final<<<1, 256, 0, g_ctx.stream>>>(pv.p, pi.p, qv.p, nb, out.p);
  g_ctx.launches += 2; post_launch(__func__);
  double h[3];
  d2h(h, out.p, sizeof h);
  *amax = h[0]; *aidx = (i64)h[1]; *bmax = h[2];
}

i64 count_nonzero(const double *v, i64 n) {
  if (n <= 0) return 0;
  Buf<unsigned long long> c(1);
  c.zero();
  i64 nb = (n + 255) / 256;
  if (nb > 1184) nb = 1184;
  k_count_nonzero<<<(unsigned)nb, 256, 0, g_ctx.stream>>>(v, n, c.p);
  g_ctx.launches++; post_launch(__func__);
  return (i64)c.get(0);
}

#else
// =======================================================================================
// host emulation build (tests only)
// =======================================================================================
void ctx_init(int) {}
void post_launch(const char *) {}
void *dev_alloc(size_t bytes) { void *p = malloc(bytes ? bytes : 8); if (!p) throw Error(-2, "out of memory"); return p; }
void dev_free(void *p) { free(p); }
void dev_release_cache() {}
size_t dev_peak_bytes() { return 0; }
void dev_memset(void *p, int v, size_t bytes) { memset(p, v, bytes); }
void h2d(void *dst, const void *src, size_t bytes) { memcpy(dst, src, bytes); }
void d2h(void *dst, const void *src, size_t bytes) { memcpy(dst, src, bytes); g_ctx.syncs++; }
void d2h_async(void *dst, const void *src, size_t bytes) { memcpy(dst, src, bytes); }
void d2d(void *dst, const void *src, size_t bytes) { memcpy(dst, src, bytes); }
void stream_sync() { g_ctx.syncs++; }

i64 exclusive_scan(const int *in, int *out, i64 n) {
  int s = 0;
  for (i64 i = 0; i < n; i++) { int v = in[i]; out[i] = s; s += v; }
  out[n] = s;
  return s;
}
void exclusive_scan_dev(const int *in, int *out, i64 n) { exclusive_scan(in, out, n); }
i64 exclusive_scan64(const i64 *in, i64 *out, i64 n) {
  i64 s = 0;
  for (i64 i = 0; i < n; i++) { i64 v = in[i]; out[i] = s; s += v; }
  out[n] = s;
  return s;
}
static double chunk_tree_host(const double *v, i64 m, const double *b) {
  double s[256];
  for (int t = 0; t < 256; t++) {
    double x = 0.0;
    for (int k = 0; k < 4; k++) {
      i64 i = t + 256 * k;
      double p = (i < m) ? (b ? v[i] * b[i] : v[i]) : 0.0;
      x = (k == 0) ? p : x + p;
    }
    s[t] = x;
  }
  double w[8];
  for (int k = 0; k < 8; k++) {
    double *x = s + 32 * k;
    for (int off = 16; off >= 1; off >>= 1) for (int l = 0; l < off; l++) x[l] = x[l] + x[l + off];
    w[k] = x[0];
  }
  for (int off = 4; off >= 1; off >>= 1) for (int l = 0; l < off; l++) w[l] = w[l] + w[l + off];
  return w[0];
}
static double tree_host(const double *v, const double *b, i64 n) {
  if (n <= 0) return 0.0;
  i64 nc = (n + 1023) / 1024;
  std::vector<double> part((size_t)nc);
  for (i64 c = 0; c < nc; c++) {
    i64 m = n - c * 1024; if (m > 1024) m = 1024;
    part[(size_t)c] = chunk_tree_host(v + c * 1024, m, b ? b + c * 1024 : nullptr);
  }
  if (nc == 1) return part[0];
  return tree_host(part.data(), nullptr, nc);
}
double tree_sum(const double *v, i64 n) { return tree_host(v, nullptr, n); }
double tree_dot(const double *a, const double *b, i64 n) { return tree_host(a, b, n); }
double seq_sum(const double *v, i64 n) { double r = 0; for (i64 i = 0; i < n; i++) r += v[i]; return r; }
double seq_dot(const double *a, const double *b, i64 n) { double r = 0; for (i64 i = 0; i < n; i++) r += b ? a[i] * b[i] : a[i]; return r; }
void seq_dot_dev(double *out, const double *a, const double *b, i64 n) { *out = seq_dot(a, b, n); }
void tree_dot_dev(double *out, const double *a, const double *b, i64 n) { *out = tree_host(a, b, n); }
void max_first(const double *v, i64 n, double *val, i64 *idx) {
  if (n <= 0) throw Error(-3, "max_first on an empty vector");
  double m = v[0]; i64 k = 0;
  for (i64 i = 1; i < n; i++) if (v[i] > m) { m = v[i]; k = i; }
  *val = m; if (idx) *idx = k;
}
i64 count_nonzero(const double *v, i64 n) { i64 c = 0; for (i64 i = 0; i < n; i++) c += (v[i] != 0.0); return c; }
void max_first2(const double *a, const double *b, i64 n, double *amax, i64 *aidx, double *bmax) {
  max_first(a, n, amax, aidx);
  max_first(b, n, bmax, nullptr);
}
#endif

bool test_small_bins() { const char *e = getenv("AMGB_TEST_SMALL_BINS"); return e && *e == '1'; }
bool test_force(char what) { const char *e = getenv("AMGB_TEST_FORCE"); return e && strchr(e, what) != nullptr; }

// ---- sub-stage profile ----
static std::map<std::string, std::pair<double, long>> g_stage;
static int g_stage_on = -1;
static double wall_now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
StageTimer::StageTimer(const char *n) : name(n), t0(0), on(false) {
  if (g_stage_on < 0) { const char *e = getenv("AMGB_STAGE_LOG"); g_stage_on = (e && *e && *e != '0') ? 1 : 0; }
  on = g_stage_on > 0;
  if (on) { stream_sync(); t0 = wall_now(); }
}
StageTimer::~StageTimer() {
  if (!on) return;
  try { stream_sync(); } catch (...) {}
  auto &r = g_stage[name];
  r.first += wall_now() - t0; r.second++;
}
static std::map<std::string, long> g_counts;
void stage_count(const char *name, long n) { if (g_stage_on > 0) g_counts[name] += n; }
void stage_report() {
  if (g_stage_on <= 0) return;
  std::vector<std::pair<double, std::string>> v;
  for (auto &kv : g_stage) v.push_back({kv.second.first, kv.first});
  std::sort(v.begin(), v.end());
  fprintf(stderr, "---- stage profile (inclusive, synchronised) rank %d ----\n", comm_rank());
  for (auto it = v.rbegin(); it != v.rend(); ++it)
    fprintf(stderr, "%-28s %10.3f ms  calls %ld\n", it->second.c_str(), it->first * 1e3, g_stage[it->second].second);
  for (auto &kv : g_counts) fprintf(stderr, "count %-24s %ld\n", kv.first.c_str(), kv.second);
  g_counts.clear();
  g_stage.clear();
}

// ---- trace ----
static uint64_t fnv1a(const void *p, size_t n) {
  const unsigned char *b = (const unsigned char *)p;
  uint64_t h = 1469598103934665603ULL;
  for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ULL; }
  return h;
}
void trace_dev(const char *tag, const void *dptr, size_t bytes) {
  Context &c = ctx();
  if (!c.trace_on) return;
  std::vector<unsigned char> h(bytes ? bytes : 1);
  if (bytes) d2h(h.data(), dptr, bytes);
  c.trace.push_back(TraceRec{c.trace_prefix + tag, fnv1a(h.data(), bytes), (i64)bytes});
}

}  // namespace amgb

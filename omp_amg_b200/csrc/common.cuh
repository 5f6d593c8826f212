// common.cuh -- device memory, launch helpers, scans and reductions shared by the kernels.
//
// Two build modes:
//   default      real CUDA (sm_100a).  This is the product.
//   AMGB_EMU     the same sources compiled for the host: parallel_for runs its body in a
//                loop, atomics are plain operations.  A development aid used by the CPU-only
//                test-suite to exercise the host orchestration; it is built into
//                tests/_emu/ and is never loaded by the omp_amg_b200 package.
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cfloat>
#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>
#include <utility>

#ifndef AMGB_EMU
#include <cuda_runtime.h>
#define HD __host__ __device__
#define DEV __device__
#else
#define HD
#define DEV
#endif

namespace amgb {

typedef long long i64;

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

#ifndef AMGB_EMU
#define CUDA_CHECK(x)                                                                         \
  do {                                                                                        \
    cudaError_t e_ = (x);                                                                     \
    if (e_ != cudaSuccess)                                                                    \
      throw ::amgb::Error(-100, std::string("CUDA error ") + cudaGetErrorString(e_) + " at " + \
                                    __FILE__ + ":" + std::to_string(__LINE__));               \
  } while (0)
#endif

// ---------------------------------------------------------------------------------------
// Context: one stream, launch counter, trace
// ---------------------------------------------------------------------------------------
struct TraceRec { std::string tag; uint64_t hash; i64 bytes; };

struct Context {
#ifndef AMGB_EMU
  cudaStream_t stream = nullptr;
#endif
  i64 launches = 0;          // kernels launched by this library (bench.py "gpu_launches")
  i64 syncs = 0;             // host<->device synchronisations
  // order of the vector-length reductions (dot products, 2-norms):
  //   1  left to right, as the reference's vv_dot/array_op (amg_setup.c:3193,3309): one thread
  //      carries the sum, so results are bit-identical to the reference
  //   0  fixed 1024-chunk tree: fast, deterministic, but rounds differently from the reference
  int reduce_seq = 1;
  bool trace_on = false;
  std::string trace_prefix;
  std::vector<TraceRec> trace;
  int sm_count = 148;
};
Context &ctx();
void ctx_init(int device);
// checks the launch; with AMGB_DEBUG_SYNC=1 also synchronises so that faults are attributed
void post_launch(const char *what);

// ---------------------------------------------------------------------------------------
// device buffers (stream-ordered pool allocations)
// ---------------------------------------------------------------------------------------
void *dev_alloc(size_t bytes);
void dev_free(void *p);
void dev_release_cache();      // hand cached blocks back to the driver
size_t dev_peak_bytes();       // high-water mark of live device memory
void dev_memset(void *p, int v, size_t bytes);
void h2d(void *dst, const void *src, size_t bytes);
void d2h(void *dst, const void *src, size_t bytes);   // synchronises the stream
void d2h_async(void *dst, const void *src, size_t bytes);   // no synchronisation: call stream_sync() before reading
void d2d(void *dst, const void *src, size_t bytes);
void stream_sync();

template <class T>
struct Buf {
  T *p = nullptr;
  i64 n = 0;
  Buf() {}
  explicit Buf(i64 n_) { alloc(n_); }
  Buf(const Buf &) = delete;
  Buf &operator=(const Buf &) = delete;
  Buf(Buf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  Buf &operator=(Buf &&o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~Buf() { release(); }
  void alloc(i64 n_) {
    release();
    n = n_;
    p = (T *)dev_alloc(sizeof(T) * (size_t)(n_ > 0 ? n_ : 1));
  }
  void release() { if (p) { dev_free(p); p = nullptr; } n = 0; }
  void zero() { dev_memset(p, 0, sizeof(T) * (size_t)n); }
  void upload(const T *h, i64 cnt) { h2d(p, h, sizeof(T) * (size_t)cnt); }
  std::vector<T> download() const {
    std::vector<T> v((size_t)n);
    if (n) d2h(v.data(), p, sizeof(T) * (size_t)n);
    return v;
  }
  T get(i64 i) const { T v; d2h(&v, p + i, sizeof(T)); return v; }
  Buf clone() const { Buf b(n); if (n) d2d(b.p, p, sizeof(T) * (size_t)n); return b; }
};

// ---------------------------------------------------------------------------------------
// parallel_for: one logical thread per index
// ---------------------------------------------------------------------------------------
#ifndef AMGB_EMU
template <class F>
__global__ void __launch_bounds__(256) k_parallel_for(i64 n, F f) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  i64 stride = (i64)gridDim.x * blockDim.x;
  for (; i < n; i += stride) f(i);
}
template <class F>
inline void parallel_for(i64 n, F f) {
  if (n <= 0) return;
  Context &c = ctx();
  i64 blocks = (n + 255) / 256;
  i64 cap = (i64)c.sm_count * 32;
  if (blocks > cap) blocks = cap;
  k_parallel_for<<<(unsigned)blocks, 256, 0, c.stream>>>(n, f);
  c.launches++;
  post_launch("parallel_for");
}
#else
template <class F>
inline void parallel_for(i64 n, F f) {
  for (i64 i = 0; i < n; i++) f(i);
  ctx().launches++;
}
#endif

// atomics usable from HD lambdas
template <class T>
HD inline T atomic_add(T *p, T v) {
#ifdef __CUDA_ARCH__
  return atomicAdd(p, v);
#else
  T o = *p; *p = o + v; return o;
#endif
}
HD inline unsigned long long atomic_max_u64(unsigned long long *p, unsigned long long v) {
#ifdef __CUDA_ARCH__
  return atomicMax(p, v);
#else
  unsigned long long o = *p; if (v > o) *p = v; return o;
#endif
}
HD inline int atomic_max_i32(int *p, int v) {
#ifdef __CUDA_ARCH__
  return atomicMax(p, v);
#else
  int o = *p; if (v > o) *p = v; return o;
#endif
}
HD inline int atomic_min_i32(int *p, int v) {
#ifdef __CUDA_ARCH__
  return atomicMin(p, v);
#else
  int o = *p; if (v < o) *p = v; return o;
#endif
}
HD inline int atomic_cas_i32(int *p, int cmp, int val) {
#ifdef __CUDA_ARCH__
  return atomicCAS(p, cmp, val);
#else
  int o = *p; if (o == cmp) *p = val; return o;
#endif
}

// order-preserving map double -> u64 (for atomic max on doubles of either sign)
HD inline unsigned long long dbl_key(double x) {
  unsigned long long b;
#ifdef __CUDA_ARCH__
  b = (unsigned long long)__double_as_longlong(x);
#else
  memcpy(&b, &x, 8);
#endif
  return (b & 0x8000000000000000ULL) ? ~b : (b | 0x8000000000000000ULL);
}
HD inline double key_dbl(unsigned long long k) {
  unsigned long long b = (k & 0x8000000000000000ULL) ? (k & 0x7fffffffffffffffULL) : ~k;
  double x;
#ifdef __CUDA_ARCH__
  x = __longlong_as_double((long long)b);
#else
  memcpy(&x, &b, 8);
#endif
  return x;
}

// ---------------------------------------------------------------------------------------
// scans and reductions (scan.cu)
// ---------------------------------------------------------------------------------------
// out[i] = sum_{j<i} in[j] for i in [0,n]; out has n+1 entries; returns out[n] on the host.
i64 exclusive_scan(const int *in, int *out, i64 n);
i64 exclusive_scan64(const i64 *in, i64 *out, i64 n);
void exclusive_scan_dev(const int *in, int *out, i64 n);   // the total stays in out[n] on the device

// Deterministic tree sum of n doubles with the fixed shape documented in DESIGN.md
// (1024-value chunks, 256 "threads" x 4 strided values, strides 16..1 inside a warp,
// strides 4..1 over the 8 warps; chunk results reduced again by the same rule).
double tree_sum(const double *v, i64 n);
// sum_i a[i]*b[i] with the same tree (products rounded first)
double tree_dot(const double *a, const double *b, i64 n);
// the same sums accumulated left to right by one thread (reference order)
double seq_sum(const double *v, i64 n);
double seq_dot(const double *a, const double *b, i64 n);
// the reduction the context asks for
inline double vdot(const double *a, const double *b, i64 n) { return ctx().reduce_seq ? seq_dot(a, b, n) : tree_dot(a, b, n); }
inline double vsum(const double *a, i64 n) { return ctx().reduce_seq ? seq_sum(a, n) : tree_sum(a, n); }
inline double vnorm2(const double *a, i64 n) { return sqrt(vdot(a, a, n)); }
// the same reductions with the result left in a device scalar (no host synchronisation);
// b == nullptr gives the plain sum
void seq_dot_dev(double *out, const double *a, const double *b, i64 n);
void tree_dot_dev(double *out, const double *a, const double *b, i64 n);
inline void vdot_dev(double *out, const double *a, const double *b, i64 n) { if (ctx().reduce_seq) seq_dot_dev(out, a, b, n); else tree_dot_dev(out, a, b, n); }
inline void vsum_dev(double *out, const double *a, i64 n) { vdot_dev(out, a, nullptr, n); }
// largest value and the first index holding it (extr_op(max), amg_setup.c:3281)
void max_first(const double *v, i64 n, double *val, i64 *idx);
// both maxima of one coarsening round in one pass and one read-back
void max_first2(const double *a, const double *b, i64 n, double *amax, i64 *aidx, double *bmax);
// number of non-zero flags
i64 count_nonzero(const double *v, i64 n);

// optional sub-stage profile (AMGB_STAGE_LOG=1): wall time with a stream sync at scope exit,
// accumulated by name and printed at the end of setup().  Off by default (no syncs added).
struct StageTimer {
  const char *name; double t0; bool on;
  explicit StageTimer(const char *n);
  ~StageTimer();
};
void stage_report();
void stage_count(const char *name, long n);   // event counters shown by stage_report

// Test hook (env AMGB_TEST_SMALL_BINS=1, read at every call): the kernels meant for the large
// levels (long-row SpMV, block-per-column find_support, cluster / HBM Q builders, the optimistic
// block SpGEMM with its overflow path) take over at tiny sizes, so that the parity tests against
// the oracle, which only finishes small problems, run through them.  Never set in production.
bool test_small_bins();
// Test hook (env AMGB_TEST_FORCE, a string of letters, read at every call): routes that the size
// heuristics only choose on the large levels are forced on inputs small enough for the oracle.
//   't'  every SpGEMM goes through the transposed product (spgemm.cu)
//   'o'  the SpGEMM arena is far too small, so the overflow second pass runs
//   'b'  vector reductions use the many-block exact reduction whatever the length (runtime.cu)
//   'p'  the A-orthogonalisation treats every support as a large one (localsolve.cu panel kernels)
bool test_force(char what);

// trace: FNV-1a of a device array; the tests compare the tag/hash sequence with their checker
void trace_dev(const char *tag, const void *dptr, size_t bytes);

}  // namespace amgb

// comm.cu -- see comm.cuh.
#include "comm.cuh"

#ifndef AMGB_EMU
#include <dlfcn.h>
#endif

namespace amgb {

namespace {
struct CommState {
  int rank = 0, size = 1;
  i64 min_work = -1;
  i64 calls = 0, bytes = 0;
  host_allgatherv_fn host_fn = nullptr;
  void *host_user = nullptr;
#ifndef AMGB_EMU
  void *nccl = nullptr;                         // ncclComm_t
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
#endif
};
CommState g_comm;

#ifndef AMGB_EMU
// The few NCCL entry points used, bound from libnccl.so.2 at run time (the copy torch has
// already loaded when the process is a torch.distributed rank, the system one otherwise).
struct NcclId { char internal[128]; };
struct NcclApi {
  void *h = nullptr;
  int (*GetUniqueId)(NcclId *) = nullptr;
  int (*CommInitRank)(void **, int, NcclId, int) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  int (*Broadcast)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;

void nccl_load() {
  if (g_nccl.h) return;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW);
  if (!h) throw Error(-110, std::string("cannot load libnccl.so.2: ") + dlerror());
  auto sym = [&](const char *n) { void *p = dlsym(h, n); if (!p) throw Error(-110, std::string("libnccl lacks ") + n); return p; };
  g_nccl.GetUniqueId = (int (*)(NcclId *))sym("ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(void **, int, NcclId, int))sym("ncclCommInitRank");
  g_nccl.CommDestroy = (int (*)(void *))sym("ncclCommDestroy");
  g_nccl.Broadcast = (int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t))sym("ncclBroadcast");
  g_nccl.GroupStart = (int (*)())sym("ncclGroupStart");
  g_nccl.GroupEnd = (int (*)())sym("ncclGroupEnd");
  g_nccl.GetErrorString = (const char *(*)(int))sym("ncclGetErrorString");
  g_nccl.h = h;
}
void nccl_check(int rc, const char *what) {
  if (rc != 0) throw Error(-111, std::string("NCCL error in ") + what + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"));
}
#endif
}  // namespace

int comm_rank() { return g_comm.rank; }
int comm_size() { return g_comm.size; }

i64 comm_min_work() {
  if (g_comm.min_work < 0) {
    const char *e = getenv("AMGB_DIST_MIN_NNZ");
    g_comm.min_work = e ? atoll(e) : (i64)1 << 20;
  }
  return g_comm.min_work;
}

void comm_stats_reset() {
  g_comm.calls = 0; g_comm.bytes = 0;
#ifndef AMGB_EMU
  for (auto &e : g_comm.ev) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
  g_comm.ev.clear();
#endif
}
void comm_stats_get(i64 *calls, i64 *bytes, double *seconds) {
  double s = 0;
#ifndef AMGB_EMU
  for (auto &e : g_comm.ev) { float ms = 0; if (cudaEventElapsedTime(&ms, e.first, e.second) == cudaSuccess) s += ms * 1e-3; }
#endif
  *calls = g_comm.calls; *bytes = g_comm.bytes; *seconds = s;
}

void comm_finalize() {
#ifndef AMGB_EMU
  if (g_comm.nccl) { stream_sync(); g_nccl.CommDestroy(g_comm.nccl); g_comm.nccl = nullptr; }
#endif
  comm_stats_reset();
  g_comm.rank = 0; g_comm.size = 1; g_comm.host_fn = nullptr; g_comm.host_user = nullptr;
}

void comm_unique_id(unsigned char id[128]) {
#ifndef AMGB_EMU
  nccl_load();
  NcclId u;
  nccl_check(g_nccl.GetUniqueId(&u), "ncclGetUniqueId");
  memcpy(id, u.internal, 128);
#else
  (void)id;
  throw Error(-112, "the host-emulation build has no NCCL transport");
#endif
}

void comm_init_nccl(int rank, int size, const unsigned char id[128]) {
#ifndef AMGB_EMU
  if (size < 1 || rank < 0 || rank >= size) throw Error(-2, "comm_init: bad rank/size");
  comm_finalize();
  if (size == 1) return;
  nccl_load();
  NcclId u;
  memcpy(u.internal, id, 128);
  void *c = nullptr;
  nccl_check(g_nccl.CommInitRank(&c, size, u, rank), "ncclCommInitRank");
  g_comm.nccl = c; g_comm.rank = rank; g_comm.size = size;
#else
  (void)rank; (void)size; (void)id;
  throw Error(-112, "the host-emulation build has no NCCL transport");
#endif
}

void comm_init_host(int rank, int size, host_allgatherv_fn fn, void *user) {
#ifdef AMGB_EMU
  if (size < 1 || rank < 0 || rank >= size || (size > 1 && !fn)) throw Error(-2, "comm_init: bad rank/size/transport");
  comm_finalize();
  g_comm.rank = rank; g_comm.size = size; g_comm.host_fn = fn; g_comm.host_user = user;
#else
  (void)rank; (void)size; (void)fn; (void)user;
  throw Error(-112, "host transports exist only in the host-emulation build; the product exchanges through NCCL");
#endif
}

void comm_allgatherv(void *buf, const i64 *off, const char *what, bool timed) {
  const int P = g_comm.size;
  if (P <= 1) return;
  StageTimer st_(what);      // AMGB_STAGE_LOG: time waiting for + inside the exchange, per call site
  g_comm.calls++;
  g_comm.bytes += (off[P] - off[0]) - (off[g_comm.rank + 1] - off[g_comm.rank]);
#ifndef AMGB_EMU
  Context &c = ctx();
  // the timing events belong to the statistics once recorded; on an error path they are destroyed
  struct Pair {
    cudaEvent_t a = nullptr, b = nullptr;
    ~Pair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
  } ev;
  if (timed) { CUDA_CHECK(cudaEventCreate(&ev.a)); CUDA_CHECK(cudaEventCreate(&ev.b)); }
  cudaEvent_t e0 = ev.a, e1 = ev.b;
  if (timed) CUDA_CHECK(cudaEventRecord(e0, c.stream));
  nccl_check(g_nccl.GroupStart(), "ncclGroupStart");
  for (int r = 0; r < P; r++) {
    const i64 n = off[r + 1] - off[r];
    if (n <= 0) continue;
    char *seg = (char *)buf + off[r];
    nccl_check(g_nccl.Broadcast(seg, seg, (size_t)n, /*ncclChar*/ 0, r, g_comm.nccl, c.stream), "ncclBroadcast");
  }
  nccl_check(g_nccl.GroupEnd(), "ncclGroupEnd");
  if (timed) {
    CUDA_CHECK(cudaEventRecord(e1, c.stream));
    g_comm.ev.emplace_back(e0, e1);
    ev.a = ev.b = nullptr;
  }
#else
  std::vector<long long> o(off, off + P + 1);
  if (g_comm.host_fn(buf, o.data(), P, g_comm.host_user) != 0) throw Error(-113, "host transport failed");
#endif
}

}  // namespace amgb

// comm.cuh -- the exchange step of the row-partitioned stages (one process per GPU).
//
// A partitioned stage (SpGEMM rows, the local solves of the coarse columns) leaves every rank
// with one contiguous segment of a result array; the other ranks' segments arrive through an
// in-place all-gather of variable-length segments.  On the product build that is NCCL over
// NVLink (grouped ncclBroadcast of the segments, libnccl.so.2 bound at run time so that a
// single-GPU process does not need it); on the host-emulation build (tests) it is a callback the
// test harness provides (gloo).  Because every output is computed by exactly one rank with the
// single-GPU kernels, a partitioned setup is bit-identical to the single-GPU one.
#pragma once
#include "common.cuh"

namespace amgb {

int comm_rank();
int comm_size();
inline bool comm_active() { return comm_size() > 1; }

// smallest operand size (entries) from which a stage is partitioned (env AMGB_DIST_MIN_NNZ)
i64 comm_min_work();

// rank r owns bytes [off[r], off[r+1]) of buf (off has size+1 host entries); after the call
// every rank holds all segments
// timed = false: no timing events (the V-cycle's vector exchanges: thousands per solve phase)
void comm_allgatherv(void *buf, const i64 *off, const char *what = "comm.exchange", bool timed = true);

// rows [row_split(r), row_split(r+1)) of an n-row stage belong to rank r
inline i64 row_split(i64 n, int r) { return n * r / comm_size(); }

// NCCL plumbing (capi)
void comm_unique_id(unsigned char id[128]);
void comm_init_nccl(int rank, int size, const unsigned char id[128]);
typedef int (*host_allgatherv_fn)(void *buf, const long long *off, int size, void *user);
void comm_init_host(int rank, int size, host_allgatherv_fn fn, void *user);
void comm_finalize();
// calls, bytes received by this rank, device seconds inside the exchanges (product build)
void comm_stats_get(i64 *calls, i64 *bytes, double *seconds);
void comm_stats_reset();

}  // namespace amgb

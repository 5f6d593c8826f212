/*
 * omp_amg_b200.h -- C ABI of the B200 AMG hierarchy-construction engine.
 *
 * Drop-in boundary for the serial AMG setup path of nicooff/omp_amg (gslib's AMG coarse
 * solver).  Plain pointers and sizes only.  Every entry point names the reference interface it
 * replaces.  All functions return 0 on success and a negative code on failure;
 * amgb_last_error() gives the message.  There is no CPU path: without a CUDA device
 * amgb_init / amgb_setup fail with code -101.
 *
 * Data layout handed back by the accessors is the reference's struct csr_mat
 * (amg_tools.h:5): row offsets, ascending column indices per row, fp64 values -- with int32
 * indices instead of the reference's build-time "uint".
 */
#ifndef OMP_AMG_B200_H
#define OMP_AMG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct amgb_hier amgb_hier;

/* ---- runtime ---- */
int amgb_device_count(void);
int amgb_init(int device);                 /* device < 0: keep the current device */
const char *amgb_last_error(void);
const char *amgb_build_info(void);         /* "cuda sm_100a ..." or "host-emulation (tests)" */

/* ---- setup: replaces amg_setup() (amg_setup.h:5, amg_setup.c:60) ----
 * Input is the assembled matrix as COO triplets, 0-based, exactly what serial_amg.c:76-96
 * builds from amgdmp_{i,j,p}.dat.  Zero values are dropped, empty rows/columns removed
 * (build_csr, amg_setup.c:3612).  The hierarchy stays resident in HBM. */
int amgb_setup(int64_t nnz, const int32_t *Ai, const int32_t *Aj, const double *Av,
               amgb_hier **out);                                   /* HOST buffers */
int amgb_setup_device(int64_t nnz, const int32_t *dAi, const int32_t *dAj, const double *dAv,
                      amgb_hier **out);                            /* buffers already in HBM */
/* reads amgdmp_i.dat, amgdmp_j.dat, amgdmp_p.dat from dir (serial_amg.c main) */
int amgb_setup_from_dump(const char *dir, amgb_hier **out);
void amgb_free(amgb_hier *h);

/* ---- struct amg_setup_data accessors (amg_tools.h:29) ---- */
int amgb_nlevels(const amgb_hier *h);      /* data->nlevels   */
int amgb_nullspace(const amgb_hier *h);    /* data->nullspace */
/* info[0]=n info[1]=nnz(A) info[2]=nf info[3]=nc info[4]=nnz(Af) info[5]=nnz(W) info[6]=nnz(AfP)
   info[7]=coarsening rounds info[8]=Lanczos iterations info[9]=interpolation rounds */
int amgb_level_info(const amgb_hier *h, int lvl, int64_t info[10]);
/* par[0]=data->m[lvl] par[1]=data->rho[lvl] par[2],par[3]=Lanczos lambda_min, lambda_max */
int amgb_level_params(const amgb_hier *h, int lvl, double par[4]);
enum { AMGB_A = 0, AMGB_AF = 1, AMGB_W = 2, AMGB_AFP = 3 };   /* data->A/Af/W/AfP[lvl] */
/* Copies one matrix to HOST buffers; pass NULL arrays to query rn/cn/nnz first. */
int amgb_get_csr(const amgb_hier *h, int lvl, int which, int32_t *rn, int32_t *cn, int64_t *nnz,
                 int32_t *row_off, int32_t *col, double *a);
enum { AMGB_C = 0, AMGB_D = 1, AMGB_IDC = 2, AMGB_IDF = 3 };  /* data->C/D/idc/idf[lvl] */
int amgb_get_vec(const amgb_hier *h, int lvl, int which, double *out);

/* ---- replaces amg_export() (amg_setup.h:9, amg_setup.c:405) ----
 * Writes amg.dat, amg_W.dat, amg_AfP.dat, amg_Aff.dat into dir in the reference's format
 * (the files amg.c:813 read_data consumes). */
int amgb_export(const amgb_hier *h, const char *dir);

/* ---- V-cycle: replaces amg_exec()+crs_solve() (amg.c:114, amg.c:171), one process ---- */
int amgb_solve(const amgb_hier *h, double *x, const double *b);                /* HOST vectors */
int amgb_solve_device(const amgb_hier *h, double *dx, const double *db);      /* HBM vectors  */

/* ---- measurement ----
 * t[0]=total t[1]=build_csr t[2]=coarsen t[3]=smoother t[4]=lanczos t[5]=interpolation
 * t[6]=galerkin (host seconds, stream synchronised at stage ends)
 * t[7]=device seconds inside SpGEMM kernels  t[8]=their algorithmic bytes  t[9]=SpGEMM calls
 * t[10]=kernel launches  t[11]=host syncs
 * t[12]=CUDA-event seconds from the first to the last kernel of the setup
 * t[13]=exchanges of the row-partitioned stages  t[14]=bytes this rank received in them
 * t[15]=device seconds inside the exchanges */
int amgb_timing(const amgb_hier *h, double t[16]);
/* Opt-in statistics of the long-row SpMV kernels (matrices with more than 24 entries per row) of
 * the setups that follow: out[0]=device seconds (CUDA events around every call), out[1]=their
 * algorithmic bytes (12 B per entry -- 8 B when the vector is an implicit vector of ones -- plus
 * 12 or 20 B per row), out[2]=calls.  Off by default (two events per call). */
int amgb_spmv_stats_enable(int on);
int amgb_spmv_stats(const amgb_hier *h, double out[3]);

/* ---- several GPUs: one process per GPU, row-partitioned stages ----
 * The reference distributes the coarse problem over MPI ranks (struct comm, crs.h:14; the setup
 * itself is serial, serial_amg.c).  Here every rank calls amgb_setup with the same matrix; the
 * SpGEMM rows (mxm, amg_setup.c:1894) and the local solves of the coarse columns (interp,
 * amg_setup.c:2053) are partitioned over the ranks and the blocks are exchanged through NCCL
 * (NVLink), so that every rank ends with the same hierarchy, bit-identical to one GPU's.
 * amgb_solve / crs_amg_solve on several ranks (same b on every rank) row-partition the
 * matrix-vector products of the large levels of the V-cycle (amg.c:114; the reference applies its
 * distributed matrices the same way, with gs halo exchanges amg.c:85-112) and exchange the blocks
 * of the result vectors; the solution is bit-identical to one GPU's on every rank.
 * Rank 0 obtains an id (amgb_comm_unique_id), the host program hands it to the other ranks
 * (MPI_Bcast / torch.distributed), then every rank calls amgb_comm_init on its own device.
 * amgb_comm_init_host installs a host transport instead and exists only in the host-emulation
 * build used by the CPU tests (the product build returns -112). */
int amgb_comm_unique_id(uint8_t id[128]);
int amgb_comm_init(int rank, int size, const uint8_t id[128]);
/* in-place all-gather: rank r owns bytes [off[r], off[r+1]) of buf; returns 0 on success */
typedef int (*amgb_allgatherv_fn)(void *buf, const long long *off, int size, void *user);
int amgb_comm_init_host(int rank, int size, amgb_allgatherv_fn fn, void *user);
int amgb_comm_finalize(void);
int amgb_comm_rank(void);
int amgb_comm_size(void);
/* Partitioned STORAGE for the solve phase.  After a setup on several ranks every rank holds the
 * whole hierarchy; this call (collective, every rank) keeps, of every matrix the V-cycle applies
 * row-partitioned (W', W, AfP, Af of the large levels, amg.c:125-152), only this rank's row block
 * and releases the rest -- the layout the reference's crs_data has after amg_setup_mats
 * (amg.c:295-377: every rank stores its rows).  amgb_solve / crs_amg_solve keep working,
 * bit-identically; the accessors of those matrices, amgb_export and amgb_hierarchy_hash then
 * return -120.  released_bytes: device bytes this rank released.  One rank: no-op. */
int amgb_partition_solve_storage(amgb_hier *h, int64_t *released_bytes);
/* exchanges issued and bytes received by this rank since the last amgb_setup (setup stages and
 * the row-partitioned V-cycles after it; amgb_setup resets the counters) */
int amgb_comm_stats(int64_t *calls, int64_t *bytes);

/* ---- device memory ----
 * Temporaries come from a caching allocator inside the library; amgb_release_memory() hands the
 * cached (free) blocks back to the driver, amgb_peak_device_bytes() is the high-water mark. */
void amgb_release_memory(void);
int64_t amgb_peak_device_bytes(void);

/* ---- arithmetic mode of the vector-length reductions (dot products, 2-norms) ----
 * AMGB_REDUCE_SEQUENTIAL (default): summed left to right like the reference's vv_dot
 *   (amg_setup.c:3193); the hierarchy is then bit-identical to the reference's.
 * AMGB_REDUCE_TREE: fixed-shape parallel tree; deterministic and much faster, values agree to
 *   rounding, but the reference algorithm's threshold decisions are sensitive to the last bit, so
 *   C/F splits and patterns can differ from the reference's (DESIGN.md "Reduction order").
 * The environment variable AMGB_REDUCE=tree|seq sets the initial mode. */
enum { AMGB_REDUCE_TREE = 0, AMGB_REDUCE_SEQUENTIAL = 1 };
int amgb_set_reduce_mode(int mode);
int amgb_get_reduce_mode(void);

/* diagnostics: sum_i a[i]*b[i] (b == NULL: sum_i a[i]) of HOST vectors with the library's
 * reduction kernels; mode as above.  Mode 1 equals the plain left-to-right loop bit for bit. */
int amgb_debug_dot(const double *a, const double *b, int64_t n, int mode, double *out);

/* diagnostics: X = A*B with the library's SpGEMM on HOST CSR arrays (columns ascending in every
 * row).  Semantics of mxm (amg_setup.c:1894): X[i][c] = sum over k ascending of B[k][c]*A[i][k],
 * separate multiply and add, entries whose sum is exactly 0 are not stored.  xro has arn+1
 * entries; xcol/xa hold up to cap entries (-30 and *xnnz set if X is larger). */
int amgb_debug_spgemm(int32_t arn, int32_t acn, const int32_t *aro, const int32_t *acol, const double *aa,
                      int32_t brn, int32_t bcn, const int32_t *bro, const int32_t *bcol, const double *ba,
                      int64_t cap, int64_t *xnnz, int32_t *xro, int32_t *xcol, double *xa);

/* rows per SpGEMM tier of the last amgb_debug_spgemm call: out[0..11) as binned, out[11..22) after the
 * optimistic hand-downs (tile8, tile32, warp512, warp2048, block4096, block8192, HBM table, global,
 * block-dense <= 8192 columns, block-dense <= 22000 columns, warp-dense <= 4608 columns) */
int amgb_debug_spgemm_tiers(int32_t out[22]);

/* ---- stage trace (debug): FNV-1a hashes of intermediate arrays, in stage order ---- */
void amgb_trace_enable(int on);
int amgb_trace_count(void);
int amgb_trace_get(int i, char *tag, int taglen, uint64_t *hash, int64_t *bytes);

/* ---- hierarchy fingerprint ----
 * One number over the whole hierarchy (level count, nullspace, and per level the row offsets,
 * columns and value bits of A, Af, W, AfP, the bits of C and D, idc, idf, m, rho), computed on the
 * device.  bench.py prints it so that runs on 1, 2, 4 and 8 GPUs can be compared; the tests compare
 * it with the same function of the oracle's hierarchy (oracle/oracle.py: hierarchy_hash). */
int amgb_hierarchy_hash(const amgb_hier *h, uint64_t *out);

/* kernels launched / host synchronisations by this library since it was loaded */
int64_t amgb_launch_count(void);
int64_t amgb_sync_count(void);

/* ---- V-cycle with workspaces kept across calls, repeated on device vectors (bench --metric solve) */
int amgb_solve_device_repeat(const amgb_hier *h, double *dx, const double *db, int repeat);

/* ---- gslib coarse-solver slot: crs.h:12-22 ----
 * Same names, argument meaning and order as the reference (with the crs_amg_ prefix gslib
 * gives its AMG variant).  gslib fixes its integer widths at build time (types.h:52-64): "uint"
 * is unsigned int, unsigned long (-DUSE_LONG) or unsigned long long (-DUSE_LONG_LONG) and
 * "ulong" likewise with -DGLOBAL_LONG / -DGLOBAL_LONG_LONG; the reference Makefile builds with
 * -DUSE_LONG -DGLOBAL_LONG.  The library exports one entry point per local width and this header
 * picks the one matching the includer's gslib flags, so that a caller compiled with the
 * reference's own CFLAGS binds to the right ABI without a conversion shim:
 *      crs_amg_setup  ->  crs_amg_setup_u32 | crs_amg_setup_u64
 * ("ulong" ids are always passed as 64-bit; with neither GLOBAL_LONG flag gslib's ulong is a
 * 32-bit unsigned int and the caller must widen the id array.)
 * One process: comm must be NULL or describe np == 1 (struct comm, comm.h:85: { uint id, np;
 * comm_ext c; } with the same build-time uint).  id[i] == 0 marks a dof that is not solved for;
 * repeated ids are the same dof (entries are summed, as gs_setup/assign_dofs do in amg.c:399). */
struct comm;                       /* gslib's struct comm (comm.h:85); only np == 1 is accepted */
struct crs_data;
struct crs_data *crs_amg_setup_u32(uint32_t n, const uint64_t *id, uint32_t nz, const uint32_t *Ai,
                                   const uint32_t *Aj, const double *A, uint32_t null_space,
                                   const struct comm *comm);
struct crs_data *crs_amg_setup_u64(uint64_t n, const uint64_t *id, uint64_t nz, const uint64_t *Ai,
                                   const uint64_t *Aj, const double *A, uint64_t null_space,
                                   const struct comm *comm);
#if defined(USE_LONG) || defined(USE_LONG_LONG)
#define crs_amg_setup crs_amg_setup_u64
#else
#define crs_amg_setup crs_amg_setup_u32
#endif
void crs_amg_solve(double *x, struct crs_data *data, double *b);
void crs_amg_stats(struct crs_data *data);
void crs_amg_free(struct crs_data *data);
amgb_hier *crs_amg_hierarchy(struct crs_data *data);

#ifdef __cplusplus
}
#endif
#endif

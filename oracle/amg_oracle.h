/*
 * oracle/amg_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the serial AMG setup of nicooff/omp_amg
 * (amg_setup.c / amg_tools.c / serial_amg.c) and of the V-cycle in amg.c.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may load this library.  The product (omp_amg_b200/)
 * never links, imports or calls it.
 *
 * Parity status: PINNED.  In reduction mode AMGO_REDUCE_SEQ every array this
 * oracle produces is compared bit-for-bit with the unmodified reference
 * compiled from /root/reference into oracle/_ref (tests/test_oracle_vs_ref.py)
 * and with committed fixtures under tests/golden/ generated from that build.
 */
#ifndef AMG_ORACLE_H
#define AMG_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* CSR matrix, nnz = ro[rn]; columns sorted ascending inside a row
   (reference: struct csr_mat, amg_tools.h:5). */
typedef struct {
  int rn, cn;
  int *ro;
  int *col;
  double *a;
} ocsr;

/* how vector-length reductions (dot, 2-norm) are ordered */
enum { AMGO_REDUCE_SEQ = 0,  /* left-to-right, as the reference (vv_dot, array_op) */
       AMGO_REDUCE_TREE = 1  /* fixed 1024-chunk tree, as the CUDA product */ };

typedef struct amgo_hier amgo_hier;

/* glibc TYPE_3 rand() restated: ring of the last 34 words; see amg_oracle.c */
typedef struct { uint32_t r[34]; int pos; int64_t count; } amgo_rng;
void amgo_rng_seed(amgo_rng *g, uint32_t seed);
int32_t amgo_rng_next(amgo_rng *g);


/* amg_setup (amg_setup.c:60).  Ai/Aj are 0-based COO indices. Returns 0 on success. */
int amgo_setup(int64_t nnz, const int32_t *Ai, const int32_t *Aj, const double *Av,
               int reduce_mode, amgo_hier **out);
void amgo_free(amgo_hier *h);

int amgo_nlevels(const amgo_hier *h);
int amgo_nullspace(const amgo_hier *h);
/* info[0]=n  info[1]=nnz(A)  info[2]=nf  info[3]=nc  info[4]=nnz(Af) info[5]=nnz(W)
   info[6]=nnz(AfP) info[7]=coarsen rounds info[8]=lanczos k info[9]=interp rounds */
int amgo_level_info(const amgo_hier *h, int lvl, int64_t info[10]);
/* par[0]=m (Chebyshev iterations) par[1]=rho par[2]=lambda_min par[3]=lambda_max */
int amgo_level_params(const amgo_hier *h, int lvl, double par[4]);
/* which: 0=A 1=Af 2=W 3=AfP.  Pass NULL pointers to query sizes only. */
int amgo_get_csr(const amgo_hier *h, int lvl, int which, int *rn, int *cn, int64_t *nnz,
                 int32_t *ro, int32_t *col, double *a);
/* which: 0=C flags (n) 1=D (nf) 2=idc (nc, as double) 3=idf (nf, as double) */
int amgo_get_vec(const amgo_hier *h, int lvl, int which, double *out);

/* amg_export (amg_setup.c:405): writes amg.dat amg_W.dat amg_AfP.dat amg_Aff.dat into dir */
int amgo_export(const amgo_hier *h, const char *dir);

/* V-cycle of amg.c:114 (amg_exec) on the hierarchy, single process:
   x = V(b), then null-space projection as crs_solve (amg.c:171). */
int amgo_solve(const amgo_hier *h, double *x, const double *b);

/* --- single stages, exported for unit pinning against oracle/_ref --- */
ocsr *amgo_csr_new(int rn, int cn, int64_t nnz);
void amgo_csr_free(ocsr *A);
ocsr *amgo_build_csr(int64_t nnz, const int32_t *Ai, const int32_t *Aj, const double *Av);
ocsr *amgo_transpose(const ocsr *A);
ocsr *amgo_spgemm(const ocsr *A, const ocsr *B);               /* mxm, amg_setup.c:1894 */
ocsr *amgo_mpm(double alpha, const ocsr *A, double beta, const ocsr *B); /* :1684 */
ocsr *amgo_mxmpoint(const ocsr *A, const ocsr *B);              /* :1807 */
ocsr *amgo_sub_mat(const ocsr *A, const double *vr, const double *vc);   /* :3058 */
int amgo_coarsen(double *vc, const ocsr *A, double ctol);       /* :2737, returns rounds */
ocsr *amgo_interpolation(const ocsr *Af, const ocsr *Ac, const ocsr *Ar, double gamma2,
                         double tol, int reduce_mode, int *rounds);       /* :598 */
int amgo_lanczos(double *lambda, const ocsr *A, int reduce_mode, amgo_rng *rng,
                 int *iters);                                    /* :2435 */
void amgo_chebsim(double *m, double *c, double rho, double tol); /* :2412 */
int amgo_pcg(double *x, const ocsr *A, double *r, const double *M, double tol,
             const double *b, int reduce_mode);                  /* :2242 */
double amgo_dot(const double *a, const double *b, int64_t n, int reduce_mode);

/* trace of stage hashes (debug aid shared with the CUDA product's trace) */
void amgo_trace_enable(int on);
int amgo_trace_count(void);
int amgo_trace_get(int i, char *tag, int taglen, uint64_t *hash, int64_t *bytes);

#ifdef __cplusplus
}
#endif
#endif

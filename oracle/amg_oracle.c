/*
 * oracle/amg_oracle.c -- TEST INFRASTRUCTURE ONLY (see amg_oracle.h).
 *
 * CPU restatement of the reference's serial AMG setup.  Every function cites the
 * reference lines it follows (amg_setup.c unless another file is named).  The
 * floating-point operation ORDER of the reference is kept (left-to-right row sums,
 * separate multiply and add, reciprocal-then-multiply where the reference does that),
 * so that in AMGO_REDUCE_SEQ mode all outputs are bit-identical to the reference
 * compiled from /root/reference (oracle/_ref).  What is NOT kept is the reference's
 * asymptotics: its mxm() is O(rows_A * rows_B) and its expand_support() builds dense
 * rank x row matrices; here they are O(nnz) row-wise algorithms with the same results.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (no FMA contraction, IEEE doubles).
 */
#include "amg_oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* small helpers                                                              */
/* ------------------------------------------------------------------------- */
static int g_verbose = -1;
static int verbose(void) {
  if (g_verbose < 0) { const char *e = getenv("AMGO_VERBOSE"); g_verbose = (e && *e && *e != '0') ? 1 : 0; }
  return g_verbose;
}
#define VLOG(...) do { if (verbose()) { fprintf(stderr, __VA_ARGS__); fflush(stderr); } } while (0)
static void *xmalloc(size_t n) {
  void *p = malloc(n ? n : 1);
  if (!p) { fprintf(stderr, "amg_oracle: out of memory (%zu bytes)\n", n); abort(); }
  return p;
}
#define NEW(T, n) ((T *)xmalloc(sizeof(T) * (size_t)(n)))

/* ---- trace: FNV-1a hashes of intermediate arrays, shared scheme with the product ---- */
typedef struct { char tag[64]; uint64_t hash; int64_t bytes; } trace_rec;
static int g_trace_on = 0, g_trace_n = 0, g_trace_cap = 0;
static trace_rec *g_trace = NULL;
static char g_trace_prefix[32] = "";

static uint64_t fnv1a(const void *p, size_t n) {
  const unsigned char *b = (const unsigned char *)p;
  uint64_t h = 1469598103934665603ULL;
  for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ULL; }
  return h;
}
static void trace(const char *tag, const void *p, size_t bytes) {
  if (!g_trace_on) return;
  if (g_trace_n == g_trace_cap) {
    g_trace_cap = g_trace_cap ? 2 * g_trace_cap : 1024;
    g_trace = (trace_rec *)realloc(g_trace, sizeof(trace_rec) * (size_t)g_trace_cap);
  }
  trace_rec *r = &g_trace[g_trace_n++];
  snprintf(r->tag, sizeof r->tag, "%s%s", g_trace_prefix, tag);
  r->hash = fnv1a(p, bytes);
  r->bytes = (int64_t)bytes;
}
static void trace_csr(const char *tag, const ocsr *A) {
  if (!g_trace_on) return;
  char t[64];
  int64_t nnz = A->ro[A->rn];
  snprintf(t, sizeof t, "%s.ro", tag);  trace(t, A->ro, sizeof(int) * (size_t)(A->rn + 1));
  snprintf(t, sizeof t, "%s.col", tag); trace(t, A->col, sizeof(int) * (size_t)nnz);
  snprintf(t, sizeof t, "%s.a", tag);   trace(t, A->a, sizeof(double) * (size_t)nnz);
}
void amgo_trace_enable(int on) { g_trace_on = on; g_trace_n = 0; }
int amgo_trace_count(void) { return g_trace_n; }
int amgo_trace_get(int i, char *tag, int taglen, uint64_t *hash, int64_t *bytes) {
  if (i < 0 || i >= g_trace_n) return -1;
  snprintf(tag, (size_t)taglen, "%s", g_trace[i].tag);
  *hash = g_trace[i].hash; *bytes = g_trace[i].bytes;
  return 0;
}

/* ------------------------------------------------------------------------- */
/* glibc rand(): the reference seeds nothing, so lanczos (amg_setup.c:2447)   */
/* consumes glibc's TYPE_3 additive-feedback stream from seed 1.              */
/* r[i] = r[i-31] + r[i-3] (mod 2^32), output r[i] >> 1, first 310 discarded. */
/* ------------------------------------------------------------------------- */
void amgo_rng_seed(amgo_rng *g, uint32_t seed) {
  int32_t init[34];
  init[0] = (int32_t)(seed ? seed : 1);
  for (int i = 1; i < 31; i++) {
    int64_t v = (16807LL * init[i - 1]) % 2147483647LL;
    if (v < 0) v += 2147483647LL;
    init[i] = (int32_t)v;
  }
  for (int i = 31; i < 34; i++) init[i] = init[i - 31];
  /* ring of the last 34 words: word i lives in r[i % 34] */
  for (int i = 0; i < 34; i++) g->r[i] = (uint32_t)init[i];
  g->count = 0;
  for (int i = 34; i < 344; i++) {
    uint32_t v = g->r[(i - 31) % 34] + g->r[(i - 3) % 34];
    g->r[i % 34] = v;
  }
  g->pos = 34 + (344 % 34);
}
int32_t amgo_rng_next(amgo_rng *g) {
  int i = g->pos;                       /* kept in [34,68): only i mod 34 matters */
  uint32_t v = g->r[(i - 31) % 34] + g->r[(i - 3) % 34];
  g->r[i % 34] = v;
  g->pos = 34 + ((i + 1) % 34);
  g->count++;
  return (int32_t)(v >> 1);
}

/* ------------------------------------------------------------------------- */
/* vector-length reductions                                                   */
/* ------------------------------------------------------------------------- */
/* Fixed tree used by the CUDA product (csrc/reduce.cuh): 1024-value chunks; inside a
   chunk "thread" t<256 adds values t, t+256, t+512, t+768 left to right, each of the 8
   "warps" folds its 32 partials with strides 16,8,4,2,1, the 8 warp results fold with
   strides 4,2,1; chunk results are reduced again by the same rule. */
static double tree_chunk(const double *v, int64_t m) {
  double s[256];
  for (int t = 0; t < 256; t++) {
    double x = (t < m) ? v[t] : 0.0;
    x = x + ((t + 256 < m) ? v[t + 256] : 0.0);
    x = x + ((t + 512 < m) ? v[t + 512] : 0.0);
    x = x + ((t + 768 < m) ? v[t + 768] : 0.0);
    s[t] = x;
  }
  double w[8];
  for (int k = 0; k < 8; k++) {
    double *x = s + 32 * k;
    for (int off = 16; off >= 1; off >>= 1)
      for (int l = 0; l < off; l++) x[l] = x[l] + x[l + off];
    w[k] = x[0];
  }
  for (int off = 4; off >= 1; off >>= 1)
    for (int l = 0; l < off; l++) w[l] = w[l] + w[l + off];
  return w[0];
}
static double tree_sum(const double *v, int64_t n) {
  if (n <= 0) return 0.0;
  if (n <= 1024) return tree_chunk(v, n);
  int64_t nc = (n + 1023) / 1024;
  double *part = NEW(double, nc);
  for (int64_t c = 0; c < nc; c++) {
    int64_t m = n - c * 1024; if (m > 1024) m = 1024;
    part[c] = tree_chunk(v + c * 1024, m);
  }
  double r = tree_sum(part, nc);
  free(part);
  return r;
}

/* vv_dot (amg_setup.c:3193) */
double amgo_dot(const double *a, const double *b, int64_t n, int mode) {
  if (mode == AMGO_REDUCE_SEQ) {
    double r = 0;
    for (int64_t i = 0; i < n; i++) r += a[i] * b[i];
    return r;
  }
  double *p = NEW(double, n);
  for (int64_t i = 0; i < n; i++) p[i] = a[i] * b[i];
  double r = tree_sum(p, n);
  free(p);
  return r;
}
/* array_op(.., norm2_op) (amg_setup.c:3309) */
static double norm2(const double *a, int64_t n, int mode) { return sqrt(amgo_dot(a, a, n, mode)); }

/* extr_op(max) (amg_setup.c:3281): first element, then strict > */
static double max_first(const double *a, int n, int *idx) {
  double m = a[0]; int k = 0;
  for (int i = 1; i < n; i++) if (a[i] > m) { m = a[i]; k = i; }
  if (idx) *idx = k;
  return m;
}

/* ------------------------------------------------------------------------- */
/* CSR primitives                                                             */
/* ------------------------------------------------------------------------- */
ocsr *amgo_csr_new(int rn, int cn, int64_t nnz) {   /* malloc_csr, :3453 */
  ocsr *A = NEW(ocsr, 1);
  A->rn = rn; A->cn = cn;
  A->ro = NEW(int, rn + 1);
  A->col = NEW(int, nnz);
  A->a = NEW(double, nnz);
  A->ro[0] = 0;
  return A;
}
void amgo_csr_free(ocsr *A) {                        /* free_csr, :3474 */
  if (!A) return;
  free(A->ro); free(A->col); free(A->a); free(A);
}
static ocsr *csr_copy(const ocsr *A) {               /* copy_csr, :3463 */
  int64_t nnz = A->ro[A->rn];
  ocsr *B = amgo_csr_new(A->rn, A->cn, nnz);
  memcpy(B->ro, A->ro, sizeof(int) * (size_t)(A->rn + 1));
  memcpy(B->col, A->col, sizeof(int) * (size_t)nnz);
  memcpy(B->a, A->a, sizeof(double) * (size_t)nnz);
  return B;
}
static int64_t nnz_of(const ocsr *A) { return A->ro[A->rn]; }

/* apply_M (amg_tools.c:71): z = alpha*y + beta*(M x), row sums left to right */
static void apply_M(double *z, double alpha, const double *y, double beta, const ocsr *M,
                    const double *x) {
  for (int i = 0; i < M->rn; i++) {
    double t = 0;
    for (int j = M->ro[i]; j < M->ro[i + 1]; j++) t += M->a[j] * x[M->col[j]];
    if (alpha == 0. || y == NULL) z[i] = beta * t;
    else z[i] = alpha * y[i] + beta * t;
  }
}
/* apply_Mt (amg_tools.c:97): z = M^t x; per column the rows arrive in ascending order */
static void apply_Mt(double *z, const ocsr *M, const double *x) {
  for (int i = 0; i < M->cn; i++) z[i] = 0;
  for (int i = 0; i < M->rn; i++) {
    double xi = x[i];
    for (int j = M->ro[i]; j < M->ro[i + 1]; j++) z[M->col[j]] += M->a[j] * xi;
  }
}

/* transpose (:2000): the reference sorts COO by (col,row); a counting pass that walks the
   rows in ascending order gives the same arrangement. */
ocsr *amgo_transpose(const ocsr *A) {
  int64_t nnz = nnz_of(A);
  ocsr *T = amgo_csr_new(A->cn, A->rn, nnz);
  int *cnt = NEW(int, A->cn + 1);
  memset(cnt, 0, sizeof(int) * (size_t)(A->cn + 1));
  for (int64_t k = 0; k < nnz; k++) cnt[A->col[k] + 1]++;
  for (int c = 0; c < A->cn; c++) cnt[c + 1] += cnt[c];
  memcpy(T->ro, cnt, sizeof(int) * (size_t)(A->cn + 1));
  for (int i = 0; i < A->rn; i++)
    for (int j = A->ro[i]; j < A->ro[i + 1]; j++) {
      int p = cnt[A->col[j]]++;
      T->col[p] = i; T->a[p] = A->a[j];
    }
  free(cnt);
  return T;
}

/* coo2csr / build_csr_dim (:3656,:3684): drop exact zeros, sort by (row,col).
   Entries must be unique (the reference would keep duplicates as repeated columns). */
typedef struct { int i, j; double v; } coo_t;
static int coo_cmp(const void *a, const void *b) {
  const coo_t *x = (const coo_t *)a, *y = (const coo_t *)b;
  if (x->i != y->i) return x->i < y->i ? -1 : 1;
  if (x->j != y->j) return x->j < y->j ? -1 : 1;
  return 0;
}
static ocsr *build_csr_dim(int64_t n, const int *Ai, const int *Aj, const double *Av, int rn,
                           int cn) {
  coo_t *c = NEW(coo_t, n);
  int64_t k = 0;
  for (int64_t i = 0; i < n; i++)
    if (Av[i] != 0.) { c[k].i = Ai[i]; c[k].j = Aj[i]; c[k].v = Av[i]; k++; }
  int sorted = 1;
  for (int64_t i = 1; i < k && sorted; i++) if (coo_cmp(&c[i - 1], &c[i]) > 0) sorted = 0;
  if (!sorted) qsort(c, (size_t)k, sizeof(coo_t), coo_cmp);
  ocsr *A = amgo_csr_new(rn, cn, k);
  memset(A->ro, 0, sizeof(int) * (size_t)(rn + 1));
  for (int64_t i = 0; i < k; i++) { A->ro[c[i].i + 1]++; A->col[i] = c[i].j; A->a[i] = c[i].v; }
  for (int r = 0; r < rn; r++) A->ro[r + 1] += A->ro[r];
  free(c);
  return A;
}

/* sub_mat (:3058): subA = A(vr != 0, vc != 0), columns renumbered */
ocsr *amgo_sub_mat(const ocsr *A, const double *vr, const double *vc) {
  int *g2l = NEW(int, A->cn);
  int subcn = 0, subrn = 0;
  int64_t nnz = 0;
  for (int i = 0; i < A->cn; i++) g2l[i] = (vc[i] != 0) ? subcn++ : -1;
  for (int i = 0; i < A->rn; i++)
    if (vr[i] != 0) {
      subrn++;
      for (int j = A->ro[i]; j < A->ro[i + 1]; j++) if (vc[A->col[j]] != 0) nnz++;
    }
  ocsr *S = amgo_csr_new(subrn, subcn, nnz);
  int r = 0; int64_t p = 0;
  for (int i = 0; i < A->rn; i++)
    if (vr[i] != 0) {
      for (int j = A->ro[i]; j < A->ro[i + 1]; j++)
        if (vc[A->col[j]] != 0) { S->col[p] = g2l[A->col[j]]; S->a[p] = A->a[j]; p++; }
      S->ro[++r] = (int)p;
    }
  free(g2l);
  return S;
}

/* build_csr (:3612): assemble, then remove empty rows (and the same-numbered columns) */
ocsr *amgo_build_csr(int64_t n, const int32_t *Ai, const int32_t *Aj, const double *Av) {
  int rn = 0, cn = 0;
  for (int64_t i = 0; i < n; i++)
    if (Av[i] != 0.) {
      if (Ai[i] + 1 > rn) rn = Ai[i] + 1;
      if (Aj[i] + 1 > cn) cn = Aj[i] + 1;
    }
  if (cn > rn) rn = cn; else cn = rn;   /* the reference indexes one flag array by row and column */
  ocsr *T = build_csr_dim(n, Ai, Aj, Av, rn, cn);
  double *keep = NEW(double, rn);
  for (int i = 0; i < rn; i++) keep[i] = (T->ro[i + 1] - T->ro[i] == 0) ? 0. : 1.;
  ocsr *A = amgo_sub_mat(T, keep, keep);
  free(keep); amgo_csr_free(T);
  return A;
}

/* mpm (:1684): X = alpha*A + beta*B; a coincident pair whose sum is exactly 0 is dropped,
   an entry present on one side only is always kept. */
ocsr *amgo_mpm(double alpha, const ocsr *A, double beta, const ocsr *B) {
  int rn = A->rn;
  int64_t cap = nnz_of(A) + nnz_of(B);
  ocsr *X = amgo_csr_new(rn, A->cn, cap);
  int64_t p = 0;
  for (int i = 0; i < rn; i++) {
    int ja = A->ro[i], ea = A->ro[i + 1], jb = B->ro[i], eb = B->ro[i + 1];
    while (ja < ea || jb < eb) {
      if (ja < ea && jb < eb && A->col[ja] == B->col[jb]) {
        double s = alpha * A->a[ja] + beta * B->a[jb];
        if (s != 0.) { X->col[p] = A->col[ja]; X->a[p] = s; p++; }
        ja++; jb++;
      } else if (jb == eb || (ja < ea && A->col[ja] < B->col[jb])) {
        X->col[p] = A->col[ja]; X->a[p] = alpha * A->a[ja]; p++; ja++;
      } else {
        X->col[p] = B->col[jb]; X->a[p] = beta * B->a[jb]; p++; jb++;
      }
    }
    X->ro[i + 1] = (int)p;
  }
  return X;
}

/* mxmpoint (:1807): X = A .* B on the intersection pattern, zeros kept */
ocsr *amgo_mxmpoint(const ocsr *A, const ocsr *B) {
  int rn = A->rn;
  int64_t cap = nnz_of(A) < nnz_of(B) ? nnz_of(A) : nnz_of(B);
  ocsr *X = amgo_csr_new(rn, A->cn, cap);
  int64_t p = 0;
  for (int i = 0; i < rn; i++) {
    int ja = A->ro[i], ea = A->ro[i + 1], jb = B->ro[i], eb = B->ro[i + 1];
    while (ja < ea && jb < eb) {
      if (A->col[ja] == B->col[jb]) {
        X->col[p] = A->col[ja]; X->a[p] = A->a[ja] * B->a[jb]; p++; ja++; jb++;
      } else if (A->col[ja] < B->col[jb]) ja++;
      else jb++;
    }
    X->ro[i + 1] = (int)p;
  }
  return X;
}

/* mxm (:1894): X = A*B.  The reference forms every (row of A) x (row of B^t) sparse dot
   product: X[i][c] = sum over k ascending of B[k][c]*A[i][k], and stores it iff != 0.
   Row-wise accumulation over k ascending gives every X[i][c] the same addition order. */
static int int_cmp(const void *a, const void *b) {
  int x = *(const int *)a, y = *(const int *)b;
  return x < y ? -1 : (x > y);
}
ocsr *amgo_spgemm(const ocsr *A, const ocsr *B) {
  int rn = A->rn, cn = B->cn;
  double *acc = NEW(double, cn);
  int *mark = NEW(int, cn), *list = NEW(int, cn);
  for (int c = 0; c < cn; c++) mark[c] = -1;
  int64_t cap = 1024, p = 0;
  ocsr *X = amgo_csr_new(rn, cn, 0);
  free(X->col); free(X->a);
  X->col = NEW(int, cap); X->a = NEW(double, cap);
  for (int i = 0; i < rn; i++) {
    int nl = 0;
    for (int ja = A->ro[i]; ja < A->ro[i + 1]; ja++) {
      int k = A->col[ja];
      double av = A->a[ja];
      for (int jb = B->ro[k]; jb < B->ro[k + 1]; jb++) {
        int c = B->col[jb];
        if (mark[c] != i) { mark[c] = i; acc[c] = 0.0; list[nl++] = c; }
        acc[c] += B->a[jb] * av;
      }
    }
    qsort(list, (size_t)nl, sizeof(int), int_cmp);
    if (p + nl > cap) {
      while (p + nl > cap) cap *= 2;
      X->col = (int *)realloc(X->col, sizeof(int) * (size_t)cap);
      X->a = (double *)realloc(X->a, sizeof(double) * (size_t)cap);
    }
    for (int q = 0; q < nl; q++) {
      int c = list[q];
      if (acc[c] != 0.0) { X->col[p] = c; X->a[p] = acc[c]; p++; }
    }
    X->ro[i + 1] = (int)p;
  }
  free(acc); free(mark); free(list);
  return X;
}

/* diag (:3363) */
static void diag_of(double *D, const ocsr *A) {
  for (int i = 0; i < A->rn; i++) {
    D[i] = 0.;
    for (int j = A->ro[i]; j < A->ro[i + 1]; j++) if (A->col[j] == i) { D[i] = A->a[j]; break; }
  }
}
/* diagcsr_op (:3389) */
static void scale_rows(ocsr *A, const double *D) {     /* dmult: A = D*A */
  for (int i = 0; i < A->rn; i++)
    for (int j = A->ro[i]; j < A->ro[i + 1]; j++) A->a[j] = A->a[j] * D[i];
}
static void scale_cols(ocsr *A, const double *D) {     /* multd: A = A*D */
  for (int i = 0; i < A->rn; i++)
    for (int j = A->ro[i]; j < A->ro[i + 1]; j++) A->a[j] = A->a[j] * D[A->col[j]];
}
static void sub_diag(ocsr *A, const double *D) {       /* dminus: A = A - D (first diag hit) */
  for (int i = 0; i < A->rn; i++)
    for (int j = A->ro[i]; j < A->ro[i + 1]; j++)
      if (A->col[j] == i) { A->a[j] = A->a[j] - D[i]; break; }
}
/* sum(s, A, 1) (:1193): column sums, entries visited in row-major order */
static void col_sums(double *s, const ocsr *A) {
  for (int c = 0; c < A->cn; c++) s[c] = 0.0;
  int64_t nnz = nnz_of(A);
  for (int64_t k = 0; k < nnz; k++) s[A->col[k]] += A->a[k];
}

/* ------------------------------------------------------------------------- */
/* coarsen (:2737) and mat_max (:3535)                                        */
/* ------------------------------------------------------------------------- */
static void mat_max(double *y, const ocsr *A, const double *f, const double *x, double tol) {
  for (int i = 0; i < A->cn; i++) y[i] = -DBL_MAX;
  for (int i = 0; i < A->rn; i++) {
    double xj = x[i], Amax = 0;
    for (int j = A->ro[i]; j < A->ro[i + 1]; j++)
      if (f[A->col[j]] != 0 && fabs(A->a[j]) > Amax) Amax = fabs(A->a[j]);
    Amax *= tol;
    for (int j = A->ro[i]; j < A->ro[i + 1]; j++) {
      int k = A->col[j];
      if (f[k] == 0 || fabs(A->a[j]) < Amax) continue;
      if (xj > y[k]) y[k] = xj;
    }
  }
}

int amgo_coarsen(double *vc, const ocsr *A, double ctol) {
  int n = A->cn, rounds = 0;
  double *D = NEW(double, n);
  diag_of(D, A);
  for (int i = 0; i < n; i++) D[i] = sqrt(D[i]);
  for (int i = 0; i < n; i++) D[i] = 1. / D[i];
  ocsr *S = csr_copy(A);
  scale_rows(S, D); scale_cols(S, D);
  for (int64_t k = 0; k < nnz_of(S); k++) S->a[k] = fabs(S->a[k]);
  diag_of(D, S); sub_diag(S, D);
  free(D);
  trace_csr("coarsen.S", S);

  double *vf = NEW(double, n), *g = NEW(double, n), *w1 = NEW(double, n), *w2 = NEW(double, n);
  double *tmp = NEW(double, n), *w = NEW(double, n), *mask = NEW(double, n), *m = NEW(double, n);
  int anyvc = 0;
  for (int i = 0; i < n; i++) { vc[i] = 0.; vf[i] = 1.; }
  for (;;) {
    rounds++;
    /* w1 = vf.*(S*(vf.*(S*vf))), w2 = vf.*(S*(vf.*(S*w1))) */
    apply_M(g, 0, vf, 1., S, vf);
    for (int i = 0; i < n; i++) g[i] = g[i] * vf[i];
    apply_M(w1, 0, g, 1., S, g);
    for (int i = 0; i < n; i++) w1[i] = w1[i] * vf[i];
    apply_M(w2, 0, w1, 1., S, w1);
    for (int i = 0; i < n; i++) w2[i] = w2[i] * vf[i];
    apply_M(tmp, 0, w2, 1., S, w2);
    for (int i = 0; i < n; i++) w2[i] = tmp[i] * vf[i];
    /* w = (1./w1).*w2, 0 where w1 == 0 */
    for (int i = 0; i < n; i++) { double inv = 1. / w1[i]; w[i] = inv * w2[i]; if (w1[i] == 0) w[i] = 0.; }
    int mi;
    double w1m = max_first(w1, n, &mi), wm = max_first(w, n, NULL);
    double b = (w1m < wm) ? sqrt(w1m) : sqrt(wm);
    if (b <= ctol) { if (!anyvc) vc[mi] = 1.; break; }
    for (int i = 0; i < n; i++) mask[i] = (w[i] > ctol * ctol) ? 1. : 0.;
    for (int i = 0; i < n; i++) tmp[i] = g[i] * mask[i];
    mat_max(m, S, vf, tmp, 0.1);
    for (int i = 0; i < n; i++) {
      double d = g[i] - m[i];
      mask[i] = (mask[i] != 0. && d >= 0.) ? 1. : 0.;
    }
    for (int i = 0; i < n; i++) tmp[i] = mask[i] * ((double)i + 1.0);
    mat_max(m, S, vf, tmp, 0.1);
    for (int i = 0; i < n; i++) {
      double d = ((double)i + 1.0) - m[i];
      mask[i] = (mask[i] != 0. && d > 0.) ? 1. : 0.;
    }
    for (int i = 0; i < n; i++) {
      if (mask[i] != 0.) { vc[i] = 1.; anyvc = 1; }
      vf[i] = ((vf[i] == 0.) != (mask[i] == 0.)) ? 1. : 0.;
    }
    if (g_trace_on) { char t[64]; snprintf(t, sizeof t, "coarsen.vc.r%d", rounds); trace(t, vc, sizeof(double) * (size_t)n); }
  }
  amgo_csr_free(S);
  free(vf); free(g); free(w1); free(w2); free(tmp); free(w); free(mask); free(m);
  return rounds;
}

/* ------------------------------------------------------------------------- */
/* pcg (:2242)                                                                */
/* ------------------------------------------------------------------------- */
int amgo_pcg(double *x, const ocsr *A, double *r, const double *M, double tol, const double *b,
             int mode) {
  int n = A->rn;
  double *p = NEW(double, n), *z = NEW(double, n), *w = NEW(double, n), *t = NEW(double, n);
  for (int i = 0; i < n; i++) { x[i] = 0.; p[i] = 0.; }
  for (int i = 0; i < n; i++) z[i] = M[i] * r[i];
  double rho = amgo_dot(r, z, n, mode);
  for (int i = 0; i < n; i++) t[i] = M[i] * b[i];
  double rho_0 = amgo_dot(t, b, n, mode);
  double rho_stop = tol * tol * rho_0;
  int nmax = n <= 100 ? n : 100, k = 0;
  double rho_old = 1;
  while (nmax > 0 && rho > rho_stop && k < nmax) {
    k++;
    double beta = rho / rho_old;
    for (int i = 0; i < n; i++) { double pb = p[i] * beta; p[i] = pb + z[i]; }
    apply_M(w, 0, NULL, 1, A, p);
    double alpha = amgo_dot(p, w, n, mode);
    alpha = rho / alpha;
    for (int i = 0; i < n; i++) { double pa = p[i] * alpha; x[i] = x[i] + pa; }
    for (int i = 0; i < n; i++) { double wa = w[i] * alpha; r[i] = r[i] - wa; }
    for (int i = 0; i < n; i++) z[i] = M[i] * r[i];
    rho_old = rho;
    rho = amgo_dot(r, z, n, mode);
  }
  free(p); free(z); free(w); free(t);
  return k;
}

/* ------------------------------------------------------------------------- */
/* chebsim (:2412)                                                            */
/* ------------------------------------------------------------------------- */
void amgo_chebsim(double *m, double *c, double rho, double tol) {
  double alpha = 0.25 * rho * rho, cp = 1, gamma = 1;
  *m = 1; *c = rho;
  while (*c > tol) {
    *m += 1;
    double d = alpha * (1 + gamma);
    gamma = d / (1 - d);
    double cn = (1 + gamma) * rho * (*c) - gamma * cp;
    cp = *c; *c = cn;
  }
}

/* ------------------------------------------------------------------------- */
/* tdeig / sec_root / rat_root / sum_3 (:2613-2726): eigenvalues of the arrowhead      */
/* matrix diag(d[1..n]) bordered by v[1..n], corner v[0]; y = last eigenvector entries */
/* ------------------------------------------------------------------------- */
#define TD_EPS (128 * DBL_EPSILON)
static double add3(double a, double b, double c) {
  if ((a >= 0 && b >= 0) || (a <= 0 && b <= 0)) return (a + b) + c;
  if ((a >= 0 && c >= 0) || (a <= 0 && c <= 0)) return (a + c) + b;
  return a + (b + c);
}
/* root of  -c/x + b + a x = 0  with the requested sign */
static double ratroot(double a, double b, double c, double sign) {
  double bh = (fabs(b) + sqrt(b * b + 4 * a * c)) / 2;
  return sign * (b * sign <= 0 ? bh / a : c / bh);
}
static double secular_root(double *y, const double *d, const double *v, int ri, int n) {
  const double dl = d[ri], dr = d[ri + 1], L = dr - dl;
  double xl = L / 2, xr = -L / 2;
  double tol = L;
  if (fabs(dl) > tol) tol = fabs(dl);
  if (fabs(dr) > tol) tol = fabs(dr);
  tol *= TD_EPS;
  for (;;) {
    if (fabs(xl) == 0 || xl < 0) { *y = 0; return dl; }
    if (fabs(xr) == 0 || xr > 0) { *y = 0; return dr; }
    double lam0 = fabs(xl) < fabs(xr) ? dl + xl : dr + xr;
    double al = 0, ar = 0, cl = 0, cr = 0, bln = 0, blp = 0, brn = 0, brp = 0, fn = 0, fp = 0;
    for (int i = 1; i <= ri; i++) {
      double den = (d[i] - dl) - xl;
      double fac = v[i] / den;
      double num = add3(d[i], -dr, -2 * xr);
      fn += v[i] * fac;
      fac *= fac;
      ar += fac;
      if (num > 0) brp += fac * num; else brn += fac * num;
      bln += fac * (d[i] - dl);
      cl += fac * xl * xl;
    }
    for (int i = ri + 1; i <= n; i++) {
      double den = (d[i] - dr) - xr;
      double fac = v[i] / den;
      double num = add3(d[i], -dl, -2 * xl);
      fp += v[i] * fac;
      fac *= fac;
      al += fac;
      if (num > 0) blp += fac * num; else bln += fac * num;
      brp += fac * (d[i] - dr);
      cr += fac * xr * xr;
    }
    if (lam0 > 0) fp += lam0; else fn += lam0;
    if (v[0] < 0) { fp -= v[0]; blp -= v[0]; brp -= v[0]; }
    else          { fn -= v[0]; bln -= v[0]; brn -= v[0]; }
    double lam;
    if (fp + fn > 0) {            /* root lies to the left */
      xl = ratroot(1 + al, add3(dl, blp, bln), cl, 1);
      lam = dl + xl; xr = xl - L;
    } else {                      /* to the right */
      xr = ratroot(1 + ar, add3(dr, brp, brn), cr, -1);
      lam = dr + xr; xl = xr + L;
    }
    if (fabs(lam - lam0) < tol) {
      double ty = 0, fac;
      for (int i = 1; i <= ri; i++) { fac = v[i] / ((d[i] - dl) - xl); ty += fac * fac; }
      for (int i = ri + 1; i <= n; i++) { fac = v[i] / ((d[i] - dr) - xr); ty += fac * fac; }
      *y = 1 / sqrt(1 + ty);
      return lam;
    }
  }
}
static void tdeig(double *lambda, double *y, double *d, const double *v, int n) {
  double v1 = 0, lo = v[0], hi = v[0];
  for (int i = 1; i <= n; i++) {
    double vi = fabs(v[i]), a = d[i] - vi, b = d[i] + vi;
    v1 += vi;
    if (a < lo) lo = a;
    if (b > hi) hi = b;
  }
  d[0] = v[0] - v1 < lo ? v[0] - v1 : lo;
  d[n + 1] = v[0] + v1 > hi ? v[0] + v1 : hi;
  for (int i = 0; i <= n; i++) lambda[i] = secular_root(&y[i], d, v, i, n);
}

/* ------------------------------------------------------------------------- */
/* lanczos (:2435).  lambda must hold 299 doubles.  Returns the number kept.  */
/* ------------------------------------------------------------------------- */
#define LANCZOS_KMAX 299
static int lanczos(double *lambda, const ocsr *A, int mode, amgo_rng *rng, int *iters) {
  int rn = A->rn, kmax = LANCZOS_KMAX;
  double *r = NEW(double, rn), *qk = NEW(double, rn), *qkm1 = NEW(double, rn), *Aqk = NEW(double, rn);
  double *l = lambda;
  double y[LANCZOS_KMAX + 1], d[LANCZOS_KMAX + 2], v[LANCZOS_KMAX + 1];
  for (int i = 0; i < rn; i++) r[i] = (double)amgo_rng_next(rng) / (double)2147483647;
  trace("lanczos.r0", r, sizeof(double) * (size_t)rn);
  double beta = norm2(r, rn, mode);
  double beta2 = beta * beta;
  beta = sqrt(beta2);
  int k = 0;
  double change = 0.0;
  /* || A - I ||_F over the stored entries (the diagonal entry is the first col==row hit) */
  {
    int64_t nnz = nnz_of(A);
    double *e = NEW(double, nnz);
    memcpy(e, A->a, sizeof(double) * (size_t)nnz);
    for (int i = 0; i < rn; i++)
      for (int j = A->ro[i]; j < A->ro[i + 1]; j++) if (A->col[j] == i) { e[j] = e[j] - 1.; break; }
    double fro = norm2(e, nnz, mode);
    free(e);
    double fro2 = fro * fro;
    fro = sqrt(fro2);
    if (fro < 1e-11) { l[0] = 1; l[1] = 1; y[0] = 0; y[1] = 0; k = 2; change = 0.0; }
    else change = 1.0;
  }
  if (rn == 1) { l[0] = A->a[0]; l[1] = A->a[0]; y[0] = 0; y[1] = 0; k = 2; change = 0.0; }
  for (int i = 0; i < rn; i++) qk[i] = 0.;
  while (k < kmax && (change > 1e-5 || y[0] > 1e-3 || y[k - 1] > 1e-3)) {
    k++;
    memcpy(qkm1, qk, sizeof(double) * (size_t)rn);
    double ib = 1. / beta;
    for (int i = 0; i < rn; i++) qk[i] = r[i] * ib;
    apply_M(Aqk, 0, NULL, 1, A, qk);
    double alpha = amgo_dot(qk, Aqk, rn, mode);
    for (int i = 0; i < rn; i++) {
      double aq = qk[i] * alpha, bq = qkm1[i] * beta;
      double t = Aqk[i] - aq;
      r[i] = t - bq;
    }
    if (k == 1) { l[0] = alpha; y[0] = 1; }
    else {
      double l0 = l[0], lkm2 = l[k - 2];
      d[0] = 0;
      for (int i = 1; i < k; i++) d[i] = l[i - 1];
      d[k] = 0;
      v[0] = alpha;
      for (int i = 1; i < k; i++) v[i] = beta * y[i - 1];
      tdeig(l, y, d, v, k - 1);
      change = fabs(l0 - l[0]) + fabs(lkm2 - l[k - 1]);
    }
    beta = norm2(r, rn, mode);
    beta2 = beta * beta;
    beta = sqrt(beta2);
    if (beta == 0) break;
  }
  if (iters) *iters = k;
  int n = 0;
  for (int i = 0; i < k; i++) if (y[i] < 0.01) lambda[n++] = l[i];
  free(r); free(qk); free(qkm1); free(Aqk);
  return n;
}
int amgo_lanczos(double *lambda, const ocsr *A, int mode, amgo_rng *rng, int *iters) {
  return lanczos(lambda, A, mode, rng, iters);
}

/* ------------------------------------------------------------------------- */
/* local energy-minimising solves: interp (:2053), interp_lmop (:1589)         */
/* ------------------------------------------------------------------------- */
/* sp_restrict_sorted (:2180): y[k] = x at index Ri[k] (0 if absent), both sorted */
static void restrict_sorted(double *y, int Rn, const int *Ri, int xn, const int *xi,
                            const double *x) {
  int p = 0;
  for (int k = 0; k < Rn; k++) {
    while (p < xn && xi[p] < Ri[k]) p++;
    y[k] = (p < xn && xi[p] == Ri[k]) ? x[p] : 0.0;
  }
}
/* mv_utt (:2122): y[i] = sum_{j<=i} U[i(i+1)/2 + j] x[j], left to right */
static void tri_t_mv(double *y, int n, const double *U, const double *x) {
  for (int i = 0; i < n; i++) {
    double v = 0;
    const double *u = U + (size_t)i * (i + 1) / 2;
    for (int j = 0; j <= i; j++) v += u[j] * x[j];
    y[i] = v;
  }
}
/* mv_ut (:2138): y[i] = sum_{j>=i} U[j(j+1)/2 + i] x[j], j ascending, starting from 0 */
static void tri_mv(double *y, int n, const double *U, const double *x) {
  for (int j = 0; j < n; j++) {
    y[j] = 0;
    const double *u = U + (size_t)j * (j + 1) / 2;
    for (int i = 0; i <= j; i++) y[i] += u[i] * x[j];
  }
}
/* The A-orthogonalisation shared by interp and interp_lmop: for the ordered support
   Qj[0..nz) builds packed upper-triangular Q with Q^t A(Qj,Qj) Q = I (:2081-2099).
   If QQt != NULL also accumulates QQt += q_k q_k^t after each column (:1637-1642). */
static void build_Q(double *Q, double *sqv1, double *sqv2, double *QQt, int nz, const int *Qj,
                    const ocsr *At) {
  double *qk = Q;
  if (QQt) for (int k = 0; k < nz * nz; k++) QQt[k] = 0;
  for (int k = 0; k < nz; k++, qk += k) {
    int s = Qj[k];
    restrict_sorted(sqv1, k + 1, Qj, At->ro[s + 1] - At->ro[s], &At->col[At->ro[s]], &At->a[At->ro[s]]);
    tri_t_mv(sqv2, k, Q, sqv1);
    tri_mv(qk, k, Q, sqv2);
    double alpha = sqv1[k];
    for (int m = 0; m < k; m++) alpha -= sqv1[m] * qk[m];
    alpha = -1.0 / sqrt(alpha);
    for (int m = 0; m < k; m++) qk[m] *= alpha;
    qk[k] = -alpha;
    if (QQt)
      for (int m = 0; m <= k; m++) {
        double qkm = qk[m];
        for (int j = 0; j <= k; j++) QQt[m * nz + j] += qkm * qk[j];
      }
  }
}
static int max_row_len(const ocsr *A) {
  int mx = 0;
  for (int i = 0; i < A->rn; i++) { int l = A->ro[i + 1] - A->ro[i]; if (l > mx) mx = l; }
  return mx;
}
/* interp (:2053): for every row i of Wt (a coarse point and its F support), overwrite the
   stored values with  Q Q^t R (B e_i + u_i lambda) */
static void interp(ocsr *Wt, const ocsr *At, const ocsr *Bt, const double *u, const double *lambda) {
  int mx = max_row_len(Wt);
  double *sqv1 = NEW(double, 2 * mx + (size_t)mx * (mx + 1) / 2 + 1);
  double *sqv2 = sqv1 + mx, *Q = sqv2 + mx;
  for (int i = 0; i < Wt->rn; i++) {
    int wir = Wt->ro[i], nz = Wt->ro[i + 1] - wir;
    const int *Qj = &Wt->col[wir];
    build_Q(Q, sqv1, sqv2, NULL, nz, Qj, At);
    restrict_sorted(sqv1, nz, Qj, Bt->ro[i + 1] - Bt->ro[i], &Bt->col[Bt->ro[i]], &Bt->a[Bt->ro[i]]);
    for (int k = 0; k < nz; k++) sqv1[k] += u[i] * lambda[Qj[k]];
    tri_t_mv(sqv2, nz, Q, sqv1);
    tri_mv(&Wt->a[wir], nz, Q, sqv2);
  }
  free(sqv1);
}
/* sp_add (:1665): y(yi) += alpha*x(xi).  The reference ASSUMES xi is a subset of yi and does
   not check: for every x it walks forward to the first stored index >= xi and adds there.
   When the assumption fails (it does whenever min_skel left zero-valued entries in column 0,
   :2229-2233) the update lands on a wrong entry, possibly in a later row, possibly past the
   end of the arrays.  Two modes:
     CHECKED  (product semantics)  an absent index is skipped; nothing else is touched.
     REFSCAN  (reference semantics) the same unchecked walk over the whole col/a arrays; a
              walk that would leave the arrays is reported (the reference is in UB there). */
enum { SPADD_CHECKED = 0, SPADD_REFSCAN = 1 };
static int g_spadd_mode = SPADD_CHECKED;
static int64_t g_spadd_miss = 0, g_spadd_ub = 0;
void amgo_set_spadd_mode(int m) { g_spadd_mode = m; }
int64_t amgo_debug_spadd_miss(void) { return g_spadd_miss; }
int64_t amgo_debug_spadd_ub(void) { return g_spadd_ub; }
void amgo_debug_reset(void) { g_spadd_miss = 0; g_spadd_ub = 0; }
/* yi/y point at row start inside arrays that extend `avail` entries from there */
static void sp_add(int yn, int64_t avail, const int *yi, double *y, double alpha, int xn,
                   const int *xi, const double *x) {
  if (yn == 0) return;
  if (g_spadd_mode == SPADD_CHECKED) {
    int p = 0;
    for (int k = 0; k < xn; k++) {
      while (p < yn && yi[p] < xi[k]) p++;
      if (p < yn && yi[p] == xi[k]) { y[p] += alpha * x[k]; p++; }
      else g_spadd_miss++;
    }
    return;
  }
  int64_t p = 0;
  for (int k = 0; k < xn; k++) {
    for (;;) {
      if (p >= avail) { g_spadd_ub++; return; }
      if (yi[p] >= xi[k]) break;
      p++;
    }
    if (yi[p] != xi[k] || p >= yn) g_spadd_miss++;
    y[p] += alpha * x[k];
    p++;
  }
}
/* interp_lmop (:1589): S = sum_i u_i * (Q_i Q_i^t) scattered to the pattern of S */
static void interp_lmop(ocsr *St, const ocsr *At, const double *u, const ocsr *Wskt) {
  int mx = max_row_len(Wskt);
  double *sqv1 = NEW(double, 2 * mx + (size_t)mx * (mx + 1) / 2 + (size_t)mx * mx + 1);
  double *sqv2 = sqv1 + mx, *Q = sqv2 + mx, *QQt = Q + (size_t)mx * (mx + 1) / 2;
  for (int64_t k = 0; k < nnz_of(St); k++) St->a[k] = 0.0;
  for (int i = 0; i < Wskt->rn; i++) {
    const int *Qj = &Wskt->col[Wskt->ro[i]];
    int nz = Wskt->ro[i + 1] - Wskt->ro[i];
    build_Q(Q, sqv1, sqv2, QQt, nz, Qj, At);
    for (int k = 0; k < nz; k++) {
      int j = Qj[k], tj = St->ro[j];
      sp_add(St->ro[j + 1] - tj, nnz_of(St) - tj, &St->col[tj], &St->a[tj], u[i], nz, Qj, QQt + (size_t)k * nz);
    }
  }
  free(sqv1);
}

/* min_skel (:2198): one entry per row at the first largest value; 1 if positive else 0 */
static ocsr *min_skel(const ocsr *R) {
  ocsr *W = amgo_csr_new(R->rn, R->cn, R->rn);
  for (int i = 0; i < R->rn; i++) {
    double ymax = -DBL_MAX; int j = 0;
    for (int k = R->ro[i]; k < R->ro[i + 1]; k++) if (R->a[k] > ymax) { ymax = R->a[k]; j = R->col[k]; }
    W->a[i] = (ymax > 0.0) ? 1.0 : 0.0;
    W->col[i] = j;
    W->ro[i] = i;
  }
  W->ro[R->rn] = R->rn;
  return W;
}

/* solve_constraint (:1499) */
static void solve_constraint(double *lam, const ocsr *Wsk, const ocsr *Wskt, const ocsr *Af,
                             const ocsr *W0, const double *alpha, const double *u, const double *v,
                             double tol, int mode) {
  int nf = Wsk->rn, nc = Wsk->cn;
  double *au2 = NEW(double, nc);
  for (int i = 0; i < nc; i++) { double uu = u[i] * u[i]; au2[i] = uu * alpha[i]; }
  ocsr *S = amgo_spgemm(Wsk, Wskt);           /* pattern of W_skel*W_skel' (nonzero products) */
  interp_lmop(S, Af, au2, Wskt);
  trace_csr("sc.S", S);
  double *resid = NEW(double, nf), *d = NEW(double, nf), *keep = NEW(double, nf);
  apply_M(resid, 1.0, v, -1.0, W0, u);
  diag_of(d, S);
  int all = 1;
  for (int i = 0; i < nf; i++) { keep[i] = (d[i] != 0.) ? 1. : 0.; if (keep[i] == 0.) { all = 0; lam[i] = 0.; } }
  if (!all) { ocsr *S2 = amgo_sub_mat(S, keep, keep); amgo_csr_free(S); S = S2; }
  int nk = 0;
  double *rc = NEW(double, nf), *dc = NEW(double, nf), *lc = NEW(double, nf);
  for (int i = 0; i < nf; i++) if (keep[i] != 0.) { rc[nk] = resid[i]; dc[nk] = d[i]; lc[nk] = lam[i]; nk++; }
  for (int i = nk; i < nf; i++) { rc[i] = 0.; dc[i] = 0.; lc[i] = 0.; }
  double *q = NEW(double, nk), *x = NEW(double, nk);
  apply_M(q, 1., rc, -1., S, lc);
  for (int i = 0; i < nk; i++) dc[i] = 1. / dc[i];
  amgo_pcg(x, S, q, dc, tol, rc, mode);
  nk = 0;
  for (int i = 0; i < nf; i++) if (keep[i] != 0.) lam[i] += x[nk++];
  amgo_csr_free(S);
  free(au2); free(resid); free(d); free(keep); free(rc); free(dc); free(lc); free(q); free(x);
}

/* solve_weights (:1437): returns W and W0 (both nf x nc) */
static void solve_weights(ocsr **W, ocsr **W0, double *lam, const ocsr *Wsk, const ocsr *Af,
                          const ocsr *Ar, const double *alpha, const double *u, const double *v,
                          double tol, int mode) {
  int nf = Af->rn, nc = Wsk->cn;
  double *au = NEW(double, nc), *zeros = NEW(double, nf);
  for (int i = 0; i < nc; i++) au[i] = alpha[i] * u[i];
  for (int i = 0; i < nf; i++) zeros[i] = 0.0;
  ocsr *W0t = amgo_transpose(Wsk);
  ocsr *Armt = amgo_transpose(Ar);
  for (int64_t k = 0; k < nnz_of(Armt); k++) Armt->a[k] = Armt->a[k] * -1.0;
  interp(W0t, Af, Armt, au, zeros);
  *W0 = amgo_transpose(W0t);
  amgo_csr_free(W0t);
  ocsr *Wskt = amgo_transpose(Wsk);
  solve_constraint(lam, Wsk, Wskt, Af, *W0, alpha, u, v, tol, mode);
  trace("sw.lam", lam, sizeof(double) * (size_t)nf);
  interp(Wskt, Af, Armt, au, lam);
  *W = amgo_transpose(Wskt);
  amgo_csr_free(Wskt); amgo_csr_free(Armt);
  free(au); free(zeros);
}

/* find_support (:1260) */
static ocsr *find_support(const ocsr *R, double goal) {
  int nf = R->rn, nc = R->cn;
  int64_t nnz = nnz_of(R), nskel = 0;
  int *ski = NEW(int, nnz), *skj = NEW(int, nnz);
  double theta = 0.5;
  double *rs = NEW(double, nf), *tmp = NEW(double, nf), *onec = NEW(double, nc);
  double *w = NEW(double, nc), *w2 = NEW(double, nc), *v = NEW(double, nc), *sumR = NEW(double, nc);
  double *maxx = NEW(double, nc);
  int *bad = NEW(int, nc), *maski = NEW(int, nc);
  for (int i = 0; i < nc; i++) onec[i] = 1.;
  ocsr *Rl = csr_copy(R);
  for (;;) {
    apply_M(rs, 0., NULL, 1., Rl, onec);
    apply_Mt(w, Rl, rs);
    apply_M(tmp, 0., NULL, 1., Rl, w);
    apply_Mt(w2, Rl, tmp);
    for (int i = 0; i < nc; i++) { v[i] = w2[i] / w[i]; if (w[i] == 0.) v[i] = 0.; }
    double mv = max_first(v, nc, NULL);
    VLOG("    find_support: nnz(R)=%ld max(v)=%.17g goal=%.17g theta=%g\n", (long)nnz_of(Rl), mv, goal, theta);
    if (mv < goal) break;
    if (nf <= 1) break;   /* the reference writes row index 1 here and never terminates */
    while (mv <= (1 + theta) * goal) theta = theta / 2.;
    col_sums(sumR, Rl);
    int nbad = 0;
    for (int i = 0; i < nc; i++) {
      bad[i] = (w[i] > (1 + theta) * goal && sumR[i] != 0.);
      if (bad[i]) { nbad++; maxx[i] = -DBL_MAX; maski[i] = (nf > 1) ? 0 : 1; }
    }
    /* per bad column: first row (ascending) with the largest R_ij * rs_i */
    if (nf > 1)
      for (int i = 0; i < nf; i++)
        for (int j = Rl->ro[i]; j < Rl->ro[i + 1]; j++) {
          int c = Rl->col[j];
          if (!bad[c]) continue;
          double x = Rl->a[j] * rs[i];
          if (x > maxx[c]) { maxx[c] = x; maski[c] = i; }
        }
    /* R = R - R.*M: the selected entries cancel exactly and leave the pattern */
    ocsr *Rn = amgo_csr_new(nf, nc, nnz_of(Rl));
    int64_t p = 0;
    for (int i = 0; i < nf; i++) {
      for (int j = Rl->ro[i]; j < Rl->ro[i + 1]; j++) {
        int c = Rl->col[j];
        if (bad[c] && maski[c] == i) {
          double s = 1. * Rl->a[j] + (-1.) * (Rl->a[j] * 1.);
          if (s == 0.) continue;
          Rn->col[p] = c; Rn->a[p] = s; p++;
        } else { Rn->col[p] = c; Rn->a[p] = 1. * Rl->a[j]; p++; }
      }
      Rn->ro[i + 1] = (int)p;
    }
    amgo_csr_free(Rl); Rl = Rn;
    for (int c = 0; c < nc; c++) if (bad[c]) { ski[nskel] = maski[c]; skj[nskel] = c; nskel++; }
    (void)nbad;
  }
  amgo_csr_free(Rl);
  double *one = NEW(double, nskel);
  for (int64_t i = 0; i < nskel; i++) one[i] = 1.;
  ocsr *Sk = build_csr_dim(nskel, ski, skj, one, nf, nc);
  free(one); free(ski); free(skj); free(rs); free(tmp); free(onec); free(w); free(w2); free(v);
  free(sumR); free(maxx); free(bad); free(maski);
  return Sk;
}

/* expand_support (:907).  The reference ranks |X| of every bad row through dense
   rank x row matrices; per row that is: sort descending (stable, ties keep column order),
   take the shortest prefix whose running sum is not below half the row sum. */
typedef struct { int col; double v; } colval;
static void sort_desc_stable(colval *e, int n) {   /* insertion sort == stable mergesort order */
  for (int i = 1; i < n; i++) {
    colval t = e[i]; int j = i - 1;
    while (j >= 0 && e[j].v < t.v) { e[j + 1] = e[j]; j--; }
    e[j + 1] = t;
  }
}
static ocsr *expand_support(const ocsr *Wsk, const ocsr *R, const ocsr *R0, double gamma) {
  int nf = Wsk->rn, nc = Wsk->cn;
  ocsr *M = find_support(R, gamma);
  trace_csr("es.M", M);
  ocsr *ns = amgo_mpm(1., M, 1., Wsk);
  amgo_csr_free(M);
  int nbad = 0;
  char *badrow = NEW(char, nf);
  for (int i = 0; i < nf; i++) {
    badrow[i] = 0;
    for (int j = ns->ro[i]; j < ns->ro[i + 1]; j++) if (ns->a[j] == 2.) { badrow[i] = 1; nbad++; break; }
  }
  if (nbad == 0) {
    for (int64_t k = 0; k < nnz_of(ns); k++) if (ns->a[k] == 2.) ns->a[k] = 1.;
    free(badrow);
    return ns;
  }
  /* X = R0 - R0.*W_skel */
  ocsr *R0W = amgo_mxmpoint(R0, Wsk);
  ocsr *X = amgo_mpm(1., R0, -1., R0W);
  amgo_csr_free(R0W);
  int mx = max_row_len(X);
  colval *e = NEW(colval, mx + 1);
  int64_t cap = nnz_of(X), nn = 0;
  int *ni = NEW(int, cap + 1), *nj = NEW(int, cap + 1);
  double *nv = NEW(double, cap + 1);
  for (int i = 0; i < nf; i++) {
    if (!badrow[i]) continue;
    int len = X->ro[i + 1] - X->ro[i];
    for (int k = 0; k < len; k++) { e[k].col = X->col[X->ro[i] + k]; e[k].v = fabs(X->a[X->ro[i] + k]); }
    sort_desc_stable(e, len);
    double tot = 0.0;                       /* sum(X,1): ranks ascending; zeros are absent */
    for (int k = 0; k < len; k++) if (e[k].v != 0.) tot += e[k].v;
    double half = tot * 0.5;
    int below = 0;
    double cs = 0.0;
    for (int k = 0; k < len; k++) {
      if (e[k].v != 0.) cs += e[k].v;
      /* mpm(SV,1,S,-1,V): V has no entry when half == 0 */
      double s = (half != 0.) ? (1. * cs + (-1.) * half) : (1. * cs);
      if (s < 0.) below++;
    }
    int take = below + 1; if (take > len) take = len;
    for (int k = 0; k < take; k++) { ni[nn] = i; nj[nn] = e[k].col; nv[nn] = 1.; nn++; }
  }
  amgo_csr_free(X);
  ocsr *N = build_csr_dim(nn, ni, nj, nv, nf, nc);
  ocsr *out = amgo_mpm(1., ns, 1., N);
  for (int64_t k = 0; k < nnz_of(out); k++) if (out->a[k] != 0.) out->a[k] = 1.;
  amgo_csr_free(N); amgo_csr_free(ns);
  free(e); free(ni); free(nj); free(nv); free(badrow);
  return out;
}

#define AMGO_MAX_INTERP_ROUNDS 100
/* interpolation (:598) */
ocsr *amgo_interpolation(const ocsr *Af, const ocsr *Ac, const ocsr *Ar, double gamma2, double tol,
                         int mode, int *rounds_out) {
  int nf = Af->rn, nc = Ac->cn;
  double *Df = NEW(double, nf), *Dfinv = NEW(double, nf), *uc = NEW(double, nc);
  double *tmp = NEW(double, nf), *v = NEW(double, nf), *b = NEW(double, nf);
  diag_of(Df, Af);
  for (int i = 0; i < nf; i++) Dfinv[i] = 1. / Df[i];
  for (int i = 0; i < nc; i++) uc[i] = 1.;
  apply_M(tmp, 0, NULL, -1, Ar, uc);
  for (int i = 0; i < nf; i++) b[i] = 1.0;
  amgo_pcg(v, Af, tmp, Df, 1e-16, b, mode);
  trace("ip.v", v, sizeof(double) * (size_t)nf);
  double *Dc = NEW(double, nc), *Dcinv = NEW(double, nc);
  diag_of(Dc, Ac);
  for (int i = 0; i < nc; i++) Dcinv[i] = 1. / Dc[i];
  /* W_skel = min_skel((Ar/Dc).*(Df\Ar)) */
  ocsr *ArD = csr_copy(Ar);
  for (int64_t k = 0; k < nnz_of(ArD); k++) ArD->a[k] = ArD->a[k] * ArD->a[k];
  scale_rows(ArD, Dfinv); scale_cols(ArD, Dcinv);
  ocsr *Wsk = min_skel(ArD);
  amgo_csr_free(ArD);
  double *lam = NEW(double, nf), *alpha = NEW(double, nc);
  for (int i = 0; i < nf; i++) lam[i] = 0.;
  memcpy(alpha, Dc, sizeof(double) * (size_t)nc);
  double *Dfsqrti = Dfinv;
  for (int i = 0; i < nf; i++) Dfsqrti[i] = sqrt(Dfsqrti[i]);
  double *Dcsqrti = NEW(double, nc), *w1 = NEW(double, nc), *w2 = NEW(double, nc), *ones = NEW(double, nc);
  double *r = NEW(double, nc);
  for (int i = 0; i < nc; i++) ones[i] = 1.0;
  ocsr *W = NULL;
  int rounds = 0;
  for (;;) {
    rounds++;
    if (rounds > AMGO_MAX_INTERP_ROUNDS) {   /* the reference has no bound and would spin forever */
      fprintf(stderr, "amg_oracle: interpolation did not converge in %d rounds\n", AMGO_MAX_INTERP_ROUNDS);
      amgo_csr_free(Wsk); Wsk = NULL; W = NULL; break;
    }
    char pfx_save[32]; memcpy(pfx_save, g_trace_prefix, sizeof pfx_save);
    if (g_trace_on) { char t[32]; snprintf(t, sizeof t, "%.20sr%d.", pfx_save, rounds); memcpy(g_trace_prefix, t, sizeof t); }
    trace_csr("ip.Wsk", Wsk);
    ocsr *Wt, *W0;
    solve_weights(&Wt, &W0, lam, Wsk, Af, Ar, alpha, uc, v, tol, mode);
    trace_csr("ip.W0", W0); trace_csr("ip.Wtmp", Wt);
    ocsr *AfW = amgo_spgemm(Af, W0);
    ocsr *Arhat0 = amgo_mpm(1., AfW, 1., Ar);
    amgo_csr_free(AfW);
    AfW = amgo_spgemm(Af, Wt);
    ocsr *Arhat = amgo_mpm(1., AfW, 1., Ar);
    amgo_csr_free(AfW);
    /* dchat = sum(W.*(Arhat+Ar),1)' + diag(Ac); Dcsqrti = 1/sqrt(dchat) */
    ocsr *Arr = amgo_mpm(1.0, Arhat, 1.0, Ar);
    ocsr *ArW = amgo_mxmpoint(Wt, Arr);
    amgo_csr_free(Arr);
    col_sums(Dcsqrti, ArW);
    amgo_csr_free(ArW);
    for (int i = 0; i < nc; i++) Dcsqrti[i] = Dcsqrti[i] + Dc[i];
    for (int i = 0; i < nc; i++) Dcsqrti[i] = 1. / Dcsqrti[i];
    for (int i = 0; i < nc; i++) Dcsqrti[i] = sqrt(Dcsqrti[i]);
    /* R = abs(Dfsqrti*Arhat)*Dcsqrti, R0 likewise from Arhat0 */
    ocsr *R = Arhat, *R0 = Arhat0;
    scale_rows(R, Dfsqrti);
    for (int64_t k = 0; k < nnz_of(R); k++) R->a[k] = fabs(R->a[k]);
    scale_cols(R, Dcsqrti);
    scale_rows(R0, Dfsqrti);
    for (int64_t k = 0; k < nnz_of(R0); k++) R0->a[k] = fabs(R0->a[k]);
    scale_cols(R0, Dcsqrti);
    trace_csr("ip.R", R); trace_csr("ip.R0", R0);
    apply_M(tmp, 0., NULL, 1., R, ones);
    apply_Mt(w1, R, tmp);
    apply_M(tmp, 0., NULL, 1., R, w1);
    apply_Mt(w2, R, tmp);
    for (int i = 0; i < nc; i++) { r[i] = w2[i] / w1[i]; if (w1[i] == 0) r[i] = 0.; }
    int nbig = 0;
    for (int i = 0; i < nc; i++) if (r[i] > gamma2) nbig++;
    double w1m = max_first(w1, nc, NULL);
    VLOG("  interp round %d: nnz(Wsk)=%ld, %d cols > gamma2, max(w1)=%g\n", rounds, (long)nnz_of(Wsk), nbig, w1m);
    if (nbig == 0 || w1m <= gamma2) {
      amgo_csr_free(W0);
      solve_weights(&W, &W0, lam, Wsk, Af, Ar, alpha, uc, v, 1e-16, mode);
      double *wuc = NEW(double, nf);
      apply_M(wuc, 0., NULL, 1., W, uc);
      /* the reference rescales only entries whose column index equals the row index (:821-837) */
      for (int i = 0; i < nf; i++)
        if (wuc[i] != 0.)
          for (int j = W->ro[i]; j < W->ro[i + 1]; j++)
            if (i == W->col[j]) { double s = v[i] / wuc[i]; W->a[j] = s * W->a[j]; }
      free(wuc);
      amgo_csr_free(Wt); amgo_csr_free(W0); amgo_csr_free(R); amgo_csr_free(R0);
      memcpy(g_trace_prefix, pfx_save, sizeof pfx_save);
      break;
    }
    for (int i = 0; i < nc; i++) { double x = w2[i] > 1e-6 ? w2[i] : 1e-6; alpha[i] = Dc[i] / x; }
    ocsr *nsk = expand_support(Wsk, R, R0, gamma2);
    amgo_csr_free(Wsk); Wsk = nsk;
    amgo_csr_free(Wt); amgo_csr_free(W0); amgo_csr_free(R); amgo_csr_free(R0);
    memcpy(g_trace_prefix, pfx_save, sizeof pfx_save);
  }
  amgo_csr_free(Wsk);
  free(Df); free(Dfinv); free(uc); free(tmp); free(v); free(b); free(Dc); free(Dcinv);
  free(lam); free(alpha); free(Dcsqrti); free(w1); free(w2); free(ones); free(r);
  if (rounds_out) *rounds_out = rounds;
  return W;
}

/* ------------------------------------------------------------------------- */
/* amg_setup (:60)                                                            */
/* ------------------------------------------------------------------------- */
typedef struct {
  ocsr *A, *Af, *W, *AfP;
  double *C, *D;
  int *idc, *idf;
  int n, nf, nc;
  double m, rho, lmin, lmax;
  int coarsen_rounds, lanczos_k, interp_rounds;
} olevel;

struct amgo_hier {
  int nlevels, nullspace, cap;
  olevel *lv;
  int n0;
};

void amgo_free(amgo_hier *h) {
  if (!h) return;
  for (int l = 0; l < h->nlevels; l++) {
    olevel *L = &h->lv[l];
    amgo_csr_free(L->A); amgo_csr_free(L->Af); amgo_csr_free(L->W); amgo_csr_free(L->AfP);
    free(L->C); free(L->D); free(L->idc); free(L->idf);
  }
  free(h->lv); free(h);
}

int amgo_setup(int64_t nnz, const int32_t *Ai, const int32_t *Aj, const double *Av, int mode,
               amgo_hier **out) {
  ocsr *A = amgo_build_csr(nnz, Ai, Aj, Av);
  const double tol = 0.5, ctol = 0.7, itol = 1e-4;
  const double gamma2 = 1. - sqrt(1. - tol);
  amgo_hier *h = NEW(amgo_hier, 1);
  h->cap = 128; h->nlevels = 0; h->nullspace = 0; h->n0 = A->rn;
  h->lv = NEW(olevel, h->cap);
  memset(h->lv, 0, sizeof(olevel) * (size_t)h->cap);
  amgo_rng rng; amgo_rng_seed(&rng, 1);
  int level = 0;
  int *id0 = NEW(int, A->rn), *idl = id0;
  for (int k = 0; k < A->rn; k++) idl[k] = k + 1;
  for (;;) {
    if (level >= h->cap) { fprintf(stderr, "amg_oracle: more than %d levels\n", h->cap); return -1; }
    olevel *L = &h->lv[level];
    int rn = A->rn;
    L->A = csr_copy(A); L->n = rn;
    h->nlevels = level + 1;
    if (g_trace_on) snprintf(g_trace_prefix, sizeof g_trace_prefix, "L%d.", level);
    trace_csr("A", A);
    /* amg_setup.c:166 tests A->a[0] < 1e-9; with an EMPTY 1x1 last level (the only entry
       cancelled exactly and mpm dropped it) the reference reads an unset value there.  Checker and
       product both treat that case as a null space. */
    if (rn <= 1) { h->nullspace = (rn == 1 && (A->ro[1] == 0 || A->a[0] < 1e-9)) ? 1 : 0; break; }

    /* coarsen (:174-186) */
    double *vc = NEW(double, rn), *vf = NEW(double, rn);
    VLOG("level %d: n=%d nnz=%ld\n", level, rn, (long)nnz_of(A));
    L->coarsen_rounds = amgo_coarsen(vc, A, ctol);
    VLOG("  coarsen: %d rounds\n", L->coarsen_rounds);
    for (int i = 0; i < rn; i++) vf[i] = (vc[i] == 0.) ? 1. : 0.;
    L->C = vc;
    trace("vc", vc, sizeof(double) * (size_t)rn);

    /* diagonal smoother (:188-228) */
    ocsr *Af = amgo_sub_mat(A, vf, vf);
    int nf = Af->rn;
    double *D = NEW(double, nf);
    for (int i = 0; i < nf; i++) {
      double s = 0;
      for (int j = Af->ro[i]; j < Af->ro[i + 1]; j++) s += Af->a[j] * Af->a[j];
      D[i] = 1. / s;
    }
    {
      double *dg = NEW(double, nf);
      diag_of(dg, Af);
      for (int i = 0; i < nf; i++) D[i] = dg[i] * D[i];
      free(dg);
    }
    if (nf >= 2) {   /* (:232-285) */
      double *Dh = NEW(double, nf);
      for (int i = 0; i < nf; i++) Dh[i] = sqrt(D[i]);
      ocsr *DAD = csr_copy(Af);
      scale_rows(DAD, Dh); scale_cols(DAD, Dh);
      double lambda[LANCZOS_KMAX + 1];
      int k = lanczos(lambda, DAD, mode, &rng, &L->lanczos_k);
      VLOG("  lanczos: %d iterations, %d kept, [%g, %g]\n", L->lanczos_k, k, lambda[0], lambda[k - 1]);
      double a = lambda[0], b = lambda[k - 1];
      double sc = 2. / (a + b);
      for (int i = 0; i < nf; i++) D[i] = D[i] * sc;
      L->rho = (b - a) / (b + a);
      L->lmin = a; L->lmax = b;
      double m, c;
      amgo_chebsim(&m, &c, L->rho, gamma2);
      L->m = m;
      free(Dh); amgo_csr_free(DAD);
    } else { L->rho = 0; L->m = 1; L->lmin = L->lmax = 0; }
    L->D = D; L->Af = Af; L->nf = nf;
    trace("D", D, sizeof(double) * (size_t)nf);

    /* interpolation (:302-334) */
    ocsr *Afc = amgo_sub_mat(A, vf, vc);
    ocsr *Ac = amgo_sub_mat(A, vc, vc);
    int nc = Ac->rn;
    L->nc = nc;
    L->idc = NEW(int, nc); L->idf = NEW(int, nf);
    { int cc = 0, cf = 0;
      for (int i = 0; i < rn; i++) { if (vc[i] == 1.) L->idc[cc++] = idl[i]; else L->idf[cf++] = idl[i]; } }
    ocsr *W = amgo_interpolation(Af, Ac, Afc, gamma2, itol, mode, &L->interp_rounds);
    if (!W) { amgo_free(h); *out = NULL; return -2; }
    L->W = W;
    trace_csr("W", W);

    /* Galerkin product (:336-372): AfP = Af*W + Afc;  A = W'*AfP + Acf*W + Ac */
    ocsr *AfW = amgo_spgemm(Af, W);
    ocsr *AfP = amgo_mpm(1., AfW, 1., Afc);
    amgo_csr_free(AfW);
    L->AfP = AfP;
    trace_csr("AfP", AfP);
    ocsr *Wt = amgo_transpose(W);
    ocsr *WtAfP = amgo_spgemm(Wt, AfP);
    ocsr *Acf = amgo_transpose(Afc);
    ocsr *AcfW = amgo_spgemm(Acf, W);
    ocsr *Atmp = amgo_mpm(1., WtAfP, 1., AcfW);
    ocsr *An = amgo_mpm(1., Atmp, 1, Ac);
    amgo_csr_free(Wt); amgo_csr_free(WtAfP); amgo_csr_free(Acf); amgo_csr_free(AcfW);
    amgo_csr_free(Atmp); amgo_csr_free(Afc); amgo_csr_free(Ac);
    amgo_csr_free(A);
    A = An;
    free(vf);
    idl = L->idc;
    level++;
  }
  free(id0);
  amgo_csr_free(A);
  g_trace_prefix[0] = 0;
  *out = h;
  return 0;
}

int amgo_nlevels(const amgo_hier *h) { return h->nlevels; }
int amgo_nullspace(const amgo_hier *h) { return h->nullspace; }
int amgo_level_info(const amgo_hier *h, int l, int64_t info[10]) {
  if (l < 0 || l >= h->nlevels) return -1;
  const olevel *L = &h->lv[l];
  info[0] = L->n; info[1] = nnz_of(L->A); info[2] = L->nf; info[3] = L->nc;
  info[4] = L->Af ? nnz_of(L->Af) : 0; info[5] = L->W ? nnz_of(L->W) : 0;
  info[6] = L->AfP ? nnz_of(L->AfP) : 0; info[7] = L->coarsen_rounds; info[8] = L->lanczos_k;
  info[9] = L->interp_rounds;
  return 0;
}
int amgo_level_params(const amgo_hier *h, int l, double par[4]) {
  if (l < 0 || l >= h->nlevels) return -1;
  const olevel *L = &h->lv[l];
  par[0] = L->m; par[1] = L->rho; par[2] = L->lmin; par[3] = L->lmax;
  return 0;
}
int amgo_get_csr(const amgo_hier *h, int l, int which, int *rn, int *cn, int64_t *nnz, int32_t *ro,
                 int32_t *col, double *a) {
  if (l < 0 || l >= h->nlevels) return -1;
  const olevel *L = &h->lv[l];
  const ocsr *M = which == 0 ? L->A : which == 1 ? L->Af : which == 2 ? L->W : which == 3 ? L->AfP : NULL;
  if (!M) return -2;
  if (rn) *rn = M->rn;
  if (cn) *cn = M->cn;
  if (nnz) *nnz = nnz_of(M);
  if (ro) memcpy(ro, M->ro, sizeof(int) * (size_t)(M->rn + 1));
  if (col) memcpy(col, M->col, sizeof(int) * (size_t)nnz_of(M));
  if (a) memcpy(a, M->a, sizeof(double) * (size_t)nnz_of(M));
  return 0;
}
int amgo_get_vec(const amgo_hier *h, int l, int which, double *out) {
  if (l < 0 || l >= h->nlevels - 1) return -1;
  const olevel *L = &h->lv[l];
  switch (which) {
    case 0: memcpy(out, L->C, sizeof(double) * (size_t)L->n); return 0;
    case 1: memcpy(out, L->D, sizeof(double) * (size_t)L->nf); return 0;
    case 2: for (int i = 0; i < L->nc; i++) out[i] = L->idc[i]; return 0;
    case 3: for (int i = 0; i < L->nf; i++) out[i] = L->idf[i]; return 0;
  }
  return -2;
}

/* ------------------------------------------------------------------------- */
/* amg_export (:405), savemats (:483), savevec (:550)                         */
/* ------------------------------------------------------------------------- */
static int save_mats(int *len, int n, int nl, const int *lvl, const amgo_hier *h, int which,
                     const char *path) {
  const double magic = 3.14159;
  FILE *f = fopen(path, "wb");
  if (!f) return -1;
  fwrite(&magic, sizeof(double), 1, f);
  int *row = NEW(int, nl + 1);
  for (int i = 0; i < nl; i++) row[i] = 0;
  int mx = 1;
  for (int i = 0; i < nl; i++) {
    const olevel *L = &h->lv[i];
    const ocsr *M = which == 0 ? L->W : which == 1 ? L->AfP : L->Af;
    int m = max_row_len(M); if (m > mx) mx = m;
  }
  double *buf = NEW(double, 2 * mx);
  for (int i = 0; i < n; i++) {
    int l = lvl[i] - 1;
    if (l >= nl) { len[i] = 0; continue; }
    const olevel *L = &h->lv[l];
    const ocsr *M = which == 0 ? L->W : which == 1 ? L->AfP : L->Af;
    const int *id = which == 2 ? L->idf : L->idc;
    int j = row[l]++;
    int kb = M->ro[j], ke = M->ro[j + 1];
    double *p = buf;
    for (int k = kb; k < ke; k++) { *p++ = id[M->col[k]]; *p++ = M->a[k]; }
    len[i] = ke - kb;
    fwrite(buf, sizeof(double), (size_t)(2 * (ke - kb)), f);
  }
  free(row); free(buf);
  fclose(f);
  return 0;
}
int amgo_export(const amgo_hier *h, const char *dir) {
  int nl = h->nlevels, n = h->lv[0].n;
  if (nl < 2) return -1;
  int *lvl = NEW(int, n);
  double *dvec = NEW(double, n);
  for (int i = 0; i < n; i++) { lvl[i] = 1; dvec[i] = 0; }
  for (int i = 0; i < nl - 1; i++)
    for (int j = 0; j < h->lv[i].nc; j++) lvl[h->lv[i].idc[j] - 1] += 1;
  for (int i = 0; i < nl - 1; i++)
    for (int j = 0; j < h->lv[i].nf; j++) dvec[h->lv[i].idf[j] - 1] = h->lv[i].D[j];
  int k = h->lv[nl - 2].idc[0] - 1;
  dvec[k] = h->nullspace ? 0. : 1. / h->lv[nl - 1].A->a[0];
  int *Wl = NEW(int, n), *Pl = NEW(int, n), *Fl = NEW(int, n);
  char path[1024];
  snprintf(path, sizeof path, "%s/amg_W.dat", dir);   if (save_mats(Wl, n, nl - 1, lvl, h, 0, path)) return -2;
  snprintf(path, sizeof path, "%s/amg_AfP.dat", dir); if (save_mats(Pl, n, nl - 1, lvl, h, 1, path)) return -2;
  snprintf(path, sizeof path, "%s/amg_Aff.dat", dir); if (save_mats(Fl, n, nl - 1, lvl, h, 2, path)) return -2;
  snprintf(path, sizeof path, "%s/amg.dat", dir);
  FILE *f = fopen(path, "wb");
  if (!f) return -2;
  const double magic = 3.14159, stamp = 2.01;
  double t;
  fwrite(&magic, sizeof(double), 1, f);
  fwrite(&stamp, sizeof(double), 1, f);
  t = nl; fwrite(&t, sizeof(double), 1, f);
  for (int i = 0; i < nl - 1; i++) { t = h->lv[i].m; fwrite(&t, sizeof(double), 1, f); }
  for (int i = 0; i < nl - 1; i++) { t = h->lv[i].rho; fwrite(&t, sizeof(double), 1, f); }
  t = n; fwrite(&t, sizeof(double), 1, f);
  for (int i = 0; i < n; i++) {
    double rec[6] = { (double)(i + 1), (double)lvl[i], (double)Wl[i], (double)Pl[i], (double)Fl[i], dvec[i] };
    fwrite(rec, sizeof(double), 6, f);
  }
  fclose(f);
  free(lvl); free(dvec); free(Wl); free(Pl); free(Fl);
  return 0;
}

/* ------------------------------------------------------------------------- */
/* V-cycle: amg_exec (amg.c:114) + crs_solve (amg.c:171), one process.        */
/* amg.c stores every level's F rows back to back; here each level keeps its   */
/* own numbering, which is the same arithmetic up to the order of the columns */
/* inside a row of W / AfP.  amg.c does not compile as a file, but its solve   */
/* path (amg.c:85-189) does in isolation: oracle/vcycle_ref_harness.c runs    */
/* those lines unchanged, and tests/test_oracle.py pins this restatement to   */
/* them bit for bit.                                                          */
/* ------------------------------------------------------------------------- */
static void vcycle(const amgo_hier *h, int l, double *x, double *b) {
  const olevel *L = &h->lv[l];
  int n = L->n;
  if (l == h->nlevels - 1) {
    double d = h->nullspace ? 0. : 1. / L->A->a[0];
    x[0] = d * b[0];
    return;
  }
  int nf = L->nf, nc = L->nc;
  double *bf = NEW(double, nf), *bc = NEW(double, nc), *xc = NEW(double, nc), *xf = NEW(double, nf);
  double *t = NEW(double, nc);
  { int cf = 0, cc = 0;
    for (int i = 0; i < n; i++) { if (L->C[i] != 0.) bc[cc++] = b[i]; else bf[cf++] = b[i]; } }
  /* b_{l+1} += W^t b_l */
  apply_Mt(t, L->W, bf);
  for (int i = 0; i < nc; i++) bc[i] = 1 * bc[i] + 1 * t[i];
  vcycle(h, l + 1, xc, bc);
  /* x_l = W x_{l+1};  b_l -= AfP x_{l+1} */
  apply_M(xf, 0, bf, 1, L->W, xc);
  apply_M(bf, 1, bf, -1, L->AfP, xc);
  double *c = NEW(double, nf), *co = NEW(double, nf), *r = NEW(double, nf);
  const double *d = L->D;
  unsigned m = (unsigned)L->m;
  double alpha = 0, beta = 0, gamma = 0;
  for (int i = 0; i < nf; i++) c[i] = d[i] * bf[i];
  if (m > 1) {
    alpha = L->rho / 2; alpha *= alpha;
    gamma = 2 * alpha / (1 - 2 * alpha); beta = 1 + gamma;
    apply_M(r, 1, bf, -1, L->Af, c);
    { double *s = c; c = co; co = s; }
    for (int i = 0; i < nf; i++) c[i] = beta * (co[i] + d[i] * r[i]);
  }
  for (unsigned ci = 3; ci <= m; ci++) {
    gamma = alpha * beta; gamma = gamma / (1 - gamma); beta = 1 + gamma;
    apply_M(r, 1, bf, -1, L->Af, c);
    { double *s = c; c = co; co = s; }
    for (int i = 0; i < nf; i++) c[i] = beta * (co[i] + d[i] * r[i]) - gamma * c[i];
  }
  for (int i = 0; i < nf; i++) xf[i] += c[i];
  { int cf = 0, cc = 0;
    for (int i = 0; i < n; i++) { if (L->C[i] != 0.) x[i] = xc[cc++]; else x[i] = xf[cf++]; } }
  free(bf); free(bc); free(xc); free(xf); free(t); free(c); free(co); free(r);
}
int amgo_solve(const amgo_hier *h, double *x, const double *b) {
  int n = h->lv[0].n;
  double *bb = NEW(double, n);
  memcpy(bb, b, sizeof(double) * (size_t)n);
  vcycle(h, 0, x, bb);
  if (h->nullspace) {
    /* crs_solve (amg.c:181-184) sums ux in ITS storage order: the unknowns sorted by the level at
       which they become F, ascending inside a level (amg.c:438-446), the last level's unknown at
       the end.  g[i] = position of top-level unknown i in that order. */
    int nl = h->nlevels;
    int *off = NEW(int, nl + 1);
    off[0] = 0;
    for (int l = 0; l < nl; l++) off[l + 1] = off[l] + (l < nl - 1 ? h->lv[l].nf : h->lv[l].n);
    int *g = NEW(int, h->lv[nl - 1].n + 1);
    for (int i = 0; i < h->lv[nl - 1].n; i++) g[i] = off[nl - 1] + i;
    for (int l = nl - 2; l >= 0; l--) {
      const olevel *L = &h->lv[l];
      int *gl = NEW(int, L->n + 1);
      int cf = 0, cc = 0;
      for (int i = 0; i < L->n; i++) gl[i] = (L->C[i] != 0.) ? g[cc++] : off[l] + cf++;
      free(g);
      g = gl;
    }
    double *ux = NEW(double, n + 1);
    for (int i = 0; i < n; i++) ux[g[i]] = x[i];
    double s = 0;
    for (int i = 0; i < n; i++) s += ux[i];
    double avg = (1 / (double)n) * s;
    for (int i = 0; i < n; i++) x[i] -= avg;
    free(ux); free(g); free(off);
  }
  free(bb);
  return 0;
}

/* refshim.c -- amg_setup / amg_export / free_data with the reference's signatures
 * (amg_setup.h:5,9; amg_setup.c:60, :405, :3487) on top of the C ABI of the CUDA engine.
 * Host-side glue only: every matrix operation happens in libomp_amg_b200.so on the GPU.
 * See include/amg_setup_b200.h. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/amg_setup_b200.h"
#include "../../include/omp_amg_b200.h"

struct amg_setup_data;   /* the caller's struct; laid out as struct amgb_amg_setup_data */

/* which hierarchy in HBM a host struct was built from (amg_export writes from the device copy) */
struct reg_entry { struct amgb_amg_setup_data *data; amgb_hier *h; };
static struct reg_entry *g_reg = NULL;
static int g_nreg = 0;

static void die(const char *what) {
  fprintf(stderr, "amg_setup (omp_amg_b200): %s: %s\n", what, amgb_last_error());
  exit(1);                       /* the reference's failure mode is fail() -> exit (fail.c) */
}
static void *xmalloc(size_t n) {
  void *p = malloc(n ? n : 1);
  if (!p) { fprintf(stderr, "amg_setup (omp_amg_b200): out of host memory\n"); exit(1); }
  return p;
}

static struct amgb_csr_mat *fetch_csr(const amgb_hier *h, int lvl, int which) {
  int32_t rn, cn;
  int64_t nnz;
  if (amgb_get_csr(h, lvl, which, &rn, &cn, &nnz, NULL, NULL, NULL)) die("amgb_get_csr");
  int32_t *ro = xmalloc(sizeof(int32_t) * ((size_t)rn + 1)), *col = xmalloc(sizeof(int32_t) * (size_t)nnz);
  struct amgb_csr_mat *M = xmalloc(sizeof *M);
  M->rn = (amgb_uint)rn; M->cn = (amgb_uint)cn;
  M->row_off = xmalloc(sizeof(amgb_uint) * ((size_t)rn + 1));
  M->col = xmalloc(sizeof(amgb_uint) * (size_t)nnz);
  M->a = xmalloc(sizeof(double) * (size_t)nnz);
  if (amgb_get_csr(h, lvl, which, NULL, NULL, NULL, ro, col, M->a)) die("amgb_get_csr");
  for (int32_t i = 0; i <= rn; i++) M->row_off[i] = (amgb_uint)ro[i];
  for (int64_t k = 0; k < nnz; k++) M->col[k] = (amgb_uint)col[k];
  free(ro); free(col);
  return M;
}

void amg_setup(amgb_uint n, const amgb_uint *Ai, const amgb_uint *Aj, const double *Av,
               struct amg_setup_data *data_) {
  struct amgb_amg_setup_data *data = (struct amgb_amg_setup_data *)data_;
  if ((uint64_t)n > 0x7fffffffULL) { fprintf(stderr, "amg_setup (omp_amg_b200): more than 2^31 entries\n"); exit(1); }
  int32_t *ai = xmalloc(sizeof(int32_t) * (size_t)n), *aj = xmalloc(sizeof(int32_t) * (size_t)n);
  for (amgb_uint k = 0; k < n; k++) { ai[k] = (int32_t)Ai[k]; aj[k] = (int32_t)Aj[k]; }
  amgb_hier *h = NULL;
  if (amgb_setup((int64_t)n, ai, aj, Av, &h)) die("amgb_setup");
  free(ai); free(aj);

  const double tol = 0.5, ctol = 0.7;                 /* amg_setup.c:71-78 */
  const double gamma2 = 1. - sqrt(1. - tol);
  data->tolc = ctol;
  data->gamma = sqrt(gamma2);
  const int nl = amgb_nlevels(h);
  const size_t cap = nl > 100 ? (size_t)nl : 100;     /* the reference allocates 100 slots (:88) */
  data->n = xmalloc(sizeof(double) * cap);      data->nnz = xmalloc(sizeof(double) * cap);
  data->nnzf = xmalloc(sizeof(double) * cap);   data->nnzfp = xmalloc(sizeof(double) * cap);
  data->m = xmalloc(sizeof(double) * cap);      data->rho = xmalloc(sizeof(double) * cap);
  data->idc = xmalloc(sizeof(amgb_uint *) * cap); data->idf = xmalloc(sizeof(amgb_uint *) * cap);
  data->C = xmalloc(sizeof(double *) * cap); data->F = xmalloc(sizeof(double *) * cap);
  data->D = xmalloc(sizeof(double *) * cap);
  data->A = xmalloc(sizeof(struct amgb_csr_mat *) * cap);   data->Af = xmalloc(sizeof(struct amgb_csr_mat *) * cap);
  data->W = xmalloc(sizeof(struct amgb_csr_mat *) * cap);   data->AfP = xmalloc(sizeof(struct amgb_csr_mat *) * cap);
  int64_t info[10];
  double par[4];
  amgb_level_info(h, 0, info);
  const int64_t n0 = info[0];
  data->id = xmalloc(sizeof(amgb_uint) * (size_t)n0);
  for (int64_t k = 0; k < n0; k++) data->id[k] = (amgb_uint)(k + 1);          /* :117 */
  for (int l = 0; l < nl; l++) {
    amgb_level_info(h, l, info);
    amgb_level_params(h, l, par);
    data->n[l] = (double)info[0];
    data->nnz[l] = (double)info[1];
    data->A[l] = fetch_csr(h, l, AMGB_A);
    printf("===================================================\n");      /* :158-162 */
    printf("Level %d, dim(A) = %d, nnz(A)/dim(A) = %lf\n", l + 1, (int)info[0], ((double)info[1]) / ((double)info[0]));
    printf("===================================================\n");
    if (l == nl - 1) break;
    const int64_t rn = info[0], nf = info[2], nc = info[3];
    data->nnzf[l] = (double)info[4];
    data->nnzfp[l] = (double)info[6];
    data->m[l] = par[0];
    data->rho[l] = par[1];
    data->Af[l] = fetch_csr(h, l, AMGB_AF);
    data->W[l] = fetch_csr(h, l, AMGB_W);
    data->AfP[l] = fetch_csr(h, l, AMGB_AFP);
    data->C[l] = xmalloc(sizeof(double) * (size_t)rn);
    data->F[l] = xmalloc(sizeof(double) * (size_t)rn);
    data->D[l] = xmalloc(sizeof(double) * (size_t)nf);
    if (amgb_get_vec(h, l, AMGB_C, data->C[l]) || amgb_get_vec(h, l, AMGB_D, data->D[l])) die("amgb_get_vec");
    for (int64_t i = 0; i < rn; i++) data->F[l][i] = (data->C[l][i] == 0.) ? 1. : 0.;   /* vf = not(vc) :183 */
    double *tmp = xmalloc(sizeof(double) * (size_t)(nf > nc ? nf : nc));
    data->idc[l] = xmalloc(sizeof(amgb_uint) * (size_t)nc);
    data->idf[l] = xmalloc(sizeof(amgb_uint) * (size_t)nf);
    if (amgb_get_vec(h, l, AMGB_IDC, tmp)) die("amgb_get_vec");
    for (int64_t i = 0; i < nc; i++) data->idc[l][i] = (amgb_uint)tmp[i];
    if (amgb_get_vec(h, l, AMGB_IDF, tmp)) die("amgb_get_vec");
    for (int64_t i = 0; i < nf; i++) data->idf[l][i] = (amgb_uint)tmp[i];
    free(tmp);
  }
  data->nlevels = (amgb_uint)nl;
  data->nullspace = (amgb_uint)amgb_nullspace(h);
  printf("===================================================\n");        /* :167-170 */
  printf("End of setup\n");
  printf("===================================================\n");
  printf("Nullspace = %u\n", (unsigned)data->nullspace);
  g_reg = realloc(g_reg, sizeof *g_reg * (size_t)(g_nreg + 1));
  g_reg[g_nreg].data = data; g_reg[g_nreg].h = h; g_nreg++;
}

static int find_reg(const struct amgb_amg_setup_data *data) {
  for (int i = 0; i < g_nreg; i++) if (g_reg[i].data == data) return i;
  return -1;
}

void amg_export(struct amg_setup_data *data_) {
  const int r = find_reg((struct amgb_amg_setup_data *)data_);
  if (r < 0) { fprintf(stderr, "amg_export (omp_amg_b200): this struct was not filled by amg_setup\n"); exit(1); }
  if (amgb_export(g_reg[r].h, ".")) die("amgb_export");
}

static void free_csr(struct amgb_csr_mat *M) { if (M) { free(M->row_off); free(M->col); free(M->a); free(M); } }

void free_data(struct amg_setup_data **data_) {
  struct amgb_amg_setup_data **pd = (struct amgb_amg_setup_data **)data_;
  if (!pd || !*pd) return;
  struct amgb_amg_setup_data *d = *pd;
  const int r = find_reg(d);
  if (r >= 0) { amgb_free(g_reg[r].h); g_reg[r] = g_reg[g_nreg - 1]; g_nreg--; }
  free(d->n); free(d->nnz); free(d->nnzf); free(d->nnzfp); free(d->m); free(d->rho);
  for (amgb_uint i = 0; i < d->nlevels; i++) free_csr(d->A[i]);
  for (amgb_uint i = 0; i + 1 < d->nlevels; i++) {
    free(d->C[i]); free(d->F[i]); free(d->D[i]); free(d->idc[i]); free(d->idf[i]);
    free_csr(d->Af[i]); free_csr(d->W[i]); free_csr(d->AfP[i]);
  }
  free(d->id); free(d->idc); free(d->idf); free(d->C); free(d->F); free(d->D);
  free(d->A); free(d->Af); free(d->W); free(d->AfP);
  free(d);
  *pd = NULL;
}

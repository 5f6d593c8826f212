/* oracle/vcycle_ref_harness.c -- TEST INFRASTRUCTURE, never part of the product.
 *
 * Runs the reference's own V-cycle code in isolation.  The reference's amg.c cannot be compiled
 * as a file in this tree (crs_setup calls amg_setup with seven arguments, amg.c:493; get_time and
 * barrier sit inside a comment, amg.c:29-43; and `struct crs_data` is defined nowhere), so the
 * recipe in oracle/Makefile cuts the four functions of the solve path out of the source where it
 * lies -- apply_Q, apply_Qt, amg_exec, crs_solve: amg.c:85-189, from the line
 * "static double apply_Q(" up to the line before "void crs_stats(" -- into the build output
 * oracle/_ref/amg_exec_fragment.inc, and this file compiles those lines UNCHANGED by including
 * them after supplying what the tree does not:
 *
 *   struct crs_data     restated from the way amg.c:114-189 and amg.c:295-470 use its fields
 *                       (the names and types are forced by those uses)
 *   barrier(), gs()     one process, every id unique: the gather-scatter is the identity
 *
 * apply_M / apply_Mt / get_time come from the reference's amg_tools.c (libamg_ref.so),
 * comm_reduce_double from its comm.c.  What this pins: the recurrence of amg_exec (restriction,
 * coarsest solve, prolongation, the Chebyshev coefficients and updates) and crs_solve's mean
 * projection, operation for operation.  What it cannot pin: the order of the entries inside a row
 * of W / AfP / Aff, which in the reference comes from reading amg.dat through the crystal router
 * (amg.c:813-947, needs the missing struct and MPI); the caller (oracle/oracle.py:
 * RefVcycle) keeps the level-local storage order of the hierarchy.
 */
#include <stddef.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include "c99.h"
#include "name.h"
#include "types.h"
#include "fail.h"
#include "mem.h"
#include "gs_defs.h"
#include "comm.h"
#include "gs.h"
#include "amg_tools.h"

struct crs_data {
  struct comm comm;
  struct gs_data *gs_top;
  uint un, *umap;
  double tni;
  int null_space;
  unsigned levels;
  unsigned *cheb_m;
  double *cheb_rho;
  uint *lvl_offset;
  double *Dff;
  struct Q *Q_W, *Q_AfP, *Q_Aff;
  struct csr_mat *W, *AfP, *Aff;
  double *b, *x, *c, *c_old, *r, *buf;
  double *timing;
  uint timing_n;
};

static void barrier(const struct comm *c) { (void)c; }
/* one process, unique ids: nothing to gather or scatter */
void gs(void *u, gs_dom dom, gs_op op, unsigned transpose, struct gs_data *gsh, buffer *buf)
{ (void)u; (void)dom; (void)op; (void)transpose; (void)gsh; (void)buf; }

#include "amg_exec_fragment.inc"

/* levels: number of levels (the last holds one unknown, or none)
 * off[levels+1]: lvl_offset (amg.c:117): the F unknowns of every level back to back
 * per level l < levels-1: W, AfP (nf_l rows, columns index the unknowns from off[l+1] on),
 *                         Aff (nf_l x nf_l); CSR with the reference's uint
 * umap[un]: position of the concatenated unknown in the caller's vector (amg.c:173)          */
int vref_solve(unsigned levels, const uint *off, const double *Dff, const unsigned *cheb_m, const double *cheb_rho,
               uint *const *ro, uint *const *col, double *const *a,   /* 3*(levels-1) matrices: W, AfP, Aff per level */
               uint un, const uint *umap, int null_space, double *x, double *b)
{
  struct crs_data d;
  unsigned l;
  uint nmax = 1, tot = off[levels];
  memset(&d, 0, sizeof d);
  d.comm.id = 0; d.comm.np = 1;
  d.levels = levels;
  d.lvl_offset = (uint *)off;
  d.Dff = (double *)Dff;
  d.cheb_m = (unsigned *)cheb_m;
  d.cheb_rho = (double *)cheb_rho;
  d.un = un; d.umap = (uint *)umap;
  d.null_space = null_space;
  d.tni = 1 / (double)un;                 /* amg.c:799: 1 / (global number of unknowns) */
  d.Q_W = tmalloc(struct Q, 3 * levels); d.Q_AfP = d.Q_W + levels; d.Q_Aff = d.Q_AfP + levels;
  d.W = tmalloc(struct csr_mat, 3 * levels); d.AfP = d.W + levels; d.Aff = d.AfP + levels;
  for (l = 0; l + 1 < levels; l++) {
    const uint nf = off[l + 1] - off[l], nrest = tot - off[l + 1];
    struct csr_mat *M[3]; unsigned k;
    M[0] = &d.W[l]; M[1] = &d.AfP[l]; M[2] = &d.Aff[l];
    for (k = 0; k < 3; k++) {
      M[k]->rn = nf; M[k]->cn = (k == 2) ? nf : nrest;
      M[k]->row_off = ro[3 * l + k]; M[k]->col = col[3 * l + k]; M[k]->a = a[3 * l + k];
    }
    d.Q_W[l].nloc = nrest; d.Q_AfP[l].nloc = nrest; d.Q_Aff[l].nloc = nf;
    d.Q_W[l].gsh = d.Q_AfP[l].gsh = d.Q_Aff[l].gsh = NULL;
    if (nf > nmax) nmax = nf;
    if (nrest > nmax) nmax = nrest;
  }
  d.b = tcalloc(double, tot + 1); d.x = tcalloc(double, tot + 1);
  d.c = tcalloc(double, nmax); d.c_old = tcalloc(double, nmax); d.r = tcalloc(double, nmax);
  d.buf = tcalloc(double, nmax);
  d.timing = tcalloc(double, 6 * levels);
  crs_solve(x, &d, b);
  free(d.b); free(d.x); free(d.c); free(d.c_old); free(d.r); free(d.buf); free(d.timing);
  free(d.Q_W); free(d.W);
  return 0;
}

// sparse.cuh -- device CSR matrices and the sparse primitives of the setup path.
// Every primitive cites the reference routine whose result it reproduces bit for bit
// (file amg_setup.c unless stated).
#pragma once
#include "common.cuh"

namespace amgb {

// CSR in HBM: ro[rn+1] (int32), col[nnz] (int32, ascending inside a row), a[nnz] (fp64).
// struct csr_mat, amg_tools.h:5.
struct Csr {
  int rn = 0, cn = 0;
  i64 nnz = 0;
  Buf<int> ro, col;
  Buf<double> a;
  // identity of the stored matrix for caches (never a recyclable device pointer): fresh for every
  // constructed or cloned matrix, carried along by moves
  unsigned long long uid = 0;
  static unsigned long long next_uid() { static unsigned long long g = 0; return ++g; }
  // rows too long for the group-per-row SpMV kernels (a few thousand entries: the column-0 pile
  // of min_skel transposed, the rows of the coarsest levels), found at the first SpMV with this
  // matrix and handled by a block each (sparse.cu: k_spmv_chain); -1 = not looked for yet
  mutable int n_long = -1;
  mutable Buf<int> long_rows;
  // Solve-phase storage partition (csr_keep_row_block): only rows [row_lo, row_lo + own_rows) are
  // stored -- ro has own_rows + 1 entries starting at 0, col/a the own_nnz entries of the block --
  // while rn, cn and nnz keep describing the whole matrix (kernel selection and thresholds must
  // not depend on the partition).  Such a matrix only serves row-partitioned products.
  bool partial = false;
  int row_lo = 0, own_rows = 0;
  i64 own_nnz = 0;
  Csr() {}
  Csr(int rn_, int cn_, i64 nnz_) : rn(rn_), cn(cn_), nnz(nnz_), ro(rn_ + 1), col(nnz_), a(nnz_), uid(next_uid()) {}
  Csr(Csr &&) = default;
  Csr &operator=(Csr &&) = default;
  Csr clone() const {                                   // copy_csr :3463
    Csr B;
    B.rn = rn; B.cn = cn; B.nnz = nnz; B.uid = next_uid();
    B.ro = ro.clone(); B.col = col.clone(); B.a = a.clone();
    B.partial = partial; B.row_lo = row_lo; B.own_rows = own_rows; B.own_nnz = own_nnz;
    return B;
  }
};

// apply_M (amg_tools.c:71): z = alpha*y + beta*(M x); y may be null (then z = beta*(M x))
void spmv(double *z, double alpha, const double *y, double beta, const Csr &M, const double *x);
// rows [r0, r1) of the same product only (z, y indexed by absolute row): the row blocks of the
// partitioned V-cycle
void spmv_rows(double *z, double alpha, const double *y, double beta, const Csr &M, const double *x, int r0, int r1);
// keeps rows [r0, r1) of M and releases the rest (see Csr::partial); returns the bytes released
i64 csr_keep_row_block(Csr &M, int r0, int r1);
// row-partitioned products over several ranks (see sparse.cu): always on inside a scope of this
// type (the V-cycle), elsewhere unless AMGB_DIST_SPMV=0
struct SpmvPartitionScope { int prev; SpmvPartitionScope(); ~SpmvPartitionScope(); };
bool spmv_is_partitioned(const Csr &M);
// values-only variant: same pattern as M, other value array
void spmv_vals(double *z, double alpha, const double *y, double beta, const Csr &M, const double *vals,
               const double *x, const double *post = nullptr,    // post != null: z[i] = (...) * post[i]
               const int *gen = nullptr, int want = 0);          // gen != null: only rows with gen[i] == want
// x == null stands for a vector of ones (row sums: no column is read, nothing is gathered)

// transpose :2000.  If tpos != null it receives, for every entry e of A, its position in A^t.
Csr transpose(const Csr &A, Buf<int> *tpos = nullptr);
// sub_mat :3058; a null flag array means "keep all"
Csr sub_mat(const Csr &A, const double *vr, const double *vc);
// mpm :1684
Csr mpm(double alpha, const Csr &A, double beta, const Csr &B);
// mxmpoint :1807
Csr mxmpoint(const Csr &A, const Csr &B);
// mxm :1894 as X = A*B (row-wise, every X[i][c] accumulated over k ascending, exact zeros dropped)
Csr spgemm(const Csr &A, const Csr &B);
// build_csr_dim :3656 from device COO arrays (zeros dropped, sorted by row then column)
Csr coo_to_csr(i64 n, const int *Ai, const int *Aj, const double *Av, int rn, int cn);

void diag_of(double *D, const Csr &A);            // diag :3363
void scale_rows(Csr &A, const double *D);         // diagcsr_op dmult :3426
void scale_cols(Csr &A, const double *D);         // diagcsr_op multd :3436
void sub_diag(Csr &A, const double *D);           // diagcsr_op dminus :3412
void col_sums(double *s, const Csr &A);           // sum(.,.,1) :1193
int max_row_len(const Csr &A);

void trace_csr(const char *tag, const Csr &A);

// opt-in statistics of the long-row SpMV kernels (more than 24 entries per row): device time and
// algorithmic bytes ((12 nnz, 8 without a gather) + 12 or 20 per row) since the last reset
void spmv_stats_enable(bool on);
void spmv_stats_reset();
void spmv_stats_get(double *seconds, i64 *bytes, i64 *calls);
// device time / algorithmic bytes of the SpGEMM kernels since the last reset (spgemm.cu)
void spgemm_stats_reset();
void spgemm_cache_reset();      // drops the cached transpose of the last large left operand
void spgemm_stats_get(double *seconds, i64 *bytes, i64 *calls);
// diagnostics: rows per SpGEMM tier of the last product ([0..11) as binned, [11..22) after hand-downs)
void spgemm_debug_collect(bool on);
void spgemm_debug_tiers(int out[22]);

// element-wise helpers
void fill(double *p, i64 n, double v);
void fill_int(int *p, i64 n, int v);

}  // namespace amgb

// localsolve.cuh -- the per-coarse-point energy-minimising solves of interp (:2053) and
// interp_lmop (:1589), factored so that the A-orthogonal basis Q of every coarse column is built
// once per skeleton and reused by the three places that need it.
#pragma once
#include "sparse.cuh"

namespace amgb {

// Packed upper-triangular Q of every row i of Wt (a coarse point and its sorted F support):
// Q_i occupies Q[qoff[i], qoff[i] + nz_i(nz_i+1)/2), column k at offset k(k+1)/2.
struct QStore {
  Buf<i64> qoff;      // rn+1
  Buf<double> Q;
  i64 total = 0;
  int maxnz = 0;
  // columns with a huge support (the column-0 pile of min_skel): handled one at a time by the
  // blocked kernels that use the whole GPU; everything else is binned by support size
  int bignz = 0x7fffffff;          // supports larger than this are "huge"
  int maxnz_small = 0;             // largest support among the others
  struct Huge { int col, nz, wb; i64 qoff; };
  std::vector<Huge> huge;
};

// Q Q^t of every column, dense nz_i x nz_i blocks at QQ[qqoff[i]]
struct QQStore {
  Buf<i64> qqoff;
  Buf<double> QQ;
};

// build Q for all rows of Wt against At (= Af, symmetric)  (:2081-2099)
void build_q_store(QStore &qs, const Csr &Wt, const Csr &At);
// values of Wt := Q Q^t R (B e_i + u_i lambda)               (:2100-2108)
void apply_q(const QStore &qs, Csr &Wt, const Csr &Bt, const double *u, const double *lambda);
// QQ_i = sum_k q_k q_k^t, k ascending                        (:1637-1642)
void form_qq(QQStore &qq, const QStore &qs, const Csr &Wt);
// S := sum_i u_i QQ_i scattered to the pattern of S, contributions in ascending i (:1644-1650);
// targets absent from the pattern are skipped (DESIGN.md "sp_add")
void lmop_accumulate(Csr &S, const QQStore &qq, const double *u, const Csr &Wskt, const Csr &Wsk,
                     const int *tpos);

}  // namespace amgb

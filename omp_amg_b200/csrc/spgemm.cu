// spgemm.cu -- Galerkin-product SpGEMM (mxm, amg_setup.c:1894).
#include "sparse.cuh"

namespace amgb {
Csr spgemm_rowhash(const Csr &A, const Csr &B);
#ifndef AMGB_EMU
Csr spgemm(const Csr &A, const Csr &B) { return spgemm_rowhash(A, B); }
#endif
}  // namespace amgb

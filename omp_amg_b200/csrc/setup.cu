// setup.cu -- coarsening, smoother estimate, interpolation and Galerkin product on the GPU.
// Host code only orchestrates: every matrix- or vector-sized operation is a kernel.
// All citations are amg_setup.c lines of the reference unless another file is named.
#include "setup.cuh"
#include "comm.cuh"
#include "localsolve.cuh"
#include <chrono>

namespace amgb {

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// =======================================================================================
// coarsen (:2737) with mat_max (:3535)
// =======================================================================================
#ifndef AMGB_EMU
// G-lanes-per-row variants of the three coarsening kernels for levels with longer rows (maxima
// are order-free, so the lanes reduce with shuffles).  G = 8 for rows of a few dozen entries: the
// kernels are chains of dependent gathers (row offsets -> column index -> thr/w/g of that row),
// so what matters is how many rows are in flight, and a warp holds four of them.
template <int G>
__global__ void __launch_bounds__(256) k_coarsen_thr(int n, const int *ro, const int *col, const double *sa,
                                                     const double *vf, double mtol, double *thr) {
  const int i = blockIdx.x * (256 / G) + threadIdx.x / G;
  if (i >= n) return;
  const int lane = threadIdx.x % G;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
  double amax = 0;
  for (int j = ro[i] + lane; j < ro[i + 1]; j += G)
    if (vf[col[j]] != 0 && fabs(sa[j]) > amax) amax = fabs(sa[j]);
  for (int off = G / 2; off >= 1; off >>= 1) amax = fmax(amax, __shfl_down_sync(gmask, amax, off, G));
  if (lane == 0) thr[i] = amax * mtol;
}
template <int STAGE, int G>
__global__ void __launch_bounds__(256) k_coarsen_gather(int n, const int *tro, const int *tcol, const double *ta,
                                                        const double *thr, const double *vf, const double *w,
                                                        const double *g, double ctol2, double *mk, double *tp,
                                                        double *vc, double *vfnext) {
  const int k = blockIdx.x * (256 / G) + threadIdx.x / G;
  if (k >= n) return;
  const int lane = threadIdx.x % G;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
  double m = -DBL_MAX;
  if (vf[k] != 0)
    for (int p = tro[k] + lane; p < tro[k + 1]; p += G) {
      const int i = tcol[p];
      if (fabs(ta[p]) < thr[i]) continue;
      double x;
      if (STAGE == 1) { const double mi0 = (w[i] > ctol2) ? 1. : 0.; x = g[i] * mi0; }
      else x = tp[i];
      if (x > m) m = x;
    }
  for (int off = G / 2; off >= 1; off >>= 1) m = fmax(m, __shfl_down_sync(gmask, m, off, G));
  if (lane != 0) return;
  if (STAGE == 1) {
    const double m0 = (w[k] > ctol2) ? 1. : 0.;
    const double d = g[k] - m;
    const double mv = (m0 != 0. && d >= 0.) ? 1. : 0.;
    mk[k] = mv;
    vfnext[k] = mv * ((double)k + 1.0);        // stage 1 writes the id vector into a scratch array
  } else {
    const double d = ((double)k + 1.0) - m;
    const double mv = (mk[k] != 0. && d > 0.) ? 1. : 0.;
    if (mv != 0.) vc[k] = 1.;
    vfnext[k] = ((vf[k] == 0.) != (mv == 0.)) ? 1. : 0.;
  }
}
#endif

// mat_max (:3535): y[k] = max over rows i that hold k as a strong neighbour (|S_ik| >= thr_i,
// thr_i = tol * max over F neighbours of |S_i.|, and f[k] != 0) of x[i].  The reference scatters
// row by row; a maximum is order-free, so every k gathers from row k of S' (which lists exactly the
// pairs (i, S_ik)) -- no atomics.  thr depends on (S, f) only and is shared by the two mat_max
// calls of a coarsening round; the element-wise mask updates around them are fused in.
int coarsen(double *vc, const Csr &A, double ctol) {
  StageTimer st_("coarsen");
  const int n = A.cn;
  int rounds = 0;
  Buf<double> D(n);
  diag_of(D.p, A);
  { double *d = D.p; parallel_for(n, [=] DEV(i64 i) { double s = sqrt(d[i]); d[i] = 1. / s; }); }
  Csr S = A.clone();
  scale_rows(S, D.p);
  scale_cols(S, D.p);
  { double *a = S.a.p; parallel_for(S.nnz, [=] DEV(i64 e) { a[e] = fabs(a[e]); }); }
  diag_of(D.p, S);
  sub_diag(S, D.p);
  trace_csr("coarsen.S", S);

  Buf<double> vf(n), g(n), w1(n), w2(n), tmp(n), w(n), mask(n), thr(n);
  Csr St = transpose(S);
  fill(vc, n, 0.);
  fill(vf.p, n, 1.);
  double *vfp = vf.p, *gp = g.p, *w1p = w1.p, *w2p = w2.p, *tp = tmp.p, *wp = w.p, *mk = mask.p, *thp = thr.p;
  const int *ro = S.ro.p, *col = S.col.p, *tro = St.ro.p, *tcol = St.col.p;
  const double *sa = S.a.p, *ta = St.a.p;
  const double ctol2 = ctol * ctol, mtol = 0.1;
  for (;;) {
    rounds++;
    stage_count("coarsen.rounds", 1);
    // w1 = vf.*(S*(vf.*(S*vf))), w2 = vf.*(S*(vf.*(S*w1))), w = (1./w1).*w2 (0 where w1 == 0)
    spmv_vals(gp, 0, nullptr, 1., S, sa, vfp, vfp);
    spmv_vals(w1p, 0, nullptr, 1., S, sa, gp, vfp);
    spmv_vals(w2p, 0, nullptr, 1., S, sa, w1p, vfp);
    spmv_vals(tp, 0, nullptr, 1., S, sa, w2p, vfp);
    parallel_for(n, [=] DEV(i64 i) {
      const double w2v = tp[i];
      w2p[i] = w2v;
      const double inv = 1. / w1p[i];
      double wv = inv * w2v;
      if (w1p[i] == 0) wv = 0.;
      wp[i] = wv;
    });
    double w1m, wm;
    i64 mi;
    max_first2(w1p, wp, n, &w1m, &mi, &wm);
    const double b = (w1m < wm) ? sqrt(w1m) : sqrt(wm);
    if (b <= ctol) {
      if (count_nonzero(vc, n) == 0) parallel_for(1, [=] DEV(i64) { vc[mi] = 1.; });
      break;
    }
#ifndef AMGB_EMU
    if ((double)S.nnz / (double)n > 16.0) {
      Context &cx = ctx();
      // stage 1 must not overwrite tmp while other warps may still read it as x: it does not read tp,
      // so tp is its output; stage 2 reads tp and writes the next vf into w2
      if ((double)S.nnz / (double)n <= 64.0) {
        const int nb = (n + 31) / 32;
        k_coarsen_thr<8><<<nb, 256, 0, cx.stream>>>(n, ro, col, sa, vfp, mtol, thp);
        k_coarsen_gather<1, 8><<<nb, 256, 0, cx.stream>>>(n, tro, tcol, ta, thp, vfp, wp, gp, ctol2, mk, nullptr, nullptr, tp);
        k_coarsen_gather<2, 8><<<nb, 256, 0, cx.stream>>>(n, tro, tcol, ta, thp, vfp, wp, gp, ctol2, mk, tp, vc, w2p);
      } else {
        const int nb = (n + 7) / 8;
        k_coarsen_thr<32><<<nb, 256, 0, cx.stream>>>(n, ro, col, sa, vfp, mtol, thp);
        k_coarsen_gather<1, 32><<<nb, 256, 0, cx.stream>>>(n, tro, tcol, ta, thp, vfp, wp, gp, ctol2, mk, nullptr, nullptr, tp);
        k_coarsen_gather<2, 32><<<nb, 256, 0, cx.stream>>>(n, tro, tcol, ta, thp, vfp, wp, gp, ctol2, mk, tp, vc, w2p);
      }
      cx.launches += 3; post_launch("coarsen_warp_kernels");
      { double *t = vfp; vfp = w2p; w2p = t; }
      if (ctx().trace_on) {
        char t[64];
        snprintf(t, sizeof t, "coarsen.vc.r%d", rounds);
        trace_dev(t, vc, sizeof(double) * (size_t)n);
      }
      if (rounds > 100000) throw Error(-7, "coarsen: no convergence");
      continue;
    }
#endif
    // thr_i of this round
    parallel_for(n, [=] DEV(i64 i) {
      double amax = 0;
      for (int j = ro[i]; j < ro[i + 1]; j++)
        if (vfp[col[j]] != 0 && fabs(sa[j]) > amax) amax = fabs(sa[j]);
      thp[i] = amax * mtol;
    });
    // mask = (w > ctol^2) & (g - mat_max(S, vf, mask.*g) >= 0);  tmp = mask.*id
    parallel_for(n, [=] DEV(i64 k) {
      double m = -DBL_MAX;
      if (vfp[k] != 0)
        for (int p = tro[k]; p < tro[k + 1]; p++) {
          const int i = tcol[p];
          if (fabs(ta[p]) < thp[i]) continue;
          const double mi0 = (wp[i] > ctol2) ? 1. : 0.;
          const double x = gp[i] * mi0;
          if (x > m) m = x;
        }
      const double m0 = (wp[k] > ctol2) ? 1. : 0.;
      const double d = gp[k] - m;
      const double mv = (m0 != 0. && d >= 0.) ? 1. : 0.;
      mk[k] = mv;
      tp[k] = mv * ((double)k + 1.0);
    });
    // mask = mask & (id - mat_max(S, vf, mask.*id) > 0);  vc |= mask;  vf ^= mask
    parallel_for(n, [=] DEV(i64 k) {
      double m = -DBL_MAX;
      if (vfp[k] != 0)
        for (int p = tro[k]; p < tro[k + 1]; p++) {
          const int i = tcol[p];
          if (fabs(ta[p]) < thp[i]) continue;
          if (tp[i] > m) m = tp[i];
        }
      const double d = ((double)k + 1.0) - m;
      const double mv = (mk[k] != 0. && d > 0.) ? 1. : 0.;
      if (mv != 0.) vc[k] = 1.;
      w2p[k] = ((vfp[k] == 0.) != (mv == 0.)) ? 1. : 0.;     // next vf, applied after the kernel
    });
    { double *t = vfp; vfp = w2p; w2p = t; }
    if (ctx().trace_on) {
      char t[64];
      snprintf(t, sizeof t, "coarsen.vc.r%d", rounds);
      trace_dev(t, vc, sizeof(double) * (size_t)n);
    }
    if (rounds > 100000) throw Error(-7, "coarsen: no convergence");
  }
  return rounds;
}

// =======================================================================================
// pcg (:2242)
// =======================================================================================
int pcg(double *x, const Csr &A, double *r, const double *M, double tol, const double *b) {
  StageTimer st_("pcg");
  const int n = A.rn;
  Buf<double> p(n), z(n), w(n), t(n);
  double *pp = p.p, *zp = z.p, *wp = w.p, *tp = t.p;
  parallel_for(n, [=] DEV(i64 i) { x[i] = 0.; pp[i] = 0.; zp[i] = M[i] * r[i]; tp[i] = M[i] * b[i]; });
  double rho = vdot(r, zp, n);
  const double rho_0 = vdot(tp, b, n);
  const double rho_stop = tol * tol * rho_0;
  const int nmax = n <= 100 ? n : 100;
  int k = 0;
  double rho_old = 1;
  while (nmax > 0 && rho > rho_stop && k < nmax) {
    k++;
    stage_count("pcg.iterations", 1);
    const double beta = rho / rho_old;
    parallel_for(n, [=] DEV(i64 i) { const double pb = pp[i] * beta; pp[i] = pb + zp[i]; });
    spmv(wp, 0, nullptr, 1, A, pp);
    double alpha = vdot(pp, wp, n);
    alpha = rho / alpha;
    parallel_for(n, [=] DEV(i64 i) {
      const double pa = pp[i] * alpha;
      x[i] = x[i] + pa;
      const double wa = wp[i] * alpha;
      const double rn = r[i] - wa;
      r[i] = rn;
      zp[i] = M[i] * rn;
    });
    rho_old = rho;
    rho = vdot(r, zp, n);
  }
  return k;
}

// =======================================================================================
// lanczos (:2435); tdeig on the host (k x k, k <= 299)
// =======================================================================================
int lanczos(double *lambda, const Csr &A, hostmath::GlibcRand &rng, int *iters) {
  StageTimer st_("lanczos");
  const int rn = A.rn, kmax = 299;
  std::vector<double> hr((size_t)rn);
  for (int i = 0; i < rn; i++) hr[(size_t)i] = rng.uniform();
  Buf<double> r(rn), qk(rn), qkm1(rn), Aqk(rn);
  r.upload(hr.data(), rn);
  trace_dev("lanczos.r0", r.p, sizeof(double) * (size_t)rn);
  double *l = lambda;
  double y[300], d[301], v[300];
  double beta = vnorm2(r.p, rn);
  double beta2 = beta * beta;
  beta = sqrt(beta2);
  int k = 0;
  double change = 0.0;
  {
    Buf<double> e = A.a.clone();
    const int *ro = A.ro.p, *col = A.col.p;
    double *ep = e.p;
    parallel_for(rn, [=] DEV(i64 i) {
      for (int j = ro[i]; j < ro[i + 1]; j++) if (col[j] == i) { ep[j] = ep[j] - 1.; break; }
    });
    double fro = vnorm2(e.p, A.nnz);
    const double fro2 = fro * fro;
    fro = sqrt(fro2);
    if (fro < 1e-11) { l[0] = 1; l[1] = 1; y[0] = 0; y[1] = 0; k = 2; change = 0.0; }
    else change = 1.0;
  }
  if (rn == 1) { const double a00 = A.a.get(0); l[0] = a00; l[1] = a00; y[0] = 0; y[1] = 0; k = 2; change = 0.0; }
  qk.zero();
  double *rp = r.p, *qp = qk.p, *qmp = qkm1.p, *Ap = Aqk.p;
  while (k < kmax && (change > 1e-5 || y[0] > 1e-3 || y[k - 1] > 1e-3)) {
    k++;
    stage_count("lanczos.iterations", 1);
    const double ib = 1. / beta;
    parallel_for(rn, [=] DEV(i64 i) { qmp[i] = qp[i]; qp[i] = rp[i] * ib; });
    spmv(Ap, 0, nullptr, 1, A, qp);
    const double alpha = vdot(qp, Ap, rn);
    const double bt = beta;
    parallel_for(rn, [=] DEV(i64 i) {
      const double aq = qp[i] * alpha, bq = qmp[i] * bt;
      const double t = Ap[i] - aq;
      rp[i] = t - bq;
    });
    if (k == 1) { l[0] = alpha; y[0] = 1; }
    else {
      const double l0 = l[0], lkm2 = l[k - 2];
      d[0] = 0;
      for (int i = 1; i < k; i++) d[i] = l[i - 1];
      d[k] = 0;
      v[0] = alpha;
      for (int i = 1; i < k; i++) v[i] = beta * y[i - 1];
      hostmath::tdeig(l, y, d, v, k - 1);
      change = fabs(l0 - l[0]) + fabs(lkm2 - l[k - 1]);
    }
    beta = vnorm2(rp, rn);
    beta2 = beta * beta;
    beta = sqrt(beta2);
    if (beta == 0) break;
  }
  if (iters) *iters = k;
  int n = 0;
  for (int i = 0; i < k; i++) if (y[i] < 0.01) lambda[n++] = l[i];
  return n;
}

// =======================================================================================
// interpolation weights: the local solves live in localsolve.cu
// =======================================================================================
// per-skeleton data shared by the solves of one interpolation round: W_skel^t with its position
// map, the packed Q of every coarse column, their Q Q^t blocks and the pattern of W_skel*W_skel^t
struct SkelCache {
  bool valid = false, has_qq = false, has_S = false;
  Csr Wskt, Spat;
  Buf<int> tpos;
  QStore qs;
  QQStore qq;
};

// min_skel (:2198)
Csr min_skel(const Csr &R) {
  Csr W(R.rn, R.cn, R.rn);
  const int *ro = R.ro.p, *col = R.col.p;
  const double *a = R.a.p;
  int *wro = W.ro.p, *wcol = W.col.p;
  double *wa = W.a.p;
  const int rn = R.rn;
  parallel_for(rn, [=] DEV(i64 i) {
    double ymax = -DBL_MAX;
    int j = 0;
    for (int k = ro[i]; k < ro[i + 1]; k++) if (a[k] > ymax) { ymax = a[k]; j = col[k]; }
    wa[i] = (ymax > 0.0) ? 1.0 : 0.0;
    wcol[i] = j;
    wro[i] = (int)i;
    if (i == rn - 1) wro[rn] = rn;
  });
  if (rn == 0) W.ro.zero();
  return W;
}

// solve_constraint (:1499)
void solve_constraint(double *lam, const Csr &Wsk, SkelCache &sk, const Csr &W0, const double *alpha,
                      const double *u, const double *v, double tol) {
  StageTimer st_("solve_constraint(incl pcg)");
  const int nf = Wsk.rn, nc = Wsk.cn;
  Buf<double> au2(nc);
  { double *p = au2.p; parallel_for(nc, [=] DEV(i64 i) { const double uu = u[i] * u[i]; p[i] = uu * alpha[i]; }); }
  if (!sk.has_qq) { StageTimer t_("sc.form_qq"); form_qq(sk.qq, sk.qs, sk.Wskt); sk.has_qq = true; }
  Csr S = sk.Spat.clone();
  { StageTimer t_("sc.lmop_acc"); lmop_accumulate(S, sk.qq, au2.p, sk.Wskt, Wsk, sk.tpos.p); }
  trace_csr("sc.S", S);
  Buf<double> resid(nf), d(nf), keep(nf);
  spmv(resid.p, 1.0, v, -1.0, W0, u);
  diag_of(d.p, S);
  double *dp = d.p, *kp = keep.p, *rp = resid.p;
  parallel_for(nf, [=] DEV(i64 i) { kp[i] = (dp[i] != 0.) ? 1. : 0.; });
  const i64 nk = count_nonzero(kp, nf);
  if (nk == nf) {
    Buf<double> q(nf), x(nf);
    spmv(q.p, 1., rp, -1., S, lam);
    parallel_for(nf, [=] DEV(i64 i) { dp[i] = 1. / dp[i]; });
    pcg(x.p, S, q.p, dp, tol, rp);
    double *xp = x.p;
    parallel_for(nf, [=] DEV(i64 i) { lam[i] = lam[i] + xp[i]; });
    return;
  }
  // some rows of S are empty: solve on the kept rows only (:1532-1574)
  Csr S2 = sub_mat(S, kp, kp);
  Buf<int> kflag(nf + 1), pos(nf + 1);
  int *kf = kflag.p;
  parallel_for(nf, [=] DEV(i64 i) { kf[i] = (kp[i] != 0.) ? 1 : 0; if (kp[i] == 0.) lam[i] = 0.; });
  exclusive_scan(kflag.p, pos.p, nf);
  Buf<double> rc(nk), dc(nk), lc(nk), q(nk), x(nk);
  const int *ps = pos.p;
  double *rcp = rc.p, *dcp = dc.p, *lcp = lc.p, *xp = x.p;
  parallel_for(nf, [=] DEV(i64 i) {
    if (!kf[i]) return;
    rcp[ps[i]] = rp[i]; dcp[ps[i]] = dp[i]; lcp[ps[i]] = lam[i];
  });
  spmv(q.p, 1., rcp, -1., S2, lcp);
  parallel_for(nk, [=] DEV(i64 i) { dcp[i] = 1. / dcp[i]; });
  pcg(xp, S2, q.p, dcp, tol, rcp);
  parallel_for(nf, [=] DEV(i64 i) { if (kf[i]) lam[i] = lam[i] + xp[ps[i]]; });
}

// solve_weights (:1437).  W and W0 share the pattern of W_skel; their values come back from the
// transposed solves through the transpose position map, which equals transposing them again.
void solve_weights(Csr &W, Csr &W0, double *lam, const Csr &Wsk, const Csr &Af, const Csr &Armt,
                   const double *alpha, const double *u, const double *v, double tol, SkelCache &sk) {
  const int nf = Af.rn, nc = Wsk.cn;
  Buf<double> au(nc), zeros(nf);
  { double *p = au.p; parallel_for(nc, [=] DEV(i64 i) { p[i] = alpha[i] * u[i]; }); }
  zeros.zero();
  if (!sk.valid) {       // the basis Q of every coarse column depends on the skeleton and Af only
    { StageTimer t_("sw.transpose"); sk.Wskt = transpose(Wsk, &sk.tpos); }
    { StageTimer t_("sw.spgemm_S"); sk.Spat = spgemm(Wsk, sk.Wskt); }  // pattern of W_skel*W_skel' (:1513), while Wskt holds 0/1
    { StageTimer t_("sw.build_q"); build_q_store(sk.qs, sk.Wskt, Af); }
    sk.valid = true; sk.has_qq = false; sk.has_S = true;
  }
  Csr &Wskt = sk.Wskt;
  const int *tp = sk.tpos.p;
  { StageTimer t_("sw.apply_q"); apply_q(sk.qs, Wskt, Armt, au.p, zeros.p); }   // interp(W0t, Af, -Ar', au, 0) :1467
  W0 = Wsk.clone();
  { double *dst = W0.a.p; const double *src = Wskt.a.p; parallel_for(Wsk.nnz, [=] DEV(i64 e) { dst[e] = src[tp[e]]; }); }
  solve_constraint(lam, Wsk, sk, W0, alpha, u, v, tol);
  trace_dev("sw.lam", lam, sizeof(double) * (size_t)nf);
  { StageTimer t_("sw.apply_q"); apply_q(sk.qs, Wskt, Armt, au.p, lam); }       // interp(Wt, Af, -Ar', au, lam) :1482
  W = Wsk.clone();
  { double *dst = W.a.p; const double *src = Wskt.a.p; parallel_for(Wsk.nnz, [=] DEV(i64 e) { dst[e] = src[tp[e]]; }); }
}

#ifndef AMGB_EMU
// One warp per column of R: the column is "bad" if w > thr and it still holds a non-zero (its
// entries are non-negative, so "sum != 0" of the reference is "some entry != 0"); the warp then
// finds the first row (ascending) with the largest R_ij * rs_i -- a max with smallest index, which
// is order-free -- and removes that entry from R and R'.
__global__ void __launch_bounds__(256) k_find_support_cols(int nc, const int *tro, const int *tcol, double *rtv,
                                                           const double *rs, const double *w, double thr, int *alive,
                                                           double *rv, const int *src, int *skel, int *gen, int next) {
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c >= nc) return;
  const int lane = threadIdx.x & 31;
  if (!(w[c] > thr)) return;
  const int b = tro[c], e = tro[c + 1];
  int any = 0;
  double best = -DBL_MAX;
  int arg = 0x7fffffff;
  for (int p = b + lane; p < e; p += 32) {
    const double v = rtv[p];
    any |= (v != 0.);
    if (!alive[p]) continue;
    const double x = v * rs[tcol[p]];
    if (x > best) { best = x; arg = p; }          // p ascending per lane: first maximum kept
  }
  any = __any_sync(0xffffffffu, any);
  if (!any) return;
  for (int off = 16; off >= 1; off >>= 1) {
    const double ob = __shfl_down_sync(0xffffffffu, best, off);
    const int oa = __shfl_down_sync(0xffffffffu, arg, off);
    if (oa != 0x7fffffff && (arg == 0x7fffffff || ob > best || (ob == best && oa < arg))) { best = ob; arg = oa; }
  }
  if (lane == 0 && arg != 0x7fffffff) {
    alive[arg] = 0; rtv[arg] = 0.0;
    rv[src[arg]] = 0.0; skel[src[arg]] = 1;
    gen[tcol[arg]] = next;                        // row tcol[arg] of R lost an entry: its sum is redone next round
  }
}

// The same with one block of 256 threads per column, for the coarse levels where R has few, long
// columns (1400 columns of 3000 entries): a warp per column leaves most of the GPU idle and walks
// ~100 dependent gather rounds.
__global__ void __launch_bounds__(256) k_find_support_cols_block(int nc, const int *tro, const int *tcol, double *rtv,
                                                                 const double *rs, const double *w, double thr,
                                                                 int *alive, double *rv, const int *src, int *skel,
                                                                 int *gen, int next) {
  __shared__ double sbest[8];
  __shared__ int sarg[8], sany[8];
  const int c = blockIdx.x;
  if (c >= nc) return;
  if (!(w[c] > thr)) return;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int b = tro[c], e = tro[c + 1];
  int any = 0;
  double best = -DBL_MAX;
  int arg = 0x7fffffff;
  for (int p = b + threadIdx.x; p < e; p += 256) {
    const double v = rtv[p];
    any |= (v != 0.);
    if (!alive[p]) continue;
    const double x = v * rs[tcol[p]];
    if (x > best) { best = x; arg = p; }          // p ascending per thread: first maximum kept
  }
  any = __any_sync(0xffffffffu, any);
  for (int off = 16; off >= 1; off >>= 1) {
    const double ob = __shfl_down_sync(0xffffffffu, best, off);
    const int oa = __shfl_down_sync(0xffffffffu, arg, off);
    if (oa != 0x7fffffff && (arg == 0x7fffffff || ob > best || (ob == best && oa < arg))) { best = ob; arg = oa; }
  }
  if (lane == 0) { sbest[wid] = best; sarg[wid] = arg; sany[wid] = any; }
  __syncthreads();
  if (threadIdx.x == 0) {
    int a_any = 0;
    for (int q = 0; q < 8; q++) {
      a_any |= sany[q];
      const double ob = sbest[q];
      const int oa = sarg[q];
      if (q == 0) { best = ob; arg = oa; }
      else if (oa != 0x7fffffff && (arg == 0x7fffffff || ob > best || (ob == best && oa < arg))) { best = ob; arg = oa; }
    }
    if (a_any && arg != 0x7fffffff) {
      alive[arg] = 0; rtv[arg] = 0.0;
      rv[src[arg]] = 0.0; skel[src[arg]] = 1;
      gen[tcol[arg]] = next;
    }
  }
}

// expand_support (:965-1114) for one bad row per warp: |X| ranked descending (stable: ties keep
// column order), the shortest prefix whose running sum reaches half of the row sum is taken (the
// two sums run left to right over the ranking, as sum()/cumsum() do), and the chosen columns are
// put back in ascending order.  Scratch: the row's own storage in X plus a same-sized buffer.
__global__ void __launch_bounds__(256) k_expand_rows(int nf, const double *badrow, const int *xro, int *xcol,
                                                     double *xa, int *tcol, double *tval, int *take) {
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= nf) return;
  const int lane = threadIdx.x & 31;
  if (badrow[i] == 0.) { if (lane == 0) take[i] = 0; return; }
  const int b = xro[i], len = xro[i + 1] - b;
  int *c = xcol + b, *tc = tcol + b;
  double *v = xa + b, *tv = tval + b;
  for (int k = lane; k < len; k += 32) v[k] = fabs(v[k]);
  __syncwarp();
  for (int k = lane; k < len; k += 32) {           // rank = #larger + #equal before it
    const double x = v[k];
    int r = 0;
    for (int f = 0; f < len; f++) { const double y = v[f]; r += (y > x) || (y == x && f < k); }
    tv[r] = x; tc[r] = c[k];
  }
  __syncwarp();
  int t = 0;
  if (lane == 0) {
    double tot = 0.0;
    for (int k = 0; k < len; k++) if (tv[k] != 0.) tot = tot + tv[k];
    const double half = tot * 0.5;
    int below = 0;
    double cs = 0.0;
    for (int k = 0; k < len; k++) {
      if (tv[k] != 0.) cs = cs + tv[k];
      const double s = (half != 0.) ? (1. * cs + (-1.) * half) : (1. * cs);
      if (s < 0.) below++;
    }
    t = below + 1;
    if (t > len) t = len;
    take[i] = t;
  }
  t = __shfl_sync(0xffffffffu, t, 0);
  for (int k = lane; k < t; k += 32) {             // chosen columns back in ascending order
    const int x = tc[k];
    int r = 0;
    for (int f = 0; f < t; f++) r += (tc[f] < x);
    c[r] = x;
  }
}
#endif

// find_support (:1260).  Entries leave R one per bad column and round; here they are zeroed and
// flagged dead instead of compacting R and its transpose every round (a zero adds nothing to any
// of the sums involved, so all vectors are bit-identical).
Csr find_support(const Csr &R, Csr &Rt, Buf<int> &tpos, double goal) {   // Rt = R' (values are consumed)
  StageTimer st_("find_support");
  const int nf = R.rn, nc = R.cn;
  Buf<int> src(R.nnz), skel(R.nnz), alive_t(R.nnz);
  Buf<double> rv = R.a.clone();
  skel.zero();
  fill_int(alive_t.p, R.nnz, 1);
  { int *s = src.p; const int *tp = tpos.p; parallel_for(R.nnz, [=] DEV(i64 e) { s[tp[e]] = (int)e; }); }
  Buf<double> rs(nf), tmp(nf), w(nc), w2(nc), v(nc);
  // rs = R*1 changes only in the rows that lost an entry in the round before: gen[i] is the round
  // in which row i has to be summed again (round 0: every row); the other rows keep their sum,
  // which a full recomputation would reproduce bit for bit
  Buf<int> gen(nf);
  gen.zero();
  int *genp = gen.p;
  int round = 0;
  double theta = 0.5;
  double *rvp = rv.p, *rtv = Rt.a.p, *rsp = rs.p, *wp = w.p, *w2p = w2.p, *vp = v.p;
  const int *tro = Rt.ro.p, *tcol = Rt.col.p, *srcp = src.p;
  int *skp = skel.p, *alp = alive_t.p;
  int guard = 0;
  for (;;) {
    spmv_vals(rsp, 0., nullptr, 1., R, rvp, nullptr, nullptr, genp, round);      // x = ones, rows of this round
    spmv_vals(wp, 0., nullptr, 1., Rt, rtv, rsp);
    spmv_vals(tmp.p, 0., nullptr, 1., R, rvp, wp);
    spmv_vals(w2p, 0., nullptr, 1., Rt, rtv, tmp.p);
    parallel_for(nc, [=] DEV(i64 i) { double q = w2p[i] / wp[i]; if (wp[i] == 0.) q = 0.; vp[i] = q; });
    double mv;
    max_first(vp, nc, &mv, nullptr);
    if (mv < goal) break;
    if (nf <= 1) break;   // the reference writes row index 1 here and never terminates
    while (mv <= (1 + theta) * goal) theta = theta / 2.;
    const double thr = (1 + theta) * goal;
#ifdef AMGB_EMU
    if (getenv("AMGB_FS_LOG")) {
      int nb = 0;
      for (int c = 0; c < nc; c++) if (wp[c] > thr) nb++;
      fprintf(stderr, "fs nf %d nc %d nnz %lld bad %d mv %.6g thr %.6g\n", nf, nc, (long long)R.nnz, nb, mv, thr);
    }
    parallel_for(nc, [=] DEV(i64 c) {
      double sum = 0.0;
      for (int p = tro[c]; p < tro[c + 1]; p++) sum = sum + rtv[p];
      if (!(wp[c] > thr && sum != 0.)) return;
      double maxx = -DBL_MAX;
      int arg = -1;
      for (int p = tro[c]; p < tro[c + 1]; p++) {
        if (!alp[p]) continue;
        const double x = rtv[p] * rsp[tcol[p]];
        if (x > maxx) { maxx = x; arg = p; }
      }
      if (arg < 0) return;
      alp[arg] = 0; rtv[arg] = 0.0;
      rvp[srcp[arg]] = 0.0; skp[srcp[arg]] = 1;
      genp[tcol[arg]] = round + 1;
    });
#else
    {
      Context &cx = ctx();
      if (R.nnz > 512 * (i64)nc || test_small_bins())
        k_find_support_cols_block<<<nc, 256, 0, cx.stream>>>(nc, tro, tcol, rtv, rsp, wp, thr, alp, rvp, srcp, skp, genp, round + 1);
      else
        k_find_support_cols<<<(nc + 7) / 8, 256, 0, cx.stream>>>(nc, tro, tcol, rtv, rsp, wp, thr, alp, rvp, srcp, skp, genp, round + 1);
      cx.launches++; post_launch("find_support_cols");
    }
#endif
    stage_count("find_support.rounds", 1);
    round++;
    if (++guard > 100000) throw Error(-8, "find_support: no convergence");
  }
  // Skel = sparse(skel_i, skel_j, 1): the flagged entries, in place
  Buf<int> cnt(nf + 1);
  const int *ro = R.ro.p, *col = R.col.p;
  { int *c = cnt.p; parallel_for(nf, [=] DEV(i64 i) { int k = 0; for (int j = ro[i]; j < ro[i + 1]; j++) k += skp[j]; c[i] = k; }); }
  Buf<int> mro(nf + 1);
  const i64 nnz = exclusive_scan(cnt.p, mro.p, nf);
  Csr M(nf, nc, nnz);
  M.ro = std::move(mro);
  const int *mr = M.ro.p;
  int *mcol = M.col.p;
  double *ma = M.a.p;
  parallel_for(nf, [=] DEV(i64 i) {
    int p = mr[i];
    for (int j = ro[i]; j < ro[i + 1]; j++) if (skp[j]) { mcol[p] = col[j]; ma[p] = 1.; p++; }
  });
  return M;
}

// expand_support (:907)
Csr expand_support(const Csr &Wsk, const Csr &R, Csr &Rt, Buf<int> &rtpos, const Csr &R0, double gamma) {
  StageTimer st_("expand_support(incl find)");
  const int nf = Wsk.rn, nc = Wsk.cn;
  Csr M = find_support(R, Rt, rtpos, gamma);
  trace_csr("es.M", M);
  Csr ns = mpm(1., M, 1., Wsk);
  Buf<double> badrow(nf);
  double *br = badrow.p;
  const int *nro = ns.ro.p;
  double *na = ns.a.p;
  parallel_for(nf, [=] DEV(i64 i) {
    double b = 0.;
    for (int j = nro[i]; j < nro[i + 1]; j++) if (na[j] == 2.) { b = 1.; break; }
    br[i] = b;
  });
  const i64 nbad = count_nonzero(br, nf);
  if (nbad == 0) {
    parallel_for(ns.nnz, [=] DEV(i64 e) { if (na[e] == 2.) na[e] = 1.; });
    return ns;
  }
  // X = R0 - R0.*W_skel, then per bad row: |X| sorted descending (stable), shortest prefix whose
  // running sum is not below half of the row sum (:965-1114 without the dense rank matrices)
  Csr R0W = mxmpoint(R0, Wsk);
  Csr X = mpm(1., R0, -1., R0W);
  Buf<int> take(nf + 1);
  const int *xro = X.ro.p;
  int *xcol = X.col.p, *tk = take.p;
  double *xa = X.a.p;
#ifdef AMGB_EMU
  parallel_for(nf, [=] DEV(i64 i) {
    if (br[i] == 0.) { tk[i] = 0; return; }
    const int b = xro[i], len = xro[i + 1] - b;
    int *c = xcol + b;
    double *v = xa + b;
    for (int k = 0; k < len; k++) v[k] = fabs(v[k]);
    for (int a = 1; a < len; a++) {          // stable insertion sort, descending by value
      const double tv = v[a]; const int tc = c[a];
      int j = a - 1;
      while (j >= 0 && v[j] < tv) { v[j + 1] = v[j]; c[j + 1] = c[j]; j--; }
      v[j + 1] = tv; c[j + 1] = tc;
    }
    double tot = 0.0;
    for (int k = 0; k < len; k++) if (v[k] != 0.) tot = tot + v[k];
    const double half = tot * 0.5;
    int below = 0;
    double cs = 0.0;
    for (int k = 0; k < len; k++) {
      if (v[k] != 0.) cs = cs + v[k];
      const double s = (half != 0.) ? (1. * cs + (-1.) * half) : (1. * cs);
      if (s < 0.) below++;
    }
    int t = below + 1;
    if (t > len) t = len;
    for (int a = 1; a < t; a++) {            // the chosen columns, ascending
      const int tc = c[a];
      int j = a - 1;
      while (j >= 0 && c[j] > tc) { c[j + 1] = c[j]; j--; }
      c[j + 1] = tc;
    }
    tk[i] = t;
  });
#else
  {
    Buf<int> tcolb(X.nnz);
    Buf<double> tvalb(X.nnz);
    Context &cx = ctx();
    k_expand_rows<<<(nf + 7) / 8, 256, 0, cx.stream>>>(nf, br, xro, xcol, xa, tcolb.p, tvalb.p, tk);
    cx.launches++; post_launch("expand_rows");
  }
#endif
  Buf<int> nro2(nf + 1);
  const i64 nn = exclusive_scan(take.p, nro2.p, nf);
  Csr N(nf, nc, nn);
  N.ro = std::move(nro2);
  const int *nr = N.ro.p;
  int *ncol = N.col.p;
  double *nva = N.a.p;
  parallel_for(nf, [=] DEV(i64 i) {
    const int b = xro[i];
    int p = nr[i];
    for (int k = 0; k < tk[i]; k++) { ncol[p] = xcol[b + k]; nva[p] = 1.; p++; }
  });
  Csr out = mpm(1., ns, 1., N);
  { double *oa = out.a.p; parallel_for(out.nnz, [=] DEV(i64 e) { if (oa[e] != 0.) oa[e] = 1.; }); }
  return out;
}

// s[c] = sum over the entries (i, w) of row c of Tt, i ascending, of w * M[i][c] where M holds
// (i,c); equals sum(Tt' .* M, 1) of the reference (mxmpoint :1807 + sum :1193)
#ifndef AMGB_EMU
__global__ void __launch_bounds__(256) k_col_dot_transposed(int nc, const int *tro, const int *tcol, const double *ta,
                                                            const int *mro, const int *mcol, const double *ma,
                                                            double *s) {
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c >= nc) return;
  const int lane = threadIdx.x & 31;
  const int end = tro[c + 1];
  double t = 0.0;
  for (int base = tro[c]; base < end; base += 32) {
    const int p = base + lane;
    double prod = 0.0;
    if (p < end) {
      const int i = tcol[p];
      int lo = mro[i], hi = mro[i + 1];
      const int e = hi;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (mcol[mid] < c) lo = mid + 1; else hi = mid; }
      if (lo < e && mcol[lo] == c) prod = ta[p] * ma[lo];
    }
    const int m = min(32, end - base);
    for (int l = 0; l < m; l++) t = t + __shfl_sync(0xffffffffu, prod, l);
  }
  if (lane == 0) s[c] = t;
}
#endif
void col_dot_transposed(double *s, const Csr &Tt, const Csr &M) {
  const int *tro = Tt.ro.p, *tcol = Tt.col.p, *mro = M.ro.p, *mcol = M.col.p;
  const double *ta = Tt.a.p, *ma = M.a.p;
#ifdef AMGB_EMU
  parallel_for(Tt.rn, [=] DEV(i64 c) {
    double t = 0.0;
    for (int p = tro[c]; p < tro[c + 1]; p++) {
      const int i = tcol[p];
      for (int q = mro[i]; q < mro[i + 1]; q++) if (mcol[q] == c) { t = t + ta[p] * ma[q]; break; }
    }
    s[c] = t;
  });
#else
  if (Tt.rn == 0) return;
  Context &cx = ctx();
  k_col_dot_transposed<<<(Tt.rn + 7) / 8, 256, 0, cx.stream>>>(Tt.rn, tro, tcol, ta, mro, mcol, ma, s);
  cx.launches++; post_launch("col_dot_transposed");
#endif
}

// interpolation (:598)
Csr interpolation(const Csr &Af, const Csr &Ac, const Csr &Ar, double gamma2, double tol, int *rounds_out) {
  const int nf = Af.rn, nc = Ac.cn;
  spgemm_cache_reset();
  Buf<double> Df(nf), Dfsqrti(nf), uc(nc), tmp(nf), v(nf), b(nf), Dc(nc), Dcinv(nc);
  diag_of(Df.p, Af);
  double *dfp = Df.p, *dfi = Dfsqrti.p;
  parallel_for(nf, [=] DEV(i64 i) { dfi[i] = 1. / dfp[i]; });
  fill(uc.p, nc, 1.);
  spmv(tmp.p, 0, nullptr, -1, Ar, uc.p);
  fill(b.p, nf, 1.0);
  pcg(v.p, Af, tmp.p, Df.p, 1e-16, b.p);
  trace_dev("ip.v", v.p, sizeof(double) * (size_t)nf);
  diag_of(Dc.p, Ac);
  double *dcp = Dc.p, *dci = Dcinv.p;
  parallel_for(nc, [=] DEV(i64 i) { dci[i] = 1. / dcp[i]; });
  Csr Wsk;
  {
    Csr ArD = Ar.clone();
    double *a = ArD.a.p;
    parallel_for(ArD.nnz, [=] DEV(i64 e) { a[e] = a[e] * a[e]; });
    scale_rows(ArD, dfi);
    scale_cols(ArD, dci);
    Wsk = min_skel(ArD);
  }
  Buf<double> lam(nf), alpha(nc), Dcsqrti(nc), w1(nc), w2(nc), r(nc);
  lam.zero();
  d2d(alpha.p, Dc.p, sizeof(double) * (size_t)nc);
  parallel_for(nf, [=] DEV(i64 i) { dfi[i] = sqrt(dfi[i]); });
  Csr Armt = transpose(Ar);
  { double *a = Armt.a.p; parallel_for(Armt.nnz, [=] DEV(i64 e) { a[e] = a[e] * -1.0; }); }
  double *dsq = Dcsqrti.p, *w1p = w1.p, *w2p = w2.p, *rp = r.p, *alp = alpha.p;
  Csr W;
  SkelCache sk;
  int rounds = 0;
  const std::string pfx_save = ctx().trace_prefix;
  for (;;) {
    rounds++;
    if (rounds > 100) throw Error(-9, "interpolation did not converge in 100 rounds (the reference would not terminate)");
    if (ctx().trace_on) ctx().trace_prefix = pfx_save + "r" + std::to_string(rounds) + ".";
    trace_csr("ip.Wsk", Wsk);
    Csr Wt, W0;
    solve_weights(Wt, W0, lam.p, Wsk, Af, Armt, alp, uc.p, v.p, tol, sk);
    trace_csr("ip.W0", W0);
    trace_csr("ip.Wtmp", Wt);
    Csr R0, R;
    // R0 = Af*W0 + Ar is only read by expand_support; in the last round of a level it is never
    // needed, so it is formed lazily (eagerly when tracing, to keep the trace order of the checker)
    const bool eager_R0 = ctx().trace_on;
    if (eager_R0) { StageTimer t_("ip.AfW_spgemm+mpm"); Csr AfW = spgemm(Af, W0); R0 = mpm(1., AfW, 1., Ar); }
    { StageTimer t_("ip.AfW_spgemm+mpm"); Csr AfW = spgemm(Af, Wt); R = mpm(1., AfW, 1., Ar); }
    {
      // dchat = sum(W .* (Arhat + Ar), 1): column c collects W_ic * Arr_ic over the rows i of its
      // support in ascending order -- row c of W_skel' (whose values are W' after the solve)
      // against row i of Arr, no transpose of the product needed
      Csr Arr = mpm(1.0, R, 1.0, Ar);
      col_dot_transposed(dsq, sk.Wskt, Arr);
    }
    parallel_for(nc, [=] DEV(i64 i) { double t = dsq[i] + dcp[i]; t = 1. / t; dsq[i] = sqrt(t); });
    scale_rows(R, dfi);
    { double *a = R.a.p; parallel_for(R.nnz, [=] DEV(i64 e) { a[e] = fabs(a[e]); }); }
    scale_cols(R, dsq);
    if (eager_R0) {
      scale_rows(R0, dfi);
      { double *a = R0.a.p; parallel_for(R0.nnz, [=] DEV(i64 e) { a[e] = fabs(a[e]); }); }
      scale_cols(R0, dsq);
    }
    trace_csr("ip.R", R);
    if (eager_R0) trace_csr("ip.R0", R0);
    Csr Rt;
    Buf<int> rtpos;
    {
      StageTimer t_("ip.Rt_w1w2");
      Rt = transpose(R, &rtpos);
      spmv_vals(tmp.p, 0., nullptr, 1., R, R.a.p, nullptr);      // R * ones: row sums, nothing to gather
      spmv(w1p, 0., nullptr, 1., Rt, tmp.p);
      spmv(tmp.p, 0., nullptr, 1., R, w1p);
      spmv(w2p, 0., nullptr, 1., Rt, tmp.p);
    }
    parallel_for(nc, [=] DEV(i64 i) {
      double q = w2p[i] / w1p[i];
      if (w1p[i] == 0) q = 0.;
      rp[i] = (q > gamma2) ? 1. : 0.;
    });
    const i64 nbig = count_nonzero(rp, nc);
    double w1m;
    max_first(w1p, nc, &w1m, nullptr);
    if (nbig == 0 || w1m <= gamma2) {
      Csr W0b;
      solve_weights(W, W0b, lam.p, Wsk, Af, Armt, alp, uc.p, v.p, 1e-16, sk);
      Buf<double> wuc(nf);
      spmv(wuc.p, 0., nullptr, 1., W, uc.p);
      // the reference rescales only entries whose column index equals the row index (:821-837)
      const int *wro = W.ro.p, *wcol = W.col.p;
      double *wa = W.a.p, *wu = wuc.p, *vp = v.p;
      parallel_for(nf, [=] DEV(i64 i) {
        if (wu[i] == 0.) return;
        for (int j = wro[i]; j < wro[i + 1]; j++)
          if ((int)i == wcol[j]) { const double s = vp[i] / wu[i]; wa[j] = s * wa[j]; }
      });
      break;
    }
    if (!eager_R0) {
      { StageTimer t_("ip.AfW_spgemm+mpm"); Csr AfW = spgemm(Af, W0); R0 = mpm(1., AfW, 1., Ar); }
      scale_rows(R0, dfi);
      { double *a = R0.a.p; parallel_for(R0.nnz, [=] DEV(i64 e) { a[e] = fabs(a[e]); }); }
      scale_cols(R0, dsq);
    }
    parallel_for(nc, [=] DEV(i64 i) { const double x = w2p[i] > 1e-6 ? w2p[i] : 1e-6; alp[i] = dcp[i] / x; });
    Wsk = expand_support(Wsk, R, Rt, rtpos, R0, gamma2);
    sk = SkelCache();
  }
  ctx().trace_prefix = pfx_save;
  if (rounds_out) *rounds_out = rounds;
  return W;
}

// =======================================================================================
// amg_setup (:60)
// =======================================================================================
// build_csr (:3612): assemble, then drop empty rows and the same-numbered columns
Csr build_csr(i64 n, const int *Ai, const int *Aj, const double *Av) {
  // dimensions: 1 + largest index among the non-zero entries
  Buf<int> dims(2);
  dims.zero();
  int *dp = dims.p;
  parallel_for(n, [=] DEV(i64 e) {
    if (Av[e] == 0.) return;
    // (no __CUDA_ARCH__ test inside an extended lambda: it would change the capture order
    // between the host and device passes and with it the kernel's mangled name)
    atomic_max_i32(&dp[0], Ai[e] + 1);
    atomic_max_i32(&dp[1], Aj[e] + 1);
  });
  std::vector<int> hd = dims.download();
  int rn = hd[0] > hd[1] ? hd[0] : hd[1];
  Csr T = coo_to_csr(n, Ai, Aj, Av, rn, rn);
  Buf<double> keep(rn);
  const int *ro = T.ro.p;
  double *kp = keep.p;
  parallel_for(rn, [=] DEV(i64 i) { kp[i] = (ro[i + 1] - ro[i] == 0) ? 0. : 1.; });
  if (count_nonzero(kp, rn) == rn) return T;
  return sub_mat(T, kp, kp);
}

void setup(i64 nnz, const int *dAi, const int *dAj, const double *dAv, Hierarchy &H) {
  Context &c = ctx();
  const i64 launches0 = c.launches, syncs0 = c.syncs;
  const double t_begin = now_s();
  double t0 = t_begin;
  spgemm_stats_reset();
  spmv_stats_reset();
  comm_stats_reset();
#ifndef AMGB_EMU
  cudaEvent_t ev0, ev1;
  CUDA_CHECK(cudaEventCreate(&ev0)); CUDA_CHECK(cudaEventCreate(&ev1));
  CUDA_CHECK(cudaEventRecord(ev0, c.stream));
#endif
  auto lap = [&](double &acc) { stream_sync(); const double t = now_s(); acc += t - t0; t0 = t; };
  Csr A = build_csr(nnz, dAi, dAj, dAv);
  lap(H.t.build);
  const double tol = 0.5, ctol = 0.7, itol = 1e-4;
  const double gamma2 = 1. - sqrt(1. - tol);
  hostmath::GlibcRand rng(1);
  H.lv.clear();
  H.n0 = A.rn;
  H.nullspace = 0;
  Buf<int> idl(A.rn);
  { int *p = idl.p; parallel_for(A.rn, [=] DEV(i64 i) { p[i] = (int)i + 1; }); }
  for (int level = 0;; level++) {
    if (level >= 128) throw Error(-10, "more than 128 levels");
    H.lv.emplace_back();
    Level &L = H.lv.back();
    const int rn = A.rn;
    L.n = rn;
    if (c.trace_on) c.trace_prefix = "L" + std::to_string(level) + ".";
    trace_csr("A", A);
    if (rn <= 1) {
      // a 1x1 last level whose only entry cancelled exactly (mpm drops exact zeros of a singular
      // operator) is a null space as much as a tiny entry is: project the mean, dvec = 0
      H.nullspace = (rn == 1 && (A.nnz == 0 || A.a.get(0) < 1e-9)) ? 1 : 0;
      L.A = std::move(A);
      break;
    }
    // ---- coarsen (:174-186)
    L.C.alloc(rn);
    Buf<double> vf(rn);
    L.coarsen_rounds = coarsen(L.C.p, A, ctol);
    const double *vc = L.C.p;
    { double *f = vf.p; parallel_for(rn, [=] DEV(i64 i) { f[i] = (vc[i] == 0.) ? 1. : 0.; }); }
    trace_dev("vc", vc, sizeof(double) * (size_t)rn);
    lap(H.t.coarsen);

    // ---- diagonal smoother (:188-228)
    L.Af = sub_mat(A, vf.p, vf.p);
    const int nf = L.Af.rn;
    L.nf = nf;
    L.D.alloc(nf);
    {
      const int *ro = L.Af.ro.p, *col = L.Af.col.p;
      const double *a = L.Af.a.p;
      double *D = L.D.p;
      parallel_for(nf, [=] DEV(i64 i) {
        double s = 0, dg = 0.;
        bool found = false;
        for (int j = ro[i]; j < ro[i + 1]; j++) {
          s = s + a[j] * a[j];
          if (!found && col[j] == i) { dg = a[j]; found = true; }
        }
        const double inv = 1. / s;
        D[i] = dg * inv;
      });
    }
    lap(H.t.smoother);
    if (nf >= 2) {   // (:232-285)
      Buf<double> Dh(nf);
      double *dh = Dh.p, *D = L.D.p;
      parallel_for(nf, [=] DEV(i64 i) { dh[i] = sqrt(D[i]); });
      Csr DAD = L.Af.clone();
      scale_rows(DAD, dh);
      scale_cols(DAD, dh);
      double lambda[300];
      const int k = lanczos(lambda, DAD, rng, &L.lanczos_k);
      if (k < 1) throw Error(-11, "lanczos returned no eigenvalue estimate");
      const double a = lambda[0], b = lambda[k - 1];
      const double sc = 2. / (a + b);
      parallel_for(nf, [=] DEV(i64 i) { D[i] = D[i] * sc; });
      L.rho = (b - a) / (b + a);
      L.lmin = a; L.lmax = b;
      double m, cc;
      hostmath::chebsim(&m, &cc, L.rho, gamma2);
      L.m = m;
    } else { L.rho = 0; L.m = 1; }
    trace_dev("D", L.D.p, sizeof(double) * (size_t)nf);
    lap(H.t.lanczos);

    // ---- interpolation (:302-334)
    Csr Afc = sub_mat(A, vf.p, vc);
    Csr Ac = sub_mat(A, vc, vc);
    const int nc = Ac.rn;
    L.nc = nc;
    {
      Buf<int> cflag(rn + 1), fflag(rn + 1);
      L.cpos.alloc(rn + 1); L.fpos.alloc(rn + 1);
      int *cf = cflag.p, *ff = fflag.p;
      parallel_for(rn, [=] DEV(i64 i) { cf[i] = (vc[i] == 1.) ? 1 : 0; ff[i] = (vc[i] == 1.) ? 0 : 1; });
      exclusive_scan(cflag.p, L.cpos.p, rn);
      exclusive_scan(fflag.p, L.fpos.p, rn);
      L.idc.alloc(nc); L.idf.alloc(nf);
      const int *cp = L.cpos.p, *fp = L.fpos.p, *il = idl.p;
      int *ic = L.idc.p, *jf = L.idf.p;
      parallel_for(rn, [=] DEV(i64 i) { if (cf[i]) ic[cp[i]] = il[i]; else jf[fp[i]] = il[i]; });
    }
    L.W = interpolation(L.Af, Ac, Afc, gamma2, itol, &L.interp_rounds);
    spgemm_cache_reset();
    trace_csr("W", L.W);
    lap(H.t.interp);

    // ---- Galerkin product (:336-372): AfP = Af*W + Afc;  A = W'*AfP + Acf*W + Ac
    {
      Csr AfW = spgemm(L.Af, L.W);
      L.AfP = mpm(1., AfW, 1., Afc);
    }
    trace_csr("AfP", L.AfP);
    L.Wt = transpose(L.W);
    Csr An;
    {
      Csr WtAfP = spgemm(L.Wt, L.AfP);
      Csr Acf = transpose(Afc);
      Csr AcfW = spgemm(Acf, L.W);
      Csr Atmp = mpm(1., WtAfP, 1., AcfW);
      An = mpm(1., Atmp, 1, Ac);
    }
    idl = L.idc.clone();
    L.A = std::move(A);
    A = std::move(An);
    lap(H.t.galerkin);
  }
  c.trace_prefix.clear();
#ifndef AMGB_EMU
  CUDA_CHECK(cudaEventRecord(ev1, c.stream));
#endif
  stream_sync();
  H.t.total = now_s() - t_begin;
#ifndef AMGB_EMU
  { float ms = 0; CUDA_CHECK(cudaEventElapsedTime(&ms, ev0, ev1)); H.t.device_total = ms * 1e-3; }
  cudaEventDestroy(ev0); cudaEventDestroy(ev1);
#else
  H.t.device_total = H.t.total;
#endif
  spgemm_stats_get(&H.t.spgemm, &H.t.spgemm_bytes, &H.t.spgemm_calls);
  spmv_stats_get(&H.t.spmv, &H.t.spmv_bytes, &H.t.spmv_calls);
  spmv_stats_reset();
  comm_stats_get(&H.t.comm_calls, &H.t.comm_bytes, &H.t.comm);
  spgemm_cache_reset();
  stage_report();
  H.launches = c.launches - launches0;
  H.syncs = c.syncs - syncs0;
}

}  // namespace amgb

// hostmath.h -- the scalar host-side pieces of the setup: the k x k arrowhead eigen-solver
// that Lanczos calls every iteration (k <= 299), the Chebyshev iteration count, and the
// random stream of the Lanczos start vector.  None of this touches matrix-sized data.
//
// PLAINLY: add3, ratroot, secular_root and tdeig below RESTATE sum_3, rat_root, sec_root and tdeig
// of the reference (amg_setup.c:2613-2726) operation for operation.  The hierarchy must be
// bit-identical to the reference's, and rho/m of every level come out of this iteration, so the
// order of every floating-point operation in it is fixed by the reference; only the identifiers
// differ.  It is 70 lines of scalar host code, k <= 299.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>

namespace amgb {
namespace hostmath {

// glibc rand() (TYPE_3): the reference fills the Lanczos start vector with rand()/RAND_MAX
// and never seeds it (amg_setup.c:2445-2448), i.e. the stream of seed 1.
struct GlibcRand {
  uint32_t r[34];
  int pos;
  explicit GlibcRand(uint32_t seed = 1) {
    int32_t w[34];
    w[0] = (int32_t)(seed ? seed : 1);
    for (int i = 1; i < 31; i++) {
      long long v = (16807LL * w[i - 1]) % 2147483647LL;
      if (v < 0) v += 2147483647LL;
      w[i] = (int32_t)v;
    }
    for (int i = 31; i < 34; i++) w[i] = w[i - 31];
    for (int i = 0; i < 34; i++) r[i] = (uint32_t)w[i];
    for (int i = 34; i < 344; i++) r[i % 34] = r[(i - 31) % 34] + r[(i - 3) % 34];
    pos = 34 + (344 % 34);
  }
  int32_t next() {
    const int i = pos;
    const uint32_t v = r[(i - 31) % 34] + r[(i - 3) % 34];
    r[i % 34] = v;
    pos = 34 + ((i + 1) % 34);
    return (int32_t)(v >> 1);
  }
  double uniform() { return (double)next() / (double)2147483647; }
};

// chebsim (amg_setup.c:2412)
inline void chebsim(double *m, double *c, double rho, double tol) {
  const double alpha = 0.25 * rho * rho;
  double cp = 1, gamma = 1;
  *m = 1; *c = rho;
  while (*c > tol) {
    *m += 1;
    const double d = alpha * (1 + gamma);
    gamma = d / (1 - d);
    const double cn = (1 + gamma) * rho * (*c) - gamma * cp;
    cp = *c; *c = cn;
  }
}

// sum_3 (amg_setup.c:2616)
inline double add3(double a, double b, double c) {
  if ((a >= 0 && b >= 0) || (a <= 0 && b <= 0)) return (a + b) + c;
  if ((a >= 0 && c >= 0) || (a <= 0 && c <= 0)) return (a + c) + b;
  return a + (b + c);
}
// rat_root (amg_setup.c:2627): root of -c/x + b + a x = 0 of the requested sign
inline double ratroot(double a, double b, double c, double sign) {
  const double bh = (std::fabs(b) + std::sqrt(b * b + 4 * a * c)) / 2;
  return sign * (b * sign <= 0 ? bh / a : c / bh);
}
// sec_root (amg_setup.c:2638): the root of the secular equation in (d[ri], d[ri+1])
inline double secular_root(double *y, const double *d, const double *v, int ri, int n) {
  const double eps128 = 128 * DBL_EPSILON;
  const double dl = d[ri], dr = d[ri + 1], L = dr - dl;
  double xl = L / 2, xr = -L / 2, tol = L;
  if (std::fabs(dl) > tol) tol = std::fabs(dl);
  if (std::fabs(dr) > tol) tol = std::fabs(dr);
  tol *= eps128;
  for (;;) {
    if (std::fabs(xl) == 0 || xl < 0) { *y = 0; return dl; }
    if (std::fabs(xr) == 0 || xr > 0) { *y = 0; return dr; }
    const double lam0 = std::fabs(xl) < std::fabs(xr) ? dl + xl : dr + xr;
    double al = 0, ar = 0, cl = 0, cr = 0, bln = 0, blp = 0, brn = 0, brp = 0, fn = 0, fp = 0;
    for (int i = 1; i <= ri; i++) {
      const double den = (d[i] - dl) - xl;
      double fac = v[i] / den;
      const double num = add3(d[i], -dr, -2 * xr);
      fn += v[i] * fac;
      fac *= fac;
      ar += fac;
      if (num > 0) brp += fac * num; else brn += fac * num;
      bln += fac * (d[i] - dl);
      cl += fac * xl * xl;
    }
    for (int i = ri + 1; i <= n; i++) {
      const double den = (d[i] - dr) - xr;
      double fac = v[i] / den;
      const double num = add3(d[i], -dl, -2 * xl);
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
    if (fp + fn > 0) {
      xl = ratroot(1 + al, add3(dl, blp, bln), cl, 1);
      lam = dl + xl; xr = xl - L;
    } else {
      xr = ratroot(1 + ar, add3(dr, brp, brn), cr, -1);
      lam = dr + xr; xl = xr + L;
    }
    if (std::fabs(lam - lam0) < tol) {
      double ty = 0, fac;
      for (int i = 1; i <= ri; i++) { fac = v[i] / ((d[i] - dl) - xl); ty += fac * fac; }
      for (int i = ri + 1; i <= n; i++) { fac = v[i] / ((d[i] - dr) - xr); ty += fac * fac; }
      *y = 1 / std::sqrt(1 + ty);
      return lam;
    }
  }
}
// tdeig (amg_setup.c:2712): eigenvalues of diag(d[1..n]) bordered by v[1..n], corner v[0];
// y receives the last component of every normalised eigenvector.
inline void tdeig(double *lambda, double *y, double *d, const double *v, int n) {
  double v1 = 0, lo = v[0], hi = v[0];
  for (int i = 1; i <= n; i++) {
    const double vi = std::fabs(v[i]), a = d[i] - vi, b = d[i] + vi;
    v1 += vi;
    if (a < lo) lo = a;
    if (b > hi) hi = b;
  }
  d[0] = v[0] - v1 < lo ? v[0] - v1 : lo;
  d[n + 1] = v[0] + v1 > hi ? v[0] + v1 : hi;
  for (int i = 0; i <= n; i++) lambda[i] = secular_root(&y[i], d, v, i, n);
}

}  // namespace hostmath
}  // namespace amgb

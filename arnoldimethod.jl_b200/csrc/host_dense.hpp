// host_dense.hpp - the m x m dense algebra of the Krylov-Schur restart, on the host.
//
// north_star keeps "the tiny m x m Hessenberg Schur factorisation on the host"; this is
// that host code in C++ (Float64 / ComplexF64), so that the time between two GPU sweeps
// is O(100 us) instead of an interpreter's O(0.1 s).  It follows the reference's
//   src/schurfact.jl, src/schursort.jl, src/restore_hessenberg.jl, src/eigvals.jl:6-54,
//   src/eigenvector_uppertriangular.jl, src/targets.jl, src/run.jl:197-208,394-545
// and Julia's LinearAlgebra.givensAlgorithm (LAPACK dlartg/zlartg; not in the reference
// tree).  Matrices are column-major views with 1-BASED accessors so indices read like
// the cited lines.
#pragma once

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <limits>
#include <stdexcept>
#include <vector>

namespace b2a {
namespace host {

using cplx = std::complex<double>;
constexpr double kEps = std::numeric_limits<double>::epsilon();

template <class T> struct is_cplx { static constexpr bool value = false; };
template <> struct is_cplx<cplx> { static constexpr bool value = true; };

inline double cj(double x) { return x; }
inline cplx cj(const cplx &x) { return std::conj(x); }
inline double re(double x) { return x; }
inline double re(const cplx &x) { return x.real(); }
inline double im(double) { return 0.0; }
inline double im(const cplx &x) { return x.imag(); }
inline double abs2(double x) { return x * x; }
inline double abs2(const cplx &x) { return x.real() * x.real() + x.imag() * x.imag(); }
inline bool iszero(double x) { return x == 0.0; }
inline bool iszero(const cplx &x) { return x.real() == 0.0 && x.imag() == 0.0; }

// column-major view, 1-based
template <class T> struct Mat {
  T *p;
  int rows, cols, ld;
  T &operator()(int i, int j) const { return p[(i - 1) + (int64_t)(j - 1) * ld]; }
  bool null() const { return p == nullptr; }
};

struct QRNoConvergence : std::runtime_error {
  QRNoConvergence() : std::runtime_error("QR algorithm did not converge") {}
};

// ---------------------------------------------------------------- givensAlgorithm
template <class T> struct Givens {
  double c;
  T s;
  T r;
};

inline Givens<double> givens(double f, double g) {
  const double safmn2 = std::ldexp(1.0, -485), safmx2 = std::ldexp(1.0, 485);
  if (g == 0.0) return {1.0, 0.0, f};
  if (f == 0.0) return {0.0, 1.0, g};
  double f1 = f, g1 = g, c, s, r;
  double scale = std::max(std::fabs(f1), std::fabs(g1));
  if (scale >= safmx2) {
    int count = 0;
    do {
      ++count;
      f1 *= safmn2;
      g1 *= safmn2;
      scale = std::max(std::fabs(f1), std::fabs(g1));
    } while (scale >= safmx2 && count < 20);
    r = std::sqrt(f1 * f1 + g1 * g1);
    c = f1 / r;
    s = g1 / r;
    for (int i = 0; i < count; ++i) r *= safmx2;
  } else if (scale <= safmn2) {
    int count = 0;
    do {
      ++count;
      f1 *= safmx2;
      g1 *= safmx2;
      scale = std::max(std::fabs(f1), std::fabs(g1));
    } while (scale <= safmn2);
    r = std::sqrt(f1 * f1 + g1 * g1);
    c = f1 / r;
    s = g1 / r;
    for (int i = 0; i < count; ++i) r *= safmn2;
  } else {
    r = std::sqrt(f1 * f1 + g1 * g1);
    c = f1 / r;
    s = g1 / r;
  }
  if (std::fabs(f) > std::fabs(g) && c < 0.0) {
    c = -c;
    s = -s;
    r = -r;
  }
  return {c, s, r};
}

inline Givens<cplx> givens(const cplx &f, const cplx &g) {
  const double safmn2 = std::ldexp(1.0, -485), safmx2 = std::ldexp(1.0, 485);
  const double safmin = std::numeric_limits<double>::min();
  auto abs1 = [](const cplx &z) { return std::max(std::fabs(z.real()), std::fabs(z.imag())); };
  double scale = std::max(abs1(f), abs1(g));
  cplx fs = f, gs = g;
  int count = 0;
  if (scale >= safmx2) {
    do {
      ++count;
      fs *= safmn2;
      gs *= safmn2;
      scale *= safmn2;
    } while (scale >= safmx2 && count < 20);
  } else if (scale <= safmn2) {
    if (iszero(g)) return {1.0, cplx(0.0), f};
    do {
      --count;
      fs *= safmx2;
      gs *= safmx2;
      scale *= safmx2;
    } while (scale <= safmn2);
  }
  const double f2 = abs2(fs), g2 = abs2(gs);
  if (f2 <= std::max(g2, 1.0) * safmin) {
    if (iszero(f)) {
      const double d = std::abs(gs);
      return {0.0, cplx(gs.real() / d, -gs.imag() / d), cplx(std::abs(g))};
    }
    const double f2s = std::abs(fs), g2s = std::sqrt(g2);
    const double c = f2s / g2s;
    cplx ff;
    if (abs1(f) > 1.0) {
      const double d = std::abs(f);
      ff = cplx(f.real() / d, f.imag() / d);
    } else {
      const double dr = safmx2 * f.real(), di = safmx2 * f.imag();
      const double d = std::hypot(dr, di);
      ff = cplx(dr / d, di / d);
    }
    const cplx s = ff * cplx(gs.real() / g2s, -gs.imag() / g2s);
    return {c, s, c * f + s * g};
  }
  const double f2s = std::sqrt(1.0 + g2 / f2);
  cplx r(f2s * fs.real(), f2s * fs.imag());
  const double c = 1.0 / f2s;
  const double d = f2 + g2;
  cplx s = cplx(r.real() / d, r.imag() / d) * std::conj(gs);
  if (count > 0)
    for (int i = 0; i < count; ++i) r *= safmx2;
  else if (count < 0)
    for (int i = 0; i < -count; ++i) r *= safmn2;
  return {c, s, r};
}

// ---------------------------------------------------------------------- rotations
// schurfact.jl:19-35.  lmul: rows i..; [c s; -conj(s) c].  rmul: columns, times G^H.
template <class T> struct Rot2 {
  double c;
  T s;
  int i;
};
template <class T> struct Rot3 {
  double c1;
  T s1;
  double c2;
  T s2;
  int i;
};

template <class T> inline Rot2<T> get_rotation(const T &p1, const T &p2, int i, T *nrm = nullptr) {
  auto g = givens(p1, p2);
  if (nrm) *nrm = g.r;
  return {g.c, g.s, i};
}
template <class T>
inline Rot3<T> get_rotation(const T &p1, const T &p2, const T &p3, int i, T *nrm = nullptr) {
  auto g1 = givens(p2, p3);
  auto g2 = givens(p1, g1.r);
  if (nrm) *nrm = g2.r;
  return {g1.c, g1.s, g2.c, g2.s, i};
}

template <class T> inline void lmul(const Rot2<T> &G, const Mat<T> &A, int from, int to) {
  if (A.null()) return;
  for (int j = from; j <= to; ++j) {
    const T a1 = A(G.i, j), a2 = A(G.i + 1, j);
    A(G.i, j) = G.c * a1 + G.s * a2;
    A(G.i + 1, j) = -cj(G.s) * a1 + G.c * a2;
  }
}
template <class T> inline void rmul(const Mat<T> &A, const Rot2<T> &G, int from, int to) {
  if (A.null()) return;
  for (int j = from; j <= to; ++j) {
    const T a1 = A(j, G.i), a2 = A(j, G.i + 1);
    A(j, G.i) = a1 * G.c + a2 * cj(G.s);
    A(j, G.i + 1) = a1 * -G.s + a2 * G.c;
  }
}
template <class T> inline void lmul(const Rot3<T> &G, const Mat<T> &A, int from, int to) {
  if (A.null()) return;
  for (int j = from; j <= to; ++j) {
    const T a1 = A(G.i, j), a2 = A(G.i + 1, j), a3 = A(G.i + 2, j);
    const T a2p = G.c1 * a2 + G.s1 * a3;
    const T a3p = -cj(G.s1) * a2 + G.c1 * a3;
    A(G.i, j) = G.c2 * a1 + G.s2 * a2p;
    A(G.i + 1, j) = -cj(G.s2) * a1 + G.c2 * a2p;
    A(G.i + 2, j) = a3p;
  }
}
template <class T> inline void rmul(const Mat<T> &A, const Rot3<T> &G, int from, int to) {
  if (A.null()) return;
  for (int j = from; j <= to; ++j) {
    const T a1 = A(j, G.i), a2 = A(j, G.i + 1), a3 = A(j, G.i + 2);
    const T a2p = a2 * G.c1 + a3 * cj(G.s1);
    const T a3p = a2 * -G.s1 + a3 * G.c1;
    A(j, G.i) = a1 * G.c2 + a2p * cj(G.s2);
    A(j, G.i + 1) = a1 * -G.s2 + a2p * G.c2;
    A(j, G.i + 2) = a3p;
  }
}
// whole-matrix forms (schurfact.jl:76-77)
template <class T, class R> inline void lmul(const R &G, const Mat<T> &A) { lmul(G, A, 1, A.cols); }
template <class T, class R> inline void rmul(const Mat<T> &A, const R &G) { rmul(A, G, 1, A.rows); }

template <class T> inline bool is_offdiagonal_small(const Mat<T> &H, int i, double tol = kEps) {
  return std::abs(H(i + 1, i)) <= tol * (std::abs(H(i, i)) + std::abs(H(i + 1, i + 1)));
}

// ------------------------------------------------------------- implicit QR sweeps
// schurfact.jl:251-320
template <class T, class S>
inline void single_shift_schur(const Mat<T> &H, int from, int to, const S &mu, const Mat<T> &Q) {
  const int m = H.rows, n = H.cols;
  const T p1 = H(from, from) - T(mu);
  const T p2 = H(from + 1, from);
  auto G1 = get_rotation(p1, p2, from);
  lmul(G1, H, from, n);
  rmul(H, G1, 1, std::min(from + 2, m));
  rmul(Q, G1);
  for (int i = from + 1; i <= to - 1; ++i) {
    T nrm;
    auto G = get_rotation(H(i, i - 1), H(i + 1, i - 1), i, &nrm);
    H(i, i - 1) = nrm;
    H(i + 1, i - 1) = T(0);
    lmul(G, H, i, n);
    rmul(H, G, 1, std::min(i + 2, m));
    rmul(Q, G);
  }
}

// schurfact.jl:150-249 (real only)
inline void double_shift_schur(const Mat<double> &H, int from, int to, double trace, double determinant,
                               const Mat<double> &Q) {
  const int m = H.rows, n = H.cols;
  const double H11 = H(from, from), H21 = H(from + 1, from);
  const double H12 = H(from, from + 1), H22 = H(from + 1, from + 1), H32 = H(from + 2, from + 1);
  const double p1 = H11 * H11 + H12 * H21 - trace * H11 + determinant;
  const double p2 = H21 * (H11 + H22 - trace);
  const double p3 = H32 * H21;
  auto G1 = get_rotation(p1, p2, p3, from);
  lmul(G1, H, from, n);
  rmul(H, G1, 1, std::min(from + 3, m));
  rmul(Q, G1);
  for (int i = from + 1; i <= to - 2; ++i) {
    double nrm;
    auto G = get_rotation(H(i, i - 1), H(i + 1, i - 1), H(i + 2, i - 1), i, &nrm);
    H(i, i - 1) = nrm;
    H(i + 1, i - 1) = 0.0;
    H(i + 2, i - 1) = 0.0;
    lmul(G, H, i, n);
    rmul(H, G, 1, std::min(i + 3, m));
    rmul(Q, G);
  }
  double nrm;
  auto Gn = get_rotation(H(to - 1, to - 2), H(to, to - 2), to - 1, &nrm);
  H(to - 1, to - 2) = nrm;
  H(to, to - 2) = 0.0;
  lmul(Gn, H, to - 1, n);
  rmul(H, Gn, 1, to);
  rmul(Q, Gn);
}

inline double sgn(double x) { return (x > 0.0) - (x < 0.0); }

// schurfact.jl:327-357
inline bool upper_triangular_2x2(double H11, double H12, double H21, double H22, double &c, double &s) {
  c = 1.0;
  s = 0.0;
  if (H21 == 0.0 || (H11 - H22 == 0.0 && sgn(H12) != sgn(H21))) return false;
  if (H12 == 0.0) {
    c = 0.0;
    s = 1.0;
    return true;
  }
  const double p = (H11 - H22) / 2;
  const double bcmax = std::max(std::fabs(H12), std::fabs(H21));
  const double bcmis = std::min(std::fabs(H12), std::fabs(H21)) * sgn(H12) * sgn(H21);
  const double scale = std::max(std::fabs(p), bcmax);
  const double z = (p / scale) * p + (bcmax / scale) * bcmis;
  if (z < 0.0) return false;
  const double H11_min_lam = p + std::copysign(std::sqrt(scale) * std::sqrt(z), p);
  const double nrm = std::hypot(H21, H11_min_lam);
  c = H11_min_lam / nrm;
  s = H21 / nrm;
  return true;
}

// schurfact.jl:363-388
inline bool use_single_shift(double H11, double H12, double H21, double H22, double &lam) {
  const double scale = std::fabs(H11) + std::fabs(H12) + std::fabs(H21) + std::fabs(H22);
  H11 /= scale;
  H12 /= scale;
  H21 /= scale;
  H22 /= scale;
  const double t = (H11 + H22) / 2;
  const double d = (H11 - t) * (H22 - t) - H12 * H21;
  lam = 0.0;
  if (d > 0.0) return false;
  const double sq = std::sqrt(std::fabs(d));
  const double l1 = t + sq, l2 = t - sq;
  lam = (std::fabs(H22 - l1) < std::fabs(H22 - l2) ? l1 : l2) * scale;
  return true;
}

// local_schurfact!, real arithmetic (schurfact.jl:393-487); throws on non-convergence (:406)
inline bool local_schurfact(const Mat<double> &H, int start, int to, const Mat<double> &Q,
                            double tol = kEps, int maxiter = -1) {
  if (maxiter < 0) maxiter = 100 * H.rows;
  int iter = 0;
  while (to > start) {
    if (++iter > maxiter) throw QRNoConvergence();
    int from = to;
    while (from > start) {
      if (is_offdiagonal_small(H, from - 1, tol)) {
        H(from, from - 1) = 0.0;
        break;
      }
      --from;
    }
    if (from == to) {
      --to;
      continue;
    }
    const double C11 = H(to - 1, to - 1), C12 = H(to - 1, to);
    const double C21 = H(to, to - 1), C22 = H(to, to);
    if (from + 1 == to) {
      double cs, sn;
      if (upper_triangular_2x2(C11, C12, C21, C22, cs, sn)) {
        Rot2<double> G{cs, sn, from};
        lmul(G, H, from, H.cols);
        rmul(H, G, 1, to);
        rmul(Q, G);
        H(to, to - 1) = 0.0;
      }
      to -= 2;
      continue;
    }
    double mu;
    if (use_single_shift(C11, C12, C21, C22, mu)) {
      single_shift_schur(H, from, to, mu, Q);
    } else {
      double_shift_schur(H, from, to, C11 + C22, C11 * C22 - C12 * C21, Q);
    }
  }
  return true;
}

// local_schurfact!, generic/complex (schurfact.jl:492-538); returns false on non-convergence
inline bool local_schurfact(const Mat<cplx> &H, int start, int to, const Mat<cplx> &Q, double tol = kEps,
                            int maxiter = -1) {
  if (maxiter < 0) maxiter = 100 * H.rows;
  int iter = 0;
  while (true) {
    if (++iter > maxiter) return false;
    int from = to;
    while (from > start && !is_offdiagonal_small(H, from - 1, tol)) --from;
    if (from == to) {
      if (from >= 2) H(from, from - 1) = cplx(0.0);  // guard of the latent from == 1 case
      --to;
    } else {
      const cplx H11 = H(to - 1, to - 1), H12 = H(to - 1, to);
      const cplx H21 = H(to, to - 1), H22 = H(to, to);
      const cplx d = H11 * H22 - H21 * H12;
      const cplx t = H11 + H22;
      const cplx sq = std::sqrt(t * t - 4.0 * d);
      const cplx l1 = (t + sq) / 2.0, l2 = (t - sq) / 2.0;
      const cplx lam = std::abs(H22 - l1) < std::abs(H22 - l2) ? l1 : l2;
      single_shift_schur(H, from, to, lam, Q);
    }
    if (to <= start) break;
  }
  return true;
}

// ------------------------------------------------------------ eigenvalues (eigvals.jl)
template <class T> inline cplx pair_sqrt(const T &x, const T &d) { return std::sqrt(cplx(x * x - d)); }

template <class T>
inline void copy_eigenvalues(cplx *lams, const Mat<T> &A, int first, int last, double tol = kEps) {
  int i = first;
  while (i < last) {
    if (is_offdiagonal_small(A, i, tol)) {
      lams[i - 1] = cplx(A(i, i));
      ++i;
    } else {
      const T d = A(i, i) * A(i + 1, i + 1) - A(i, i + 1) * A(i + 1, i);
      const T x = (A(i, i) + A(i + 1, i + 1)) / 2.0;
      const cplx y = pair_sqrt(x, d);
      lams[i - 1] = cplx(x) + y;
      lams[i] = cplx(x) - y;
      i += 2;
    }
  }
  if (i == last) lams[i - 1] = cplx(A(i, i));
}

template <class T> inline cplx eigenvalue(const Mat<T> &R, int i) {
  const int n = std::min(R.rows, R.cols);
  if (i == n || iszero(R(i + 1, i))) return cplx(R(i, i));
  const T d = R(i, i) * R(i + 1, i + 1) - R(i, i + 1) * R(i + 1, i);
  const T x = (R(i, i) + R(i + 1, i + 1)) / 2.0;
  return cplx(x) + pair_sqrt(x, d);
}

template <class T> inline bool is_start_of_11_block(const Mat<T> &R, int i) {
  return i == R.cols || iszero(R(i + 1, i));
}
template <class T> inline bool is_end_of_11_block(const Mat<T> &R, int i) {
  return i == 1 || iszero(R(i, i - 1));
}

// ------------------------------------------- tiny Sylvester solves (schursort.jl:61-202)
// completely pivoted LU of an N x N system (N = 2 or 4), LINPACK-style interleaved pivots.
template <class T, int N> struct SmallLU {
  T A[N][N];  // [row][col], 0-based storage
  int p[N], q[N];
  bool singular = false;

  void factor() {
    for (int i = 0; i < N; ++i) p[i] = q[i] = N;
    for (int k = 1; k <= N - 1; ++k) {
      int m = 1, n = 1;
      double maxval = 0.0;
      for (int j = k; j <= N; ++j)
        for (int i = k; i <= N; ++i)
          if (std::abs(A[i - 1][j - 1]) > maxval) {
            m = i;
            n = j;
            maxval = std::abs(A[i - 1][j - 1]);
          }
      p[k - 1] = m;
      q[k - 1] = n;
      for (int j = k; j <= N; ++j) std::swap(A[k - 1][j - 1], A[m - 1][j - 1]);
      for (int j = k; j <= N; ++j) std::swap(A[j - 1][k - 1], A[j - 1][n - 1]);
      const T Akk = A[k - 1][k - 1];
      if (iszero(Akk)) {
        singular = true;
        break;
      }
      for (int i = k + 1; i <= N; ++i) A[i - 1][k - 1] /= Akk;
      for (int j = k + 1; j <= N; ++j) {
        const T Akj = A[k - 1][j - 1];
        for (int i = k + 1; i <= N; ++i) A[i - 1][j - 1] -= A[i - 1][k - 1] * Akj;
      }
    }
    if (iszero(A[N - 1][N - 1])) singular = true;
  }

  void solve(T *x) const {
    for (int i = 1; i <= N; ++i) {
      std::swap(x[i - 1], x[p[i - 1] - 1]);
      for (int j = i + 1; j <= N; ++j) x[j - 1] -= A[j - 1][i - 1] * x[i - 1];
    }
    for (int i = N; i >= 1; --i) {
      for (int j = N; j >= i + 1; --j) x[i - 1] -= A[i - 1][j - 1] * x[j - 1];
      x[i - 1] /= A[i - 1][i - 1];
      std::swap(x[i - 1], x[q[i - 1] - 1]);
    }
  }
};

// ------------------------------------------------------ block swaps (schursort.jl:222-503)
template <class T> inline void swap22(const Mat<T> &R, int i, const Mat<T> &Q) {
  const int n = R.cols;
  const T A11 = R(i, i), A12 = R(i, i + 1), A21 = R(i + 1, i), A22 = R(i + 1, i + 1);
  const T B11 = R(i + 2, i + 2), B12 = R(i + 2, i + 3), B21 = R(i + 3, i + 2), B22 = R(i + 3, i + 3);
  SmallLU<T, 4> lu;
  const T z(0);
  T S[4][4] = {{A11 - B11, A12, -B21, z},
               {A21, A22 - B11, z, -B21},
               {-B12, z, A11 - B22, A12},
               {z, -B12, A21, A22 - B22}};
  for (int a = 0; a < 4; ++a)
    for (int b = 0; b < 4; ++b) lu.A[a][b] = S[a][b];
  lu.factor();
  if (lu.singular) return;
  // vec(C), column-major: C = R[i:i+1, i+2:i+3]
  T x[4] = {R(i, i + 2), R(i + 1, i + 2), R(i, i + 3), R(i + 1, i + 3)};
  lu.solve(x);
  const T X11 = x[0], X21 = x[1], X12 = x[2], X22 = x[3];
  const T one(1);
  auto g1 = givens(T(-X21), one);
  auto g2 = givens(T(-X11), g1.r);
  T Y22 = g1.c * -X22;
  const T Y32 = -cj(g1.s) * -X22;
  Y22 = -cj(g2.s) * -X12 + g2.c * Y22;
  auto g3 = givens(Y32, one);
  auto g4 = givens(Y22, g3.r);
  Rot3<T> G1{g1.c, g1.s, g2.c, g2.s, i};
  Rot3<T> G2{g3.c, g3.s, g4.c, g4.s, i + 1};
  lmul(G1, R, i, n);
  rmul(R, G1, 1, i + 3);
  lmul(G2, R, i, n);
  rmul(R, G2, 1, i + 3);
  R(i + 2, i) = z;
  R(i + 3, i) = z;
  R(i + 2, i + 1) = z;
  R(i + 3, i + 1) = z;
  rmul(Q, G1);
  rmul(Q, G2);
}

template <class T> inline void swap21(const Mat<T> &R, int i, const Mat<T> &Q) {
  const int n = R.cols;
  const T A11 = R(i, i), A12 = R(i, i + 1), A21 = R(i + 1, i), A22 = R(i + 1, i + 1);
  const T B11 = R(i + 2, i + 2);
  SmallLU<T, 2> lu;
  lu.A[0][0] = A11 - B11;
  lu.A[0][1] = A12;
  lu.A[1][0] = A21;
  lu.A[1][1] = A22 - B11;
  lu.factor();
  if (lu.singular) return;
  T x[2] = {R(i, i + 2), R(i + 1, i + 2)};
  lu.solve(x);
  const T one(1);
  auto g1 = givens(T(-x[1]), one);
  auto g2 = givens(T(-x[0]), g1.r);
  Rot3<T> G1{g1.c, g1.s, g2.c, g2.s, i};
  lmul(G1, R, i, n);
  rmul(R, G1, 1, i + 2);
  R(i + 1, i) = T(0);
  R(i + 2, i) = T(0);
  rmul(Q, G1);
}

template <class T> inline void swap12(const Mat<T> &R, int i, const Mat<T> &Q) {
  const int n = R.cols;
  const T A11 = R(i, i);
  const T B11 = R(i + 1, i + 1), B12 = R(i + 1, i + 2), B21 = R(i + 2, i + 1), B22 = R(i + 2, i + 2);
  SmallLU<T, 2> lu;
  lu.A[0][0] = A11 - B11;
  lu.A[0][1] = -B21;
  lu.A[1][0] = -B12;
  lu.A[1][1] = A11 - B22;
  lu.factor();
  if (lu.singular) return;
  T x[2] = {R(i, i + 1), R(i, i + 2)};
  lu.solve(x);
  const T one(1);
  auto g1 = givens(T(-x[0]), one);
  const T X22 = -cj(g1.s) * -x[1];
  auto g2 = givens(X22, one);
  Rot2<T> G1{g1.c, g1.s, i};
  Rot2<T> G2{g2.c, g2.s, i + 1};
  lmul(G1, R, i, n);
  rmul(R, G1, 1, i + 2);
  lmul(G2, R, i, n);
  rmul(R, G2, 1, i + 2);
  R(i + 2, i) = T(0);
  R(i + 2, i + 1) = T(0);
  rmul(Q, G1);
  rmul(Q, G2);
}

template <class T> inline void swap11(const Mat<T> &R, int i, const Mat<T> &Q) {
  const int n = R.cols;
  const T R11 = R(i, i), R12 = R(i, i + 1), R22 = R(i + 1, i + 1);
  auto G = get_rotation(R12, T(R22 - R11), i);
  lmul(G, R, i + 2, n);
  rmul(R, G, 1, i - 1);
  R(i, i) = R22;
  R(i + 1, i + 1) = R11;
  rmul(Q, G);
}

template <class T> inline void swap_blocks(const Mat<T> &R, int i, bool curr_11, bool next_11, const Mat<T> &Q) {
  if (curr_11) {
    if (next_11)
      swap11(R, i, Q);
    else
      swap12(R, i, Q);
  } else {
    if (next_11)
      swap21(R, i, Q);
    else
      swap22(R, i, Q);
  }
}

// schursort.jl:19-32
template <class T> inline void rotate_right(const Mat<T> &R, int from, int to, const Mat<T> &Q) {
  int i = to;
  while (i > from) {
    const bool curr_11 = is_start_of_11_block(R, i);
    const bool prev_11 = is_end_of_11_block(R, i - 1);
    const int j = prev_11 ? i - 1 : i - 2;
    swap_blocks(R, j, prev_11, curr_11, Q);
    i = j;
  }
}

// run.jl:394-457
template <class T>
inline void partition_schur_three_way(const Mat<T> &R, const Mat<T> &Q, const std::vector<int> &groups) {
  int hi = 1, mi = 1, lo = 1;
  const int len = (int)groups.size();
  while (hi <= len) {
    const int group = groups[hi - 1];
    const int blocksize = is_start_of_11_block(R, hi) ? 1 : 2;
    if (group == 3) {
      hi += blocksize;
    } else if (group == 2) {
      rotate_right(R, mi, hi, Q);
      hi += blocksize;
      mi += blocksize;
    } else {
      rotate_right(R, lo, hi, Q);
      hi += blocksize;
      mi += blocksize;
      lo += blocksize;
    }
  }
}

// ---------------------------------------------------------------- orderings (targets.jl)
inline bool isless(double a, double b) {  // Julia isless: NaN largest, -0.0 < 0.0
  if (std::isnan(a)) return false;
  if (std::isnan(b)) return true;
  if (a == 0.0 && b == 0.0) return std::signbit(a) && !std::signbit(b);
  return a < b;
}

struct Ordering {
  int which;  // b2a_which
  double key(const cplx &z) const {
    switch (which) {
      case 0: return std::abs(z);
      case 1:
      case 2: return z.real();
      default: return z.imag();
    }
  }
  bool reverse() const { return which == 0 || which == 1 || which == 3; }
  bool lt(const cplx &a, const cplx &b) const {
    return reverse() ? isless(key(b), key(a)) : isless(key(a), key(b));
  }
};

// sort!(ord, QuickSort, OrderPerm(lams, ordering)) - run.jl:289, targets.jl:61-67
inline void sort_perm(std::vector<int> &ord, const cplx *lams, const Ordering &o) {
  std::stable_sort(ord.begin(), ord.end(), [&](int i, int j) {
    const cplx &a = lams[i - 1], &b = lams[j - 1];
    if (o.lt(a, b)) return true;
    if (o.lt(b, a)) return false;
    return i < j;
  });
}

// run.jl:465-502
template <class T> inline void sortschur(const Mat<T> &R, const Mat<T> &Q, int to, const Ordering &o) {
  if (to <= 1) return;
  int next_idx = 1;
  while (next_idx <= to) {
    int curr_idx = next_idx;
    const int curr_size = is_start_of_11_block(R, curr_idx) ? 1 : 2;
    const cplx curr_lam = eigenvalue(R, curr_idx);
    while (curr_idx > 1) {
      const int prev_size = is_end_of_11_block(R, curr_idx - 1) ? 1 : 2;
      const int prev_idx = curr_idx - prev_size;
      const cplx prev_lam = eigenvalue(R, prev_idx);
      if (!o.lt(curr_lam, prev_lam)) break;
      swap_blocks(R, prev_idx, prev_size == 1, curr_size == 1, Q);
      curr_idx -= prev_size;
    }
    next_idx += curr_size;
  }
}

// ------------------------------------------------- reflectors (restore_hessenberg.jl)
template <class T> struct Reflector {
  std::vector<T> vec;
  int offset = 1, len = 0;
  T tau = T(0);
  explicit Reflector(int max_len) : vec(max_len) {}
};

// reflector!(y, k) -> tau'  (restore_hessenberg.jl:16-45)
template <class T> inline T make_reflector(T *y, int k) {
  double xnrm = 0.0;
  for (int idx = 1; idx <= k - 1; ++idx) xnrm += abs2(y[idx - 1]);
  T alpha = y[k - 1];
  if (xnrm == 0.0 && im(alpha) == 0.0) return T(0);
  xnrm = std::sqrt(xnrm);
  const double beta = -std::copysign(std::hypot(std::abs(alpha), xnrm), re(alpha));
  const T tau = (T(beta) - alpha) / beta;
  alpha = T(1) / (alpha - T(beta));
  for (int i = 1; i <= k - 1; ++i) y[i - 1] *= alpha;
  y[k - 1] = T(beta);
  return cj(tau);
}

// restore_hessenberg.jl:138-159
template <class T> inline void lmul(const Reflector<T> &G, const Mat<T> &H, int from, int to) {
  if (iszero(G.tau)) return;
  const int len = G.len, off = G.offset;
  for (int col = from; col <= to; ++col) {
    T dot(0);
    for (int i = 1; i <= len - 1; ++i) dot += cj(G.vec[i - 1]) * H(i + off - 1, col);
    dot += H(len + off - 1, col);
    dot *= G.tau;
    for (int i = 1; i <= len - 1; ++i) H(i + off - 1, col) -= dot * G.vec[i - 1];
    H(len + off - 1, col) -= dot;
  }
}
// restore_hessenberg.jl:161-182
template <class T> inline void rmul(const Mat<T> &H, const Reflector<T> &G, int from, int to) {
  if (iszero(G.tau)) return;
  const int len = G.len, off = G.offset;
  for (int row = from; row <= to; ++row) {
    T dot(0);
    for (int i = 1; i <= len - 1; ++i) dot += H(row, i + off - 1) * G.vec[i - 1];
    dot += H(row, off + len - 1);
    dot *= cj(G.tau);
    for (int i = 1; i <= len - 1; ++i) H(row, i + off - 1) -= dot * cj(G.vec[i - 1]);
    H(row, off + len - 1) -= dot;
  }
}

// restore_arnoldi! (restore_hessenberg.jl:75-134)
template <class T>
inline void restore_arnoldi(const Mat<T> &H, int from, int to, const Mat<T> &Q, Reflector<T> &G) {
  if (!(from < to)) return;
  const int m = H.rows, n = H.cols;
  T nrm = Q(n, from);
  for (int i = from; i <= to - 1; ++i) {
    auto g = givens(Q(n, i + 1), nrm);
    nrm = g.r;
    Rot2<T> rot{g.c, -g.s, i};
    rmul(H, rot, 1, std::min(i + 2, to));
    lmul(rot, H, 1, to);
    rmul(Q, rot, 1, n);
  }
  H(to + 1, to) = Q(Q.rows, to) * H(m, n);
  G.offset = from;
  for (int i = to - from; i >= 2; --i) {
    G.len = i;
    const int row = from + i;
    for (int j = 1; j <= i; ++j) G.vec[j - 1] = cj(H(row, j + from - 1));
    G.tau = make_reflector(G.vec.data(), i);
    rmul(H, G, 1, row - 1);
    for (int j = 1; j <= i - 1; ++j) H(row, j + from - 1) = T(0);
    H(row, i - 1 + from) = cj(G.vec[i - 1]);
    lmul(G, H, from, to);
    rmul(Q, G, 1, n);
  }
}

// ------------------------- eigenvectors of (quasi) triangular R (eigenvector_uppertriangular.jl)
inline void shifted_backward_sub(cplx *x, const Mat<double> &R, const cplx &lam, int k) {
  while (k > 0) {
    if (k > 1 && R(k, k - 1) != 0.0) {
      const cplx R11 = R(k - 1, k - 1) - lam, R22 = R(k, k) - lam;
      const double R12 = R(k - 1, k), R21 = R(k, k - 1);
      const cplx det = R11 * R22 - R21 * R12;
      const cplx a1 = (R22 * x[k - 2] - R12 * x[k - 1]) / det;
      const cplx a2 = (-R21 * x[k - 2] + R11 * x[k - 1]) / det;
      x[k - 2] = a1;
      x[k - 1] = a2;
      for (int i = 1; i <= k - 2; ++i) x[i - 1] -= R(i, k - 1) * x[k - 2] + R(i, k) * x[k - 1];
      k -= 2;
    } else {
      const cplx sigma = R(k, k) - lam;
      if (iszero(sigma)) {
        x[k - 1] = sigma;
      } else {
        x[k - 1] /= sigma;
        for (int i = 1; i <= k - 1; ++i) x[i - 1] -= R(i, k) * x[k - 1];
      }
      k -= 1;
    }
  }
}
inline void shifted_backward_sub(cplx *x, const Mat<cplx> &R, const cplx &lam, int k) {
  while (k > 0) {
    const cplx sigma = R(k, k) - lam;
    if (iszero(sigma)) {
      x[k - 1] = sigma;
    } else {
      x[k - 1] /= sigma;
      for (int i = 1; i <= k - 1; ++i) x[i - 1] -= R(i, k) * x[k - 1];
    }
    k -= 1;
  }
}

inline void normalize_prefix(cplx *x, int j) {
  double nrm = 0.0;
  for (int k = 1; k <= j; ++k) nrm += abs2(x[k - 1]);
  const double scale = 1.0 / std::sqrt(nrm);
  for (int k = 1; k <= j; ++k) x[k - 1] *= scale;
}

// collect_eigen! (eigenvector_uppertriangular.jl:76-128 real, :130-154 generic)
inline int collect_eigen(cplx *x, const Mat<double> &R, int j) {
  const int n = R.cols;
  if (j < n && R(j + 1, j) != 0.0) j += 1;
  if (j > 1 && R(j, j - 1) != 0.0) {
    const double R11 = R(j - 1, j - 1), R21 = R(j, j - 1), R12 = R(j - 1, j), R22 = R(j, j);
    const double det = R11 * R22 - R21 * R12;
    const double tr = R11 + R22;
    const cplx lam = (tr + std::sqrt(cplx(tr * tr - 4 * det))) / 2.0;
    x[j - 2] = -R12 / (R11 - lam);
    x[j - 1] = 1.0;
    for (int i = 1; i <= j - 2; ++i) x[i - 1] = -R(i, j - 1) * x[j - 2] - R(i, j);
    shifted_backward_sub(x, R, lam, j - 2);
  } else {
    const cplx lam = R(j, j);
    x[j - 1] = 1.0;
    for (int i = 1; i <= j - 1; ++i) x[i - 1] = -R(i, j);
    shifted_backward_sub(x, R, lam, j - 1);
  }
  normalize_prefix(x, j);
  return j;
}
inline int collect_eigen(cplx *x, const Mat<cplx> &R, int j) {
  const cplx lam = R(j, j);
  x[j - 1] = 1.0;
  for (int i = 1; i <= j - 1; ++i) x[i - 1] = -R(i, j);
  shifted_backward_sub(x, R, lam, j - 1);
  normalize_prefix(x, j);
  return j;
}

// copy_residuals! (run.jl:524-545)
template <class T>
inline void copy_residuals(double *rs, const Mat<T> &H, const Mat<T> &Q, const T &h_last, cplx *x, int first,
                           int last) {
  const int m = H.cols;
  for (int i = 0; i < m; ++i) rs[i] = 0.0;
  for (int i = first; i <= last; ++i) {
    for (int t = 0; t < m; ++t) x[t] = cplx(0.0);
    const int len = collect_eigen(x, H, i);
    cplx tmp(0.0);
    for (int j = 1; j <= len; ++j) tmp += cplx(Q(m, j)) * x[j - 1];
    rs[i - 1] = std::abs(tmp * cplx(h_last));
  }
}

// include_conjugate_pair (run.jl:510-517)
template <class T> inline int include_conjugate_pair(const cplx *lams, const std::vector<int> &ord, int i) {
  if (is_cplx<T>::value) return i;
  if (i >= (int)ord.size()) return i;
  const cplx l1 = lams[ord[i - 1] - 1], l2 = lams[ord[i] - 1];
  return (l1.imag() != 0.0 && std::conj(l1) == l2) ? i + 1 : i;
}

template <class T> inline double frobenius(const Mat<T> &H) {
  // norm(H) (run.jl:292): scaled 2-norm of all entries
  double scale = 0.0, ssq = 1.0;
  for (int j = 1; j <= H.cols; ++j)
    for (int i = 1; i <= H.rows; ++i) {
      const double a = std::abs(H(i, j));
      if (a != 0.0) {
        if (scale < a) {
          ssq = 1.0 + ssq * (scale / a) * (scale / a);
          scale = a;
        } else {
          ssq += (a / scale) * (a / scale);
        }
      }
    }
  return scale * std::sqrt(ssq);
}

// --------------------------------------------------------------- one restart (run.jl:278-360)
template <class T> struct RestartScratch {
  std::vector<cplx> x, lams;
  std::vector<double> rs;
  std::vector<int> ord, groups;
  Reflector<T> G;
  explicit RestartScratch(int maxdim)
      : x(maxdim), lams(maxdim), rs(maxdim), ord(maxdim), groups(maxdim, 0), G(maxdim) {}
};

struct RestartPlan {
  int k, purge, nlock, effective_nev;
};

// H: (maxdim+1) x maxdim, Q: maxdim x maxdim.  Mutates both.  Throws QRNoConvergence (real T).
template <class T>
inline RestartPlan restart_decision(const Mat<T> &H, const Mat<T> &Q, int maxdim, int mindim, int nev,
                                    double tol, const Ordering &ordering, int active,
                                    RestartScratch<T> &S) {
  for (int j = 1; j <= maxdim; ++j)
    for (int i = 1; i <= maxdim; ++i) Q(i, j) = (i == j) ? T(1) : T(0);  // run.jl:278

  Mat<T> Hsq{H.p, maxdim, H.cols, H.ld};  // view(H, OneTo(maxdim), :)
  local_schurfact(Hsq, active, maxdim, Q);  // run.jl:281 (generic version's false is ignored)

  for (int i = 0; i < maxdim; ++i) S.ord[i] = i + 1;  // run.jl:284
  copy_eigenvalues(S.lams.data(), H, 1, maxdim);       // run.jl:285
  copy_residuals(S.rs.data(), H, Q, H(maxdim + 1, maxdim), S.x.data(), active, maxdim);  // run.jl:286
  sort_perm(S.ord, S.lams.data(), ordering);           // run.jl:289
  const double H_frob = frobenius(H);                  // run.jl:292

  auto isconverged = [&](int i) {  // run.jl:206-208
    return S.rs[i - 1] <= std::max(kEps * H_frob, tol * std::abs(S.lams[i - 1]));
  };

  const int effective_nev = include_conjugate_pair<T>(S.lams.data(), S.ord, nev);  // run.jl:298

  int nlock = 0;
  for (int i = 1; i <= effective_nev; ++i) {  // run.jl:301-308
    if (isconverged(S.ord[i - 1])) {
      S.groups[S.ord[i - 1] - 1] = 1;
      ++nlock;
    } else {
      S.groups[S.ord[i - 1] - 1] = 2;
    }
  }

  const int ideal_size = std::min(nlock + mindim, (mindim + maxdim) / 2);  // run.jl:316
  int k = effective_nev;
  int i = effective_nev + 1;
  while (i <= maxdim) {  // run.jl:320-339
    const bool is_pair = include_conjugate_pair<T>(S.lams.data(), S.ord, i) == i + 1;
    const int num = is_pair ? 2 : 1;
    int group;
    if (k < ideal_size && !isconverged(S.ord[i - 1])) {
      group = 2;
      k += num;
    } else {
      group = 3;
    }
    if (is_pair) {
      S.groups[S.ord[i - 1] - 1] = group;
      S.groups[S.ord[i] - 1] = group;
      i += 2;
    } else {
      S.groups[S.ord[i - 1] - 1] = group;
      i += 1;
    }
  }

  int purge = 1;  // run.jl:350-353
  while (purge < active && S.groups[purge - 1] == 1) ++purge;

  partition_schur_three_way(H, Q, S.groups);  // run.jl:355
  restore_arnoldi(H, nlock + 1, k, Q, S.G);   // run.jl:360
  return {k, purge, nlock, effective_nev};
}

}  // namespace host
}  // namespace b2a

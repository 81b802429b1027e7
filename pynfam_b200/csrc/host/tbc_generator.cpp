// Full-FAM two-body-current field of the Gamow-Teller operator: the Yukawa (pion-exchange) part of the effective
// one-body field  Gamma_ac = sum_db J_abcd rho_db  in the HO basis, i.e. what the reference computes once per
// (nucleus, operator, K) and caches in <name>.tbc.
//
// Reference: exes/pnfam/pnfam_extfield_2bc.f90:26-465 (effective_2bc_extfield, the production branch debug = 0),
//            exes/pnfam/pnfam_type_extfield_2bc.f90:47-84, 182-398, 750-825, 869-1463 (LEC prefactors, z / radial
//            spatial components, spin-isospin contraction), exes/pnfam/pnfam_spatial_mtxels.f90:78-450, 609-927
//            (Gaussian fit of the Yukawa function, 1D matrix elements with derivatives, polar -> Cartesian expansion),
//            hfbtho_gogny.f90:651-686, 884-915, 991-1045, 1246-1352 (T_z, C polar->Cartesian, 1D Gaussian element),
//            hfbtho_solver.f90:1822-1856 (density matrix rk in the HO basis), :3270-3296 (the (Omega, n_r, Lambda, s)
//            classes with their n_z lists).
//
// Same sums, different organisation.  The reference loops (n_z^a, n_z^c) outermost and re-evaluates the radial
// two-body elements -- each a four-fold sum over Cartesian quanta -- for every one of the (N_z+1)^2 combinations; here
//   1. all 1D tables (seven derivative kinds x six Gaussians) are built once,
//   2. the z contraction  W[q][dir|exc][(D,B)][comp][g](z_a, z_c) = sum_{z_b z_d} J^z_comp(g; z_a z_b z_c z_d) rho^q_db
//      is made once per (z_a, z_c) for every pair of radial classes D, B (chunked over z_a under a memory bound),
//   3. the radial elements J^r(g; A B C D) are made once per pair of radial classes (A, C) = (n_r, Lambda) of the bra and
//      the ket and serve every (z_a, z_c, s_a, s_c) of that pair,
//   4. each radial element is an O(N^2) sum over an intermediate built once per (A, C, derivative kind, Gaussian)
//      (build_q below) instead of the O(N^4) Cartesian sum; exchange elements use tables with the last two indices
//      exchanged; the complex element of the momentum terms splits the same way,
//   5. the spin-isospin contraction is a sparse list of linear coefficients probed from the literal spin functions.
// OpenMP over the (A, C) pairs.  PNFAM_B200_TBC_LITERAL_RADIAL=1 evaluates step 4 with the reference's literal sums
// (self-check).  Entry points: generate_two_body_current_field(TbcProblem) on plain arrays (C ABI:
// pnfam_host_effective_2bc_extfield) and the set-up's front door that takes the density matrices from the HFB solution.
//
// Scope: gamma (particle-hole) part -- the pairing part (3rd digit of the mode 2, 3) is refused upstream exactly where the
// reference says "not yet operational"; even and blocked (odd, equal filling) nuclei, zero and finite temperature.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <stdexcept>

#include <omp.h>
#include <unistd.h>

#include "fam_setup.hpp"

namespace pnfam {
namespace {

constexpr int NG = 6;          // Gaussians of the Yukawa fit (pnfam_spatial_mtxels.f90:52-55)
constexpr int NKIND = 7;       // (d, p): derivatives on the Gaussian, momentum on particle 1 / 2
enum Kind { G00 = 0, G10, G20, G01, G02, G11, G12 };
constexpr int KIND_D[NKIND] = {0, 1, 2, 0, 0, 1, 1};
constexpr int KIND_P[NKIND] = {0, 0, 0, 1, 2, 1, 2};
// the 14 spatial components of type extfield_2bc
enum Comp { RKP = 0, RKZ, RKM, RPZ, RPM, RZM, RPZ2, RZM2, R0P, R0Z, R0M, RP0, RZ0, RM0, NCOMP };

constexpr double HBARC = 197.3269718, FPI = 92.4, MN = 939.0, MPI = 138.04;   // pnfam_constants.f90:44-67
const double PI = 3.141592653589793238462643;

struct Fact {
  double a[171];
  Fact() { a[0] = 1.0; for (int i = 1; i <= 170; i++) a[i] = (double)i * a[i - 1]; }
  double operator()(int n) const {
    if (n < 0) throw std::runtime_error("negative integer in factrl");
    return n <= 170 ? a[n] : std::exp(std::lgamma(n + 1.0));
  }
};
const Fact factrl;

double binomialco(int m, int n) {   // hfbtho_gogny.f90:1321-1352
  if (n < 0 || n > m) return 0.0;
  if (n == 0 || n == m) return 1.0;
  if (n == 1 || n == m - 1) return (double)m;
  if (m <= 170) return n <= m / 2 ? (factrl(m) / factrl(m - n)) / factrl(n) : (factrl(m) / factrl(n)) / factrl(m - n);
  return std::exp(std::lgamma(m + 1.0) - std::lgamma(n + 1.0) - std::lgamma(m - n + 1.0));
}

double hypergeom2f1(int a, int b, double c, double x) {   // hfbtho_gogny.f90:1286-1311 (a, b <= 0)
  const int d = std::abs(std::max(a, b)), e = std::abs(std::min(a, b));
  double sum = 0.0, xi = 1.0;
  for (int i = 0; i <= d; i++) {
    double upf = 1.0;
    for (int j = 0; j < i; j++) upf *= (c + (double)j);
    sum += binomialco(d, i) * binomialco(e, i) * xi * factrl(i) / upf;
    xi *= x;
  }
  return sum;
}

struct Tables {
  int nzx = 0, nrx = 0, nlx = 0, nrlx = 0, nox = 0;
  int nbig = 0;    // largest 1D quantum number of the plain Gaussian tables (two derivatives beyond the basis)
  int nsm = 0;     // largest 1D quantum number of the derivative tables
  double bz = 0, bp = 0;
  double mu[NG], pf[NG];
  std::vector<double> tz;              // T_z(n1, n2, n)
  std::vector<double> cp2c;            // C(n, k, ny)
  int nsh = 0;
  std::vector<double> mez, mer;        // [g][i][j][k][l] over 0..nbig: z (with the fit prefactor) and perpendicular
  std::vector<double> zk[NKIND];       // [g][i][j][k][l] over 0..nzx
  std::vector<double> rk[NKIND];       // [g][i][j][k][l] over 0..nsm
  std::vector<double> rkT[NKIND];      // the same with the last two indices exchanged (exchange elements)
  std::vector<double> rkg[NKIND], rkgT[NKIND];   // [i][j][k][l][g] copies of rk / rkT (Gaussian index fastest) for build_q

  size_t ibig(int g, int i, int j, int k, int l) const {
    const size_t n = nbig + 1;
    return ((((size_t)g * n + i) * n + j) * n + k) * n + l;
  }
  size_t iz(int g, int i, int j, int k, int l) const {
    const size_t n = nzx + 1;
    return ((((size_t)g * n + i) * n + j) * n + k) * n + l;
  }
  size_t ir(int g, int i, int j, int k, int l) const {
    const size_t n = nsm + 1;
    return ((((size_t)g * n + i) * n + j) * n + k) * n + l;
  }
  double T(int n1, int n2, int n) const {
    const size_t d = nbig + 1;
    return tz[((size_t)n1 * d + n2) * (2 * nbig + 1) + n];
  }
  double C(int n, int k, int ny) const {
    return cp2c[((size_t)n * (4 * nsh + 1) + (k + 2 * nsh)) * (2 * nsh + 1) + ny];
  }
};

// hfbtho_gogny.f90:991-1045 with GOGNY_HYPER = 1 (the build setting of the reference, src/Makefile:46)
double matrix_element_z(const Tables& t, int ni, int nj, int nk, int nl, double mu, double b) {
  int ini, inj, ink, inl;
  if (std::min(nj, nl) <= std::min(ni, nk)) { ini = ni; inj = nj; ink = nk; inl = nl; }
  else { ini = nj; inj = ni; ink = nl; inl = nk; }
  const double z = 1.0 + mu * mu / (2.0 * b * b);
  double vz = 0.0;
  for (int nz = std::abs(inj - inl); nz <= inj + inl; nz += 2) {
    if ((ini + ink + nz) % 2 != 0) break;
    const double xi = (ini + ink + nz + 1) * 0.5;
    const double fbar = std::tgamma(xi - ini) * std::tgamma(xi - ink) * std::tgamma(xi - nz) /
                        (std::pow(z, xi) * std::sqrt(factrl(ini) * factrl(ink) * factrl(nz))) *
                        hypergeom2f1(-ini, -ink, nz + 1 - xi, 1.0 - z);
    vz += t.T(inj, inl, nz) * fbar;
  }
  return mu / (std::sqrt(2.0 * PI * PI * PI) * b) * vz;
}

// derivative of a 1D HO function as a combination of HO functions (pnfam_spatial_mtxels.f90:865-927)
struct HoDer { int n[3]; double c[3]; int cnt; };
HoDer ho_derivative(int n, int d, double b) {
  const double p1 = 1.0 / (b * std::sqrt(2.0)), p2 = 1.0 / (b * b * 2.0);
  HoDer h{{-1, -1, -1}, {0, 0, 0}, 1};
  if (d == 0) { h.c[0] = 1.0; h.n[0] = n; }
  else if (d == 1) {
    h.c[0] = -std::sqrt((double)(n + 1)) * p1; h.n[0] = n + 1;
    if (n - 1 >= 0) { h.c[1] = std::sqrt((double)n) * p1; h.n[1] = n - 1; h.cnt = 2; }
  } else {
    h.cnt = 2;
    h.c[0] = std::sqrt((double)((n + 1) * (n + 2))) * p2; h.n[0] = n + 2;
    h.c[1] = -(2.0 * n + 1.0) * p2; h.n[1] = n;
    if (n - 2 >= 0) { h.c[2] = std::sqrt((double)(n * (n - 1))) * p2; h.n[2] = n - 2; h.cnt = 3; }
  }
  return h;
}

// pnfam_spatial_mtxels.f90:798-855
double me1d_dho(const Tables& t, const std::vector<double>& me, double b, int g, int ni, int nj, int nk, int nl, int di, int dj,
                int dk, int dl) {
  const HoDer hi = ho_derivative(ni, di, b), hj = ho_derivative(nj, dj, b), hk = ho_derivative(nk, dk, b),
              hl = ho_derivative(nl, dl, b);
  double v = 0.0;
  for (int l = 0; l < hl.cnt; l++)
    for (int k = 0; k < hk.cnt; k++)
      for (int j = 0; j < hj.cnt; j++)
        for (int i = 0; i < hi.cnt; i++) {
          if ((hi.n[i] + hj.n[j] + hk.n[k] + hl.n[l]) % 2 != 0) continue;
          v += me[t.ibig(g, hi.n[i], hj.n[j], hk.n[k], hl.n[l])] * hi.c[i] * hj.c[j] * hk.c[k] * hl.c[l];
        }
  return v;
}

// pnfam_spatial_mtxels.f90:703-787
double me1d(const Tables& t, bool z, int g, int ni, int nj, int nk, int nl, int d, int p) {
  const std::vector<double>& me = z ? t.mez : t.mer;
  const double b = z ? t.bz : t.bp;
  if (p == 0) { if ((ni + nj + nk + nl + d) % 2 != 0) return 0.0; }
  else {
    if ((ni + nj + nk + nl + d + 1) % 2 != 0) return 0.0;
    if (p == 1 && ni == nk) return 0.0;
    if (p == 2 && nj == nl) return 0.0;
  }
  auto D = [&](int a, int bb, int c, int e) { return me1d_dho(t, me, b, g, ni, nj, nk, nl, a, bb, c, e); };
  if (d == 0 && p == 0) return me[t.ibig(g, ni, nj, nk, nl)];
  if (d == 1 && p == 0) return -D(1, 0, 0, 0) - D(0, 0, 1, 0);
  if (d == 2 && p == 0) return D(2, 0, 0, 0) + D(0, 0, 2, 0) + D(1, 0, 1, 0) * 2.0;
  if (d == 1 && p == 1) return -D(0, 0, 2, 0) + D(2, 0, 0, 0);
  if (d == 1 && p == 2) return D(0, 0, 0, 2) - D(0, 2, 0, 0);
  if (d == 0 && p == 1) return D(0, 0, 1, 0) - D(1, 0, 0, 0);
  if (d == 0 && p == 2) return D(0, 0, 0, 1) - D(0, 1, 0, 0);
  throw std::runtime_error("two-body currents: derivative combination not implemented");
}

Tables build_tables(const TbcProblem& s, bool use_p) {
  Tables t;
  t.bz = s.bz; t.bp = s.bp;
  for (int i = 0; i < s.nt; i++) {
    t.nzx = std::max(t.nzx, s.nz[i]); t.nrx = std::max(t.nrx, s.nr[i]); t.nlx = std::max(t.nlx, s.nl[i]);
    t.nrlx = std::max(t.nrlx, 2 * s.nr[i] + s.nl[i]);
    t.nox = std::max(t.nox, s.nl[i] + (s.ns[i] + 1) / 2 - 1);
  }
  // prep_gaussian (pnfam_spatial_mtxels.f90:100-108)
  const double a[NG] = {34.0, 6.60, 1.44, 0.38, 0.15, 0.13};
  const double pf[NG] = {6.79, 2.41, 0.786, 0.241, -0.062, 0.078};
  for (int i = 0; i < NG; i++) {
    t.mu[i] = (1.0 / std::sqrt(a[i])) / (MPI / HBARC);
    t.pf[i] = pf[i] * (MPI / HBARC) * 0.25 / PI;
  }
  t.nbig = std::max(t.nzx + 2, std::max(2 * (t.nrx + 1), t.nlx + 2));
  t.nsm = std::max(t.nzx, std::max(2 * t.nrx, t.nlx));
  // calculateTz
  {
    const int nz = t.nbig;
    t.tz.assign((size_t)(nz + 1) * (nz + 1) * (2 * nz + 1), 0.0);
    for (int n1 = 0; n1 <= nz; n1++)
      for (int n2 = n1; n2 <= nz; n2++)
        for (int n = std::abs(n2 - n1); n <= n2 + n1; n++) {
          if (n % 2 != (n2 + n1) % 2) continue;
          const double v = (std::sqrt(factrl(n2)) / factrl((-n1 + n2 + n) / 2)) * (std::sqrt(factrl(n)) / factrl((n1 - n2 + n) / 2)) *
                           (std::sqrt(factrl(n1)) / factrl((n1 + n2 - n) / 2));
          t.tz[((size_t)n1 * (nz + 1) + n2) * (2 * nz + 1) + n] = v;
          t.tz[((size_t)n2 * (nz + 1) + n1) * (2 * nz + 1) + n] = v;
        }
  }
  // calculateCpolar2cartesian
  {
    const int nsh = std::max(2 * (t.nrx + 1), t.nlx + 2);
    t.nsh = nsh;
    t.cp2c.assign((size_t)(nsh + 1) * (4 * nsh + 1) * (2 * nsh + 1), 0.0);
    for (int n = 0; n <= nsh; n++)
      for (int k = -2 * nsh; k <= 2 * nsh; k++)
        for (int ny = 0; ny <= 2 * nsh; ny++) {
          const int ak = std::abs(k);
          if (ak > 2 * nsh - 2 * n) continue;
          if (ny > 2 * n + ak) continue;
          const int nx = 2 * n + ak - ny;
          const double A = ((n % 2) ? -1.0 : 1.0) * std::pow(2.0, -n - ak * 0.5) * std::sqrt(factrl(n + ak)) *
                           (std::sqrt(factrl(n)) / (std::sqrt(factrl(nx)) * std::sqrt(factrl(ny))));
          double xsum = 0.0;
          const int qmax = std::min(ny, n + (ak - k) / 2);
          for (int q = 0; q <= qmax; q++)
            xsum += binomialco(nx, n - q + (ak - k) / 2) * binomialco(ny, q) * (((ny - q) % 2) ? -1.0 : 1.0);
          t.cp2c[((size_t)n * (4 * nsh + 1) + (k + 2 * nsh)) * (2 * nsh + 1) + ny] = A * xsum;
        }
  }
  // calculateME1D: plain Gaussian elements
  {
    const int n = t.nbig;
    const size_t d = n + 1, tot = (size_t)NG * d * d * d * d;
    t.mez.assign(tot, 0.0); t.mer.assign(tot, 0.0);
#pragma omp parallel for collapse(2) schedule(dynamic)
    for (int ni = 0; ni <= n; ni++)
      for (int nj = 0; nj <= n; nj++)
        for (int nk = 0; nk <= n; nk++)
          for (int nl = 0; nl <= n; nl++) {
            if ((ni + nj + nk + nl) % 2 != 0) continue;
            for (int g = 0; g < NG; g++) {
              t.mez[t.ibig(g, ni, nj, nk, nl)] = matrix_element_z(t, ni, nj, nk, nl, t.mu[g], t.bz) * t.pf[g];
              t.mer[t.ibig(g, ni, nj, nk, nl)] = matrix_element_z(t, ni, nj, nk, nl, t.mu[g], t.bp);
            }
          }
  }
  // derivative kinds
  for (int kd = 0; kd < NKIND; kd++) {
    if (!use_p && KIND_P[kd] != 0) continue;
    {
      const int n = t.nzx;
      const size_t d = n + 1;
      t.zk[kd].assign((size_t)NG * d * d * d * d, 0.0);
#pragma omp parallel for collapse(2) schedule(dynamic)
      for (int ni = 0; ni <= n; ni++)
        for (int nj = 0; nj <= n; nj++)
          for (int nk = 0; nk <= n; nk++)
            for (int nl = 0; nl <= n; nl++)
              for (int g = 0; g < NG; g++) t.zk[kd][t.iz(g, ni, nj, nk, nl)] = me1d(t, true, g, ni, nj, nk, nl, KIND_D[kd], KIND_P[kd]);
    }
    {
      const int n = t.nsm;
      const size_t d = n + 1;
      t.rk[kd].assign((size_t)NG * d * d * d * d, 0.0);
#pragma omp parallel for collapse(2) schedule(dynamic)
      for (int ni = 0; ni <= n; ni++)
        for (int nj = 0; nj <= n; nj++)
          for (int nk = 0; nk <= n; nk++)
            for (int nl = 0; nl <= n; nl++)
              for (int g = 0; g < NG; g++) t.rk[kd][t.ir(g, ni, nj, nk, nl)] = me1d(t, false, g, ni, nj, nk, nl, KIND_D[kd], KIND_P[kd]);
      t.rkT[kd].assign(t.rk[kd].size(), 0.0);
      for (int g = 0; g < NG; g++)
        for (int ni = 0; ni <= n; ni++)
          for (int nj = 0; nj <= n; nj++)
            for (int nk = 0; nk <= n; nk++)
              for (int nl = 0; nl <= n; nl++) t.rkT[kd][t.ir(g, ni, nj, nl, nk)] = t.rk[kd][t.ir(g, ni, nj, nk, nl)];
      t.rkg[kd].assign(t.rk[kd].size(), 0.0); t.rkgT[kd].assign(t.rk[kd].size(), 0.0);
      for (int g = 0; g < NG; g++)
        for (int ni = 0; ni <= n; ni++)
          for (int nj = 0; nj <= n; nj++)
            for (int nk = 0; nk <= n; nk++)
              for (int nl = 0; nl <= n; nl++) {
                t.rkg[kd][(t.ir(0, ni, nj, nk, nl)) * NG + g] = t.rk[kd][t.ir(g, ni, nj, nk, nl)];
                t.rkgT[kd][(t.ir(0, ni, nj, nk, nl)) * NG + g] = t.rkT[kd][t.ir(g, ni, nj, nk, nl)];
              }
    }
  }
  return t;
}

// MatrixElement_radx (pnfam_spatial_mtxels.f90:387-450): derivative kind `kx` along x, plain Gaussian along y.
// Same terms in the same order; offsets and coefficient products hoisted, the sign alternates along the innermost index.
double radx(const Tables& t, int g, int kx, int ni, int li, int nj, int lj, int nk, int lk, int nl, int ll) {
  const size_t n1 = t.nsm + 1, n2 = n1 * n1, n3 = n2 * n1, slab = n3 * n1;
  const double* mx = t.rk[kx].data() + (size_t)g * slab;
  const double* my = t.rk[G00].data() + (size_t)g * slab;
  const int Ni = 2 * ni + std::abs(li), Nj = 2 * nj + std::abs(lj), Nk = 2 * nk + std::abs(lk), Nl = 2 * nl + std::abs(ll);
  // the x element vanishes unless (sum of the x quantum numbers + d + [p != 0]) is even; the y sum is even by construction
  if ((Ni + Nj + Nk + Nl + KIND_D[kx] + (KIND_P[kx] ? 1 : 0)) % 2 != 0) return 0.0;
  const double* Ci = &t.cp2c[((size_t)ni * (4 * t.nsh + 1) + (li + 2 * t.nsh)) * (2 * t.nsh + 1)];
  const double* Cj = &t.cp2c[((size_t)nj * (4 * t.nsh + 1) + (lj + 2 * t.nsh)) * (2 * t.nsh + 1)];
  const double* Ck = &t.cp2c[((size_t)nk * (4 * t.nsh + 1) + (lk + 2 * t.nsh)) * (2 * t.nsh + 1)];
  const double* Cl = &t.cp2c[((size_t)nl * (4 * t.nsh + 1) + (ll + 2 * t.nsh)) * (2 * t.nsh + 1)];
  double v = 0.0;
  for (int yi = 0; yi <= Ni; yi++) {
    const double ci = Ci[yi];
    const size_t xi_ = (size_t)(Ni - yi) * n3, yi_ = (size_t)yi * n3;
    for (int yj = 0; yj <= Nj; yj++) {
      const double cij = ci * Cj[yj];
      const size_t xj_ = xi_ + (size_t)(Nj - yj) * n2, yj_ = yi_ + (size_t)yj * n2;
      for (int yk = 0; yk <= Nk; yk++) {
        const double cijk = cij * Ck[yk];
        const double* px = mx + xj_ + (size_t)(Nk - yk) * n1 + Nl;
        const double* py = my + yj_ + (size_t)yk * n1;
        const int y0 = (yi + yj + yk) % 2;
        double sg = ((yi + yj + (yi + yj + yk + y0) / 2) % 2) ? -1.0 : 1.0;
        double acc = 0.0;
        for (int yl = y0; yl <= Nl; yl += 2) {
          acc += sg * Cl[yl] * px[-yl] * py[yl];
          sg = -sg;
        }
        v += cijk * acc;
      }
    }
  }
  return v;
}

// imaginary part of MatrixElement_rad_cmplx (pnfam_spatial_mtxels.f90:609-688) for kinds (kx along x, ky along y)
double rad_cmplx_imag(const Tables& t, int g, int kx, int ky, int ni, int li, int nj, int lj, int nk, int lk, int nl, int ll) {
  const size_t n1 = t.nsm + 1, n2 = n1 * n1, n3 = n2 * n1, slab = n3 * n1;
  const double* mx = t.rk[kx].data() + (size_t)g * slab;
  const double* my = t.rk[ky].data() + (size_t)g * slab;
  const int Ni = 2 * ni + std::abs(li), Nj = 2 * nj + std::abs(lj), Nk = 2 * nk + std::abs(lk), Nl = 2 * nl + std::abs(ll);
  const double* Ci = &t.cp2c[((size_t)ni * (4 * t.nsh + 1) + (li + 2 * t.nsh)) * (2 * t.nsh + 1)];
  const double* Cj = &t.cp2c[((size_t)nj * (4 * t.nsh + 1) + (lj + 2 * t.nsh)) * (2 * t.nsh + 1)];
  const double* Ck = &t.cp2c[((size_t)nk * (4 * t.nsh + 1) + (lk + 2 * t.nsh)) * (2 * t.nsh + 1)];
  const double* Cl = &t.cp2c[((size_t)nl * (4 * t.nsh + 1) + (ll + 2 * t.nsh)) * (2 * t.nsh + 1)];
  double v = 0.0;
  for (int yi = 0; yi <= Ni; yi++) {
    const size_t xi_ = (size_t)(Ni - yi) * n3, yi_ = (size_t)yi * n3;
    for (int yj = 0; yj <= Nj; yj++) {
      const double cij = Ci[yi] * Cj[yj];
      const size_t xj_ = xi_ + (size_t)(Nj - yj) * n2, yj_ = yi_ + (size_t)yj * n2;
      for (int yk = 0; yk <= Nk; yk++) {
        const double cijk = cij * Ck[yk];
        const double* px = mx + xj_ + (size_t)(Nk - yk) * n1 + Nl;
        const double* py = my + yj_ + (size_t)yk * n1;
        // i^(yi+yj-yk-yl) has an imaginary part only for an odd exponent: +1 for 1 (mod 4), -1 for 3 (mod 4)
        const int y0 = (yi + yj + yk + 1) % 2;
        int e = (((yi + yj - yk - y0) % 4) + 4) % 4;
        double sg = e == 1 ? 1.0 : -1.0;
        double acc = 0.0;
        for (int yl = y0; yl <= Nl; yl += 2) {
          acc += sg * Cl[yl] * px[-yl] * py[yl];
          sg = -sg;
        }
        v += cijk * acc;
      }
    }
  }
  return v;
}

// The four-fold Cartesian sum of MatrixElement_radx with the quantum numbers of two of the four states fixed:
//   radx(A, B, C, D) = sum_{yB, yD} cB(yB) cD(yD) (-1)^(yB + ceil((yB + yD) / 2)) Q[yB][yD][N_B - yB][N_D - yD],
//   Q[YB][YD][XB][XD] = sum_{yA, yC} cA(yA) cC(yC) (-1)^(yA + floor((yA + yC) / 2)) My[yA, YB, yC, YD] Mx[N_A - yA, XB, N_C - yC, XD]
// (the reference's sign (-1)^(yA + yB + (yA + yB + yC + yD) / 2) factorises like this on the even sums, the only ones
// with a non-zero y element).  One Q per (A, C, derivative kind, Gaussian) serves every pair (B, D), both orientations of
// their Lambda, and -- built from the tables with the last two indices exchanged -- the exchange elements
// radx(A, B, D, C).  O(N^4) per element becomes O(N^2).
// mode 0: the real sum above (y kind = plain Gaussian); modes 1, 2: the real / imaginary part of i^(yA - yC) instead of
// the sign, for MatrixElement_rad_cmplx (x kind `kind`, y kind `ky`):  Im sum i^(yA + yB - yC - yD) ... =
//   sum_{yB, yD} cB cD [ Im i^(yB - yD) Q_re + Re i^(yB - yD) Q_im ][yB][yD][N_B - yB][N_D - yD]
inline int ipow_re(int n) { n = ((n % 4) + 4) % 4; return n == 0 ? 1 : n == 2 ? -1 : 0; }
inline int ipow_im(int n) { n = ((n % 4) + 4) % 4; return n == 1 ? 1 : n == 3 ? -1 : 0; }
// Storage: Q is only needed where XB + YB = N_B and XD + YD = N_D are total Cartesian quanta of a basis state, so it is
// kept as Q[tri(N_B, yB)][tri(N_D, yD)][g], tri(N, y) = N (N + 1) / 2 + y, the six Gaussians innermost: 1.1 MB per
// (kind, orientation) at 16 shells; the sum of one element walks rows of it for all Gaussians at once, and the build
// shares its index arithmetic between them.
inline int tri(int N, int y) { return N * (N + 1) / 2 + y; }
void build_q(const Tables& t, int kind, bool exc, int nA, int lA, int nC, int lC, std::vector<double>& Q, int mode = 0, int ky = G00) {
  const size_t n1 = t.nsm + 1, n2 = n1 * n1, n3 = n2 * n1;
  const int nmax = t.nrlx;
  const size_t M = (size_t)(nmax + 1) * (nmax + 2) / 2;
  Q.assign(M * M * NG, 0.0);
  const double* mx = (exc ? t.rkgT[kind] : t.rkg[kind]).data();
  const double* my = (exc ? t.rkgT[ky] : t.rkg[ky]).data();
  const int NA = 2 * nA + std::abs(lA), NC = 2 * nC + std::abs(lC);
  const int xpar = KIND_D[kind] + (KIND_P[kind] ? 1 : 0), ypar = KIND_D[ky] + (KIND_P[ky] ? 1 : 0);
  const double* CA = &t.cp2c[((size_t)nA * (4 * t.nsh + 1) + (lA + 2 * t.nsh)) * (2 * t.nsh + 1)];
  const double* CC = &t.cp2c[((size_t)nC * (4 * t.nsh + 1) + (lC + 2 * t.nsh)) * (2 * t.nsh + 1)];
  double* q = Q.data();
  for (int yA = 0; yA <= NA; yA++)
    for (int yC = 0; yC <= NC; yC++) {
      const double sgn = mode == 0 ? ((((yA + (yA + yC) / 2) % 2) ? -1.0 : 1.0)) : mode == 1 ? (double)ipow_re(yA - yC) : (double)ipow_im(yA - yC);
      const double w = CA[yA] * CC[yC] * sgn;
      if (w == 0.0) continue;
      const double* mrow = mx + ((size_t)(NA - yA) * n3 + (size_t)(NC - yC) * n1) * NG;
      const int xd_par = (NA - yA + NC - yC + xpar) % 2;
      for (int YB = 0; YB <= nmax; YB++)
        for (int YD = (yA + YB + yC + ypar) % 2; YD <= nmax; YD += 2) {
          const double* myp = my + ((size_t)yA * n3 + (size_t)YB * n2 + (size_t)yC * n1 + YD) * NG;
          double l[NG];
          for (int g = 0; g < NG; g++) l[g] = w * myp[g];
          // target column tri(XD + YD, YD) = tri0[XD]: hoisted; the x element vanishes for an odd (sum + d + [p != 0])
          int tri0[64];
          for (int XD = 0; XD <= nmax - YD; XD++) tri0[XD] = tri(XD + YD, YD) * NG;
          for (int XB = 0; XB <= nmax - YB; XB++) {
            double* qq = q + (size_t)tri(XB + YB, YB) * M * NG;
            const double* mm = mrow + (size_t)XB * n2 * NG;
            for (int XD = (xd_par + XB) % 2; XD <= nmax - YD; XD += 2) {
              double* dst = qq + tri0[XD];
              const double* src = mm + (size_t)XD * NG;
              for (int g = 0; g < NG; g++) dst[g] += l[g] * src[g];
            }
          }
        }
    }
}

// all six Gaussians of one element
inline void radx_q(const Tables& t, const double* q, int nB, int lB, int nD, int lD, double* v) {
  const int NB = 2 * nB + std::abs(lB), ND = 2 * nD + std::abs(lD);
  const size_t M = (size_t)(t.nrlx + 1) * (t.nrlx + 2) / 2;
  const double* CB = &t.cp2c[((size_t)nB * (4 * t.nsh + 1) + (lB + 2 * t.nsh)) * (2 * t.nsh + 1)];
  const double* CD = &t.cp2c[((size_t)nD * (4 * t.nsh + 1) + (lD + 2 * t.nsh)) * (2 * t.nsh + 1)];
  for (int g = 0; g < NG; g++) v[g] = 0.0;
  for (int yB = 0; yB <= NB; yB++) {
    const double* row = q + ((size_t)tri(NB, yB) * M + tri(ND, 0)) * NG;
    double acc[NG] = {0, 0, 0, 0, 0, 0};
    for (int yD = 0; yD <= ND; yD++) {
      const double c = (((yB + (yB + yD + 1) / 2) % 2) ? -1.0 : 1.0) * CD[yD];
      for (int g = 0; g < NG; g++) acc[g] += c * row[(size_t)yD * NG + g];
    }
    for (int g = 0; g < NG; g++) v[g] += CB[yB] * acc[g];
  }
}

inline void rad_q_imag(const Tables& t, const double* qre, const double* qim, int nB, int lB, int nD, int lD, double* v) {
  const int NB = 2 * nB + std::abs(lB), ND = 2 * nD + std::abs(lD);
  const size_t M = (size_t)(t.nrlx + 1) * (t.nrlx + 2) / 2;
  const double* CB = &t.cp2c[((size_t)nB * (4 * t.nsh + 1) + (lB + 2 * t.nsh)) * (2 * t.nsh + 1)];
  const double* CD = &t.cp2c[((size_t)nD * (4 * t.nsh + 1) + (lD + 2 * t.nsh)) * (2 * t.nsh + 1)];
  for (int g = 0; g < NG; g++) v[g] = 0.0;
  for (int yB = 0; yB <= NB; yB++) {
    const size_t at = ((size_t)tri(NB, yB) * M + tri(ND, 0)) * NG;
    double acc[NG] = {0, 0, 0, 0, 0, 0};
    for (int yD = 0; yD <= ND; yD++) {
      const double ci = CD[yD] * ipow_im(yB - yD), cr = CD[yD] * ipow_re(yB - yD);
      for (int g = 0; g < NG; g++) acc[g] += ci * qre[at + (size_t)yD * NG + g] + cr * qim[at + (size_t)yD * NG + g];
    }
    for (int g = 0; g < NG; g++) v[g] += CB[yB] * acc[g];
  }
}

inline int kind_of(int d, int pp) {
  for (int k = 0; k < NKIND; k++) if (KIND_D[k] == d && KIND_P[k] == pp) return k;
  throw std::runtime_error("two-body currents: invalid derivative kind");
}

// MatrixElement_ddGr (pnfam_spatial_mtxels.f90:292-371); msum = -li - lj + lk + ll of the four states,
// X(kind) = MatrixElement_radx for that derivative kind, XC(kind) = Im MatrixElement_rad_cmplx(dx = 1, py = kind)
template <class RX, class RC>
double ddgr(int msum, RX&& X, RC&& XC, int di, int dj, int dip) {
  if (msum + di + dj != 0) return 0.0;
  const int p = dip != 0 ? 1 : 0;
  const double cr2 = std::sqrt(2.0);
  if (di == 0 && dj == 0) return X(kind_of(0, 0));
  if (std::abs(di + dj) == 2) return X(kind_of(2 - p, dip)) * 2.0;
  if (std::abs(di + dj) == 1) return dj == 0 ? X(kind_of(1 - p, dip)) * (-(di + dj) * cr2) : X(kind_of(1, 0)) * (-(di + dj) * cr2);
  double v = X(kind_of(2 - p, dip)) * (-1.0);
  if (dip != 0) v += XC(kind_of(0, dip));
  return v;
}

// calc_Jr_opt (pnfam_type_extfield_2bc.f90:294-398): radial components, no prefactors
template <class RX, class RC>
void calc_jr(int K, bool use_p, int msum, RX&& X, RC&& XC, double* J) {
  for (int c = 0; c < NCOMP; c++) J[c] = 0.0;
  auto dd = [&](int di, int dj, int dip) { return ddgr(msum, X, XC, di, dj, dip); };
  if (use_p) {
    J[R0P] = dd(K, -1, 1); J[R0Z] = dd(K, 0, 1); J[R0M] = dd(K, +1, 1);
    J[RP0] = dd(K, -1, 2); J[RZ0] = dd(K, 0, 2); J[RM0] = dd(K, +1, 2);
  }
  const double dpm = dd(+1, -1, 0), d00 = dd(0, 0, 0);
  if (K == 1) {
    const double dp0 = dd(+1, 0, 0), dpp = dd(+1, +1, 0);
    J[RKP] = dpm; J[RKZ] = dp0; J[RKM] = dpp; J[RPZ] = d00; J[RPM] = dp0; J[RZM] = dpp; J[RPZ2] = dpm;
  } else if (K == 0) {
    const double d0m = dd(0, -1, 0), dp0 = dd(+1, 0, 0);
    J[RKP] = d0m; J[RKZ] = d00; J[RKM] = dp0; J[RPZ] = d0m; J[RPM] = dpm; J[RZM] = dp0;
  } else {
    const double d0m = dd(0, -1, 0), dmm = dd(-1, -1, 0);
    J[RKP] = dmm; J[RKZ] = d0m; J[RKM] = dpm; J[RPZ] = dmm; J[RPM] = d0m; J[RZM] = d00; J[RZM2] = dpm;
  }
}

// calc_Jz_opt (pnfam_type_extfield_2bc.f90:182-288) with unit low-energy constants: the c3 components (rk*) carry
// caux, the c4 components (rp*, rz*) carry 4 caux, the momentum components caux -- what is left after strip_gamdel_lecs
struct JzCoef { double c3_c4r2, c3_8, y_c2r2, y_2, y_4, cr2, two; };
void calc_jz(const Tables& t, const JzCoef& q, int K, bool use_p, int g, int i, int j, int k, int l, double* J) {
  for (int c = 0; c < NCOMP; c++) J[c] = 0.0;
  const size_t x = t.iz(g, i, j, k, l);
  const double g00 = t.zk[G00][x], g10 = t.zk[G10][x], g20 = t.zk[G20][x];
  if (use_p) {
    double p1dp, p1d0, p1dm, p2dp, p2d0, p2dm;
    if (K == 0) { p1dp = p1dm = t.zk[G01][x]; p1d0 = t.zk[G11][x]; p2dp = p2dm = t.zk[G02][x]; p2d0 = t.zk[G12][x]; }
    else { p1dp = p1dm = p2dp = p2dm = g00; p1d0 = p2d0 = g10; }
    J[R0P] = +q.cr2 * p1dm; J[R0Z] = +q.two * p1d0; J[R0M] = -q.cr2 * p1dp;
    J[RP0] = +q.cr2 * p2dm; J[RZ0] = +q.two * p2d0; J[RM0] = -q.cr2 * p2dp;
  }
  if (K == 1) {
    J[RKP] = -q.c3_c4r2 * g00; J[RKZ] = -q.c3_8 * g10; J[RKM] = +q.c3_c4r2 * g00;
    J[RPZ] = +q.y_c2r2 * g20; J[RPM] = -q.y_2 * g10; J[RZM] = -q.y_c2r2 * g00; J[RPZ2] = -q.y_c2r2 * g00;
  } else if (K == 0) {
    J[RKP] = -q.c3_c4r2 * g10; J[RKZ] = -q.c3_8 * g20; J[RKM] = +q.c3_c4r2 * g10;
    J[RPZ] = +q.y_c2r2 * g10; J[RPM] = -q.y_4 * g00; J[RZM] = -q.y_c2r2 * g10;
  } else {
    J[RKP] = -q.c3_c4r2 * g00; J[RKZ] = -q.c3_8 * g10; J[RKM] = +q.c3_c4r2 * g00;
    J[RPZ] = +q.y_c2r2 * g00; J[RPM] = -q.y_2 * g10; J[RZM] = -q.y_c2r2 * g20; J[RZM2] = +q.y_c2r2 * g00;
  }
}

// ---- spin contraction (pnfam_type_extfield_2bc.f90:869-1463), literal; J holds the 12 merged components ------------
struct J12 { double rkp, rkz, rkm, rpz, rpm, rzm, r0p, r0z, r0m, rp0, rz0, rm0; };
constexpr double h = 0.5;
inline int db_case(int sd, int sb) {      // 0: rho_du (sd=-1, sb=+1), 1: rho_ud, 2: rho_uu, 3: rho_dd
  if (sd == -1 && sb == +1) return 0;
  if (sd == +1 && sb == -1) return 1;
  return sd + sb > 0 ? 2 : 3;
}
double jrsa_dir(const J12& J, int sa, int sc, int sd, int sb) {
  if (sd + sb == 0) return 0.0;
  if (sa == sc) return (sa * h) * J.rkz;
  return sa == +1 ? J.rkp : J.rkm;
}
double jrsx_dir(const J12& J, int sa, int sc, int sd, int sb) {
  const int c = db_case(sd, sb);
  if (sa == sc) return c == 0 ? (sa * h) * (-J.rpz) : c == 1 ? (sa * h) * J.rzm : 0.0;
  if (sa == +1) return c == 0 ? 0.0 : c == 1 ? J.rpm : c == 2 ? +h * J.rpz : -h * J.rpz;
  return c == 0 ? -J.rpm : c == 1 ? 0.0 : c == 2 ? -h * J.rzm : +h * J.rzm;
}
double jrsxp_dir(const J12& J, int sa, int sc, int sd, int sb) {
  const int c = db_case(sd, sb);
  if (sa == sc) return c == 0 ? J.r0p : c == 1 ? J.r0m : c == 2 ? (sa * h) * J.rz0 + h * J.r0z : (sa * h) * J.rz0 - h * J.r0z;
  if (sa == +1) return c >= 2 ? J.rp0 : 0.0;
  return c >= 2 ? J.rm0 : 0.0;
}
inline int ac_case(int sa, int sc) {      // 0: uu, 1: dd, 2: ud, 3: du
  if (sa == +1 && sc == +1) return 0;
  if (sa == -1 && sc == -1) return 1;
  return sa == +1 ? 2 : 3;
}
double jrsa_exc(const J12& J, int sa, int sc, int sd, int sb) {
  const int c = db_case(sd, sb);
  switch (ac_case(sa, sc)) {
    case 0: return c == 0 ? J.rkp : c == 2 ? h * J.rkz : 0.0;
    case 1: return c == 1 ? J.rkm : c == 3 ? -h * J.rkz : 0.0;
    case 2: return c == 1 ? h * J.rkz : c == 3 ? J.rkp : 0.0;
    default: return c == 0 ? -h * J.rkz : c == 2 ? J.rkm : 0.0;
  }
}
double jrsb_exc(const J12& J, int sa, int sc, int sd, int sb) {
  const int c = db_case(sd, sb);
  switch (ac_case(sa, sc)) {
    case 0: return c == 1 ? J.rkm : c == 2 ? h * J.rkz : 0.0;
    case 1: return c == 0 ? J.rkp : c == 3 ? -h * J.rkz : 0.0;
    case 2: return c == 1 ? -h * J.rkz : c == 2 ? J.rkp : 0.0;
    default: return c == 0 ? +h * J.rkz : c == 3 ? J.rkm : 0.0;
  }
}
double jrsx_exc(const J12& J, int sa, int sc, int sd, int sb) {
  const int c = db_case(sd, sb);
  switch (ac_case(sa, sc)) {
    case 0: return c == 0 ? h * J.rpz : c == 1 ? h * J.rzm : c == 2 ? 0.0 : J.rpm;
    case 1: return c == 0 ? h * J.rpz : c == 1 ? h * J.rzm : c == 2 ? -J.rpm : 0.0;
    case 2: return c >= 2 ? -h * J.rpz : 0.0;
    default: return c >= 2 ? -h * J.rzm : 0.0;
  }
}
double jrsxp_exc(const J12& J, int sa, int sc, int sd, int sb) {
  const int c = db_case(sd, sb);
  switch (ac_case(sa, sc)) {
    case 0: return c == 0 ? J.rp0 : c == 1 ? J.r0m : c == 2 ? h * J.rz0 + h * J.r0z : 0.0;
    case 1: return c == 0 ? J.r0p : c == 1 ? J.rm0 : c == 2 ? 0.0 : -h * J.rz0 - h * J.r0z;
    case 2: return c == 0 ? 0.0 : c == 1 ? h * J.rz0 - h * J.r0z : c == 2 ? J.r0p : J.rp0;
    default: return c == 0 ? -h * J.rz0 + h * J.r0z : c == 1 ? 0.0 : c == 2 ? J.rm0 : J.r0m;
  }
}

// calc_gam_sep (pnfam_type_extfield_2bc.f90:750-825) as a linear map: out[6] = sum_{src, comp} coef * product[src][comp],
// src = 0: dir x rho_n, 1: dir x rho_p, 2: exc x rho_n, 3: exc x rho_p; probed with unit vectors
struct SpinCoef { double c[6][4][NCOMP]; };
SpinCoef spin_coef(int sa, int sc, int sd, int sb, bool bminus) {
  SpinCoef o;
  std::memset(&o, 0, sizeof o);
  for (int comp = 0; comp < NCOMP; comp++) {
    J12 J{0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    switch (comp) {
      case RKP: J.rkp = 1; break; case RKZ: J.rkz = 1; break; case RKM: J.rkm = 1; break;
      case RPZ: case RPZ2: J.rpz = 1; break; case RPM: J.rpm = 1; break; case RZM: case RZM2: J.rzm = 1; break;
      case R0P: J.r0p = 1; break; case R0Z: J.r0z = 1; break; case R0M: J.r0m = 1; break;
      case RP0: J.rp0 = 1; break; case RZ0: J.rz0 = 1; break; case RM0: J.rm0 = 1; break;
    }
    const double a_d = jrsa_dir(J, sa, sc, sd, sb), x_d = jrsx_dir(J, sa, sc, sd, sb), xp_d = jrsxp_dir(J, sa, sc, sd, sb);
    const double a_e = jrsa_exc(J, sa, sc, sd, sb), b_e = jrsb_exc(J, sa, sc, sd, sb), x_e = jrsx_exc(J, sa, sc, sd, sb),
                 xp_e = jrsxp_exc(J, sa, sc, sd, sb);
    const double sg = bminus ? 1.0 : -1.0;
    o.c[0][0][comp] = a_d; o.c[0][1][comp] = a_d;                                       // c3d
    o.c[1][2][comp] = bminus ? -a_e : -b_e; o.c[1][3][comp] = bminus ? -b_e : -a_e;     // c3e
    o.c[2][0][comp] = sg * h * x_d; o.c[2][1][comp] = -sg * h * x_d;                    // c4d
    o.c[3][2][comp] = -h * x_e; o.c[3][3][comp] = -h * x_e;                             // c4e
    o.c[4][0][comp] = sg * h * xp_d; o.c[4][1][comp] = -sg * h * xp_d;                  // cpd
    o.c[5][2][comp] = -h * xp_e; o.c[5][3][comp] = -h * xp_e;                           // cpe
  }
  return o;
}

struct RadClass {            // one (Omega, n_r, Lambda, s) class of the undoubled basis with its n_z members
  int np, r, l, s, nl;       // np = Omega + 1/2, nl = 0 (Lambda = Omega - 1/2, s = +1) or 1
  std::vector<int> state;    // undoubled state indices
};

}  // namespace

TbcField generate_two_body_current_field(const TbcProblem& s) {
  const int K = s.K;
  if (K < -1 || K > 1) throw std::runtime_error("ERROR, invalid K for extfield 2bc");
  const bool use_p = s.use_p;
  const int nt = s.nt, nbx = s.nb;
  const size_t nxy = s.nxy;
  if ((int)s.id.size() != nbx || (int)s.nz.size() != nt || (int)s.ir2c.size() != 2 * nbx || (int)s.ir2m.size() != 2 * nbx)
    throw std::runtime_error("two-body-current field generator: inconsistent array sizes");
  const bool timing = getenv("PNFAM_B200_SETUP_TIMING") != nullptr;
  // self-check switch: the reference's own four-fold Cartesian sum for every radial element instead of the intermediate
  const bool literal = getenv("PNFAM_B200_TBC_LITERAL_RADIAL") != nullptr;
  double t_phase = omp_get_wtime(), t_jr = 0.0, t_con = 0.0, t_q = 0.0;
  auto lap = [&](const char* what) {
    if (timing) { const double now = omp_get_wtime(); std::fprintf(stderr, "[setup]   2BC %-18s %.3f s\n", what, now - t_phase); t_phase = now; }
  };
  const Tables t = build_tables(s, use_p);
  lap("1D tables");
  const int ncomp = use_p ? (int)NCOMP : (int)R0P;

  // prefactors with unit LECs
  const double caux = (HBARC * HBARC * HBARC) / (2.0 * MN * FPI * FPI), cr2 = std::sqrt(2.0);
  JzCoef q;
  q.c3_c4r2 = caux * 4.0 * cr2; q.c3_8 = caux * 8.0;
  const double cy = caux * 4.0;
  q.y_c2r2 = cy * 2.0 * cr2; q.y_2 = cy * 2.0; q.y_4 = cy * 4.0; q.cr2 = caux * cr2; q.two = caux * 2.0;

  std::vector<size_t> boff(nbx + 1, 0);
  for (int ib = 0; ib < nbx; ib++) boff[ib + 1] = boff[ib] + (size_t)s.id[ib] * s.id[ib];
  if (s.rho[0].size() != boff[nbx] || s.rho[1].size() != boff[nbx])
    throw std::runtime_error("two-body-current field generator: density matrices of the wrong size");
  const std::vector<double>* rho = s.rho;
  std::vector<int> ia(nbx, 0);
  for (int ib = 1; ib < nbx; ib++) ia[ib] = ia[ib - 1] + s.id[ib - 1];
  std::vector<int> blk(nt), pos(nt);
  for (int ib = 0; ib < nbx; ib++)
    for (int i = 0; i < s.id[ib]; i++) { blk[ia[ib] + i] = ib; pos[ia[ib] + i] = i; }

  // radial classes in the reference's order (Omega, n_r, Lambda, s), members by n_z
  std::vector<RadClass> cls;
  std::vector<int> cls_of(nt, -1);
  {
    std::map<std::array<int, 4>, int> idx;
    for (int i = 0; i < nt; i++) {
      const int np = s.nl[i] + (s.ns[i] + 1) / 2;
      std::array<int, 4> key{np, s.nr[i], s.nl[i], -s.ns[i]};
      auto it = idx.find(key);
      if (it == idx.end()) it = idx.emplace(key, 0).first;
    }
    int n = 0;
    for (auto& kv : idx) {
      kv.second = n++;
      RadClass c;
      c.np = kv.first[0]; c.r = kv.first[1]; c.l = kv.first[2]; c.s = -kv.first[3];
      c.nl = (c.np - c.l + 1) % 2;
      cls.push_back(c);
    }
    for (int i = 0; i < nt; i++) {
      const int np = s.nl[i] + (s.ns[i] + 1) / 2;
      const int ci = idx[{np, s.nr[i], s.nl[i], -s.ns[i]}];
      cls[ci].state.push_back(i);
      cls_of[i] = ci;
    }
  }
  const int ncls = (int)cls.size();
  // pairs (D, B) of classes with equal Omega
  std::vector<int> pair_id((size_t)ncls * ncls, -1);
  std::vector<std::array<int, 2>> pairs;
  for (int D = 0; D < ncls; D++)
    for (int B = 0; B < ncls; B++)
      if (cls[D].np == cls[B].np) { pair_id[(size_t)D * ncls + B] = (int)pairs.size(); pairs.push_back({D, B}); }
  const int npairs = (int)pairs.size();
  const size_t wrec = (size_t)NG * ncomp;             // one (pair) record: [comp][g]
  const size_t wsz = (size_t)4 * npairs * wrec;       // [src = x*2+q ... see below][pair][comp][g]

  // doubled states (time-reversed partners after the originals), grouped by (n_r, Lambda)
  struct Dst { int z, r, l, sp, sign, blk, pos; };
  std::vector<Dst> dst(2 * nt);
  for (int i = 0; i < nt; i++) {
    dst[i] = {s.nz[i], s.nr[i], s.nl[i], s.ns[i], 1, blk[i], pos[i]};
    dst[i + nt] = {s.nz[i], s.nr[i], -s.nl[i], -s.ns[i], s.ns[i] < 0 ? -1 : 1, blk[i] + nbx, pos[i]};
  }
  std::map<std::array<int, 2>, std::vector<int>> group;
  for (int i = 0; i < 2 * nt; i++) group[{dst[i].r, dst[i].l}].push_back(i);
  std::vector<std::array<int, 2>> gkeys;
  for (auto& kv : group) gkeys.push_back(kv.first);
  struct Combo { int A, C; };
  std::vector<Combo> combos;
  for (int A = 0; A < (int)gkeys.size(); A++)
    for (int C = 0; C < (int)gkeys.size(); C++) {
      bool any = false;
      for (int a : group[gkeys[A]]) {
        for (int c : group[gkeys[C]])
          if (s.ir2c[dst[a].blk] - 1 == dst[c].blk) { any = true; break; }
        if (any) break;
      }
      if (any) combos.push_back({A, C});
    }

  // spin coefficient tables for (sa, sc, sd, sb) in {-1, +1}^4
  std::vector<SpinCoef> sc_tab(16);
  auto sidx = [](int sa, int sc, int sd, int sb) { return ((sa + 1) / 2) * 8 + ((sc + 1) / 2) * 4 + ((sd + 1) / 2) * 2 + (sb + 1) / 2; };
  for (int sa = -1; sa <= 1; sa += 2)
    for (int sc = -1; sc <= 1; sc += 2)
      for (int sd = -1; sd <= 1; sd += 2)
        for (int sb = -1; sb <= 1; sb += 2) sc_tab[sidx(sa, sc, sd, sb)] = spin_coef(sa, sc, sd, sb, s.beta_minus);

  struct SpinTerm { int out, src, comp; double coef; };
  std::vector<std::vector<SpinTerm>> sc_list(16);
  for (int i = 0; i < 16; i++)
    for (int o = 0; o < 6; o++)
      for (int src = 0; src < 4; src++)
        for (int c = 0; c < ncomp; c++)
          if (sc_tab[i].c[o][src][c] != 0.0) sc_list[i].push_back({o, src, c, sc_tab[i].c[o][src][c]});

  // unsorted result: index ir2m(block of a) - 1 + pos_a + pos_c * d_a in the reference's original in-block order
  std::vector<double> raw[6];
  for (auto& v : raw) v.assign(nxy, 0.0);

  // z contraction for all (z_a, z_c) of a chunk of z_a values (memory bound: a quarter of the RAM by default)
  const int nz1 = t.nzx + 1;
  const size_t per_za = (size_t)nz1 * wsz * sizeof(double);
  // bound on the z-contracted densities kept at once: a quarter of the physical memory, at least 4 GB (PNFAM_B200_TBC_MEM_GB);
  // with several chunks the radial elements are recomputed per chunk
  double mem_gb = 4.0;
  {
    const long pages = sysconf(_SC_PHYS_PAGES), psize = sysconf(_SC_PAGE_SIZE);
    if (pages > 0 && psize > 0) mem_gb = std::max(4.0, 0.25 * (double)pages * (double)psize / 1073741824.0);
  }
  if (const char* e = getenv("PNFAM_B200_TBC_MEM_GB")) mem_gb = std::max(1e-4, atof(e));
  const int za_chunk = (int)std::max<size_t>(1, std::min<size_t>(nz1, (size_t)(mem_gb * 1073741824.0) / std::max<size_t>(1, per_za)));
  std::vector<double> W;
  for (int za0 = 0; za0 < nz1; za0 += za_chunk) {
    const int za1 = std::min(nz1, za0 + za_chunk);
    W.assign((size_t)(za1 - za0) * nz1 * wsz, 0.0);
    // W[(za, zc)][src][pair][g][comp], src = 0: dir rho_n, 1: dir rho_p, 2: exc rho_n, 3: exc rho_p
#pragma omp parallel for collapse(2) schedule(dynamic)
    for (int za = za0; za < za1; za++)
      for (int zc = 0; zc < nz1; zc++) {
        double* w = W.data() + ((size_t)(za - za0) * nz1 + zc) * wsz;
        double Jd[NCOMP], Je[NCOMP];
        for (int ib = 0; ib < nbx; ib++) {
          const int nd = s.id[ib], i0 = ia[ib];
          for (int jb = 0; jb < nd; jb++)
            for (int jd = 0; jd < nd; jd++) {
              const int sb_ = i0 + jb, sd_ = i0 + jd;
              const int pid = pair_id[(size_t)cls_of[sd_] * ncls + cls_of[sb_]];
              if (pid < 0) continue;
              const double rn = rho[0][boff[ib] + jd + (size_t)jb * nd], rp = rho[1][boff[ib] + jd + (size_t)jb * nd];
              const int zb = s.nz[sb_], zd = s.nz[sd_];
              for (int g = 0; g < NG; g++) {
                calc_jz(t, q, K, use_p, g, za, zb, zc, zd, Jd);
                calc_jz(t, q, K, use_p, g, za, zb, zd, zc, Je);
                double* w0 = w + ((size_t)0 * npairs + pid) * wrec + g;
                double* w1 = w + ((size_t)1 * npairs + pid) * wrec + g;
                double* w2 = w + ((size_t)2 * npairs + pid) * wrec + g;
                double* w3 = w + ((size_t)3 * npairs + pid) * wrec + g;
                for (int c = 0; c < ncomp; c++) { w0[c * NG] += Jd[c] * rn; w1[c * NG] += Jd[c] * rp; w2[c * NG] += Je[c] * rn; w3[c * NG] += Je[c] * rp; }
              }
            }
        }
      }

    lap("z contraction");
    // radial elements per (A, C) and the contraction
#pragma omp parallel for schedule(dynamic) reduction(+ : t_jr, t_con, t_q)
    for (int ic = 0; ic < (int)combos.size(); ic++) {
      const double tt0 = omp_get_wtime();
      const std::array<int, 2> kA = gkeys[combos[ic].A], kC = gkeys[combos[ic].C];
      const int ra = kA[0], la = kA[1], rc = kC[0], lc = kC[1];
      // JR[pair][v][g][comp], v = 0: dir, 1: exc, 2: dir with time-reversed (d, b), 3: exc with time-reversed (d, b)
      std::vector<double> JR((size_t)npairs * 4 * wrec, 0.0);
      std::vector<char> use_n(npairs, 0), use_t(npairs, 0);
      std::vector<double> Q[NKIND][2];
      auto q_of = [&](int kind, int exc) -> const double* {
        if (Q[kind][exc].empty()) { const double q0 = omp_get_wtime(); build_q(t, kind, exc != 0, ra, la, rc, lc, Q[kind][exc]); t_q += omp_get_wtime() - q0; }
        return Q[kind][exc].data();
      };
      std::vector<double> QC[NKIND][2][2];      // rad_cmplx intermediates: [y kind][exc][re | im]
      auto qc_of = [&](int ky, int exc, int part) -> const double* {
        if (QC[ky][exc][part].empty()) { const double q0 = omp_get_wtime(); build_q(t, G10, exc != 0, ra, la, rc, lc, QC[ky][exc][part], 1 + part, ky); t_q += omp_get_wtime() - q0; }
        return QC[ky][exc][part].data();
      };
      for (int p = 0; p < npairs; p++) {
        const RadClass& D = cls[pairs[p][0]];
        const RadClass& B = cls[pairs[p][1]];
        const int rb = B.r, lb = B.l, rd = D.r, ld = D.l;
        const int dn = la + lb - lc - ld, dt = la - lb - lc + ld;
        use_n[p] = dn >= K - 1 && dn <= K + 1;
        use_t[p] = dt >= K - 1 && dt <= K + 1;
        for (int v = 0; v < 4; v++) {
          if (v < 2 ? !use_n[p] : !use_t[p]) continue;
          const int exc = v & 1, sl = v < 2 ? 1 : -1;              // orientation of Lambda_b, Lambda_d
          const int msum = -la - sl * lb + lc + sl * ld;
          // the elements of all six Gaussians per derivative kind, made on first use
          double xv[NKIND][NG], xcv[NKIND][NG];
          bool have_x[NKIND] = {false, false, false, false, false, false, false}, have_xc[NKIND] = {false, false, false, false, false, false, false};
          for (int g = 0; g < NG; g++) {
            double* j = JR.data() + ((size_t)p * 4) * wrec + g;
            double tmp[NCOMP];
            auto X = [&](int kind) {
              if (literal) return exc ? radx(t, g, kind, ra, la, rb, sl * lb, rd, sl * ld, rc, lc) : radx(t, g, kind, ra, la, rb, sl * lb, rc, lc, rd, sl * ld);
              if (!have_x[kind]) { radx_q(t, q_of(kind, exc), rb, sl * lb, rd, sl * ld, xv[kind]); have_x[kind] = true; }
              return xv[kind][g];
            };
            auto XC = [&](int kind) {
              if (literal) return exc ? rad_cmplx_imag(t, g, G10, kind, ra, la, rb, sl * lb, rd, sl * ld, rc, lc)
                                      : rad_cmplx_imag(t, g, G10, kind, ra, la, rb, sl * lb, rc, lc, rd, sl * ld);
              if (!have_xc[kind]) { rad_q_imag(t, qc_of(kind, exc, 0), qc_of(kind, exc, 1), rb, sl * lb, rd, sl * ld, xcv[kind]); have_xc[kind] = true; }
              return xcv[kind][g];
            };
            calc_jr(K, use_p, msum, X, XC, tmp);
            for (int c = 0; c < ncomp; c++) j[(size_t)v * wrec + (size_t)c * NG] = tmp[c];
          }
        }
      }
      const double tt1 = omp_get_wtime();
      t_jr += tt1 - tt0;
      for (int a : group.at(kA)) {
        const Dst& sa_ = dst[a];
        if (sa_.z < za0 || sa_.z >= za1) continue;
        for (int c : group.at(kC)) {
          const Dst& sc_ = dst[c];
          if (s.ir2c[sa_.blk] - 1 != sc_.blk) continue;
          const double* w = W.data() + ((size_t)(sa_.z - za0) * nz1 + sc_.z) * wsz;
          double gam[6] = {0, 0, 0, 0, 0, 0};
          for (int p = 0; p < npairs; p++) {
            if (!use_n[p] && !use_t[p]) continue;
            const RadClass& D = cls[pairs[p][0]];
            const RadClass& B = cls[pairs[p][1]];
            const int sd = D.s, sb = B.s;
            const double sign_db = (sd + sb == 0) ? -1.0 : 1.0;
            for (int tr = 0; tr < 2; tr++) {
              if (tr == 0 ? !use_n[p] : !use_t[p]) continue;
              const std::vector<SpinTerm>& terms = tr == 0 ? sc_list[sidx(sa_.sp, sc_.sp, sd, sb)] : sc_list[sidx(sa_.sp, sc_.sp, -sd, -sb)];
              const double fac = tr == 0 ? 1.0 : sign_db;
              for (const SpinTerm& tm : terms) {
                const double* ws = w + ((size_t)tm.src * npairs + p) * wrec + (size_t)tm.comp * NG;
                const double* js = JR.data() + ((size_t)p * 4 + (tm.src / 2) + 2 * tr) * wrec + (size_t)tm.comp * NG;
                double acc = 0.0;
                for (int g = 0; g < NG; g++) acc += ws[g] * js[g];
                gam[tm.out] += fac * tm.coef * acc;
              }
            }
          }
          const int da = s.id[sa_.blk % nbx];
          const size_t at = (size_t)s.ir2m[sa_.blk] - 1 + sa_.pos + (size_t)sc_.pos * da;
          const double sg = (double)(sa_.sign * sc_.sign);
          for (int o = 0; o < 6; o++) raw[o][at] = gam[o] * sg;
        }
      }
      t_con += omp_get_wtime() - tt1;
    }
    lap("radial + spin");
  }
  if (timing) std::fprintf(stderr, "[setup]   2BC thread-seconds: radial elements %.2f (intermediates %.2f), contraction %.2f (%d classes, %d pairs, %d (A,C))\n",
                           t_jr, t_q, t_con, ncls, npairs, (int)combos.size());

  // spin sort of rows and columns (reorder_blockmatrix_basis 'a' with new_order, pnfam_extfield_2bc.f90:903-913)
  TbcField out;
  std::vector<std::vector<int>> order(2 * nbx);
  for (int ib = 0; ib < 2 * nbx; ib++) {
    const int h0 = ib < nbx ? ib : ib - nbx;
    const int flip = ib < nbx ? 1 : -1;
    for (int i = 0; i < s.id[h0]; i++) if (flip * s.ns[ia[h0] + i] > 0) order[ib].push_back(i);
    for (int i = 0; i < s.id[h0]; i++) if (flip * s.ns[ia[h0] + i] < 0) order[ib].push_back(i);
  }
  for (int o = 0; o < 6; o++) {
    out.c[o].assign(nxy, 0.0);
    for (int ibr = 0; ibr < 2 * nbx; ibr++) {
      const int ibc = s.ir2c[ibr] - 1;
      if (ibc < 0) continue;
      const int dr = s.id[ibr % nbx], dc = s.id[ibc % nbx];
      const size_t im = (size_t)s.ir2m[ibr] - 1;
      for (int c = 0; c < dc; c++)
        for (int r = 0; r < dr; r++) out.c[o][im + r + (size_t)c * dr] = s.spin_sorted ? raw[o][im + order[ibr][r] + (size_t)order[ibc][c] * dr] : raw[o][im + r + (size_t)c * dr];
    }
  }
  return out;
}

// Front door of the set-up: density matrices rho_db = rk / 2 (hfbtho_solver.f90:1822-1856) from the HFB solution:
// sum_k V_dk V_bk over the pairing window; finite temperature: V (1 - f_k) V + U f_k U with the Fermi-Dirac factor of the
// quasiparticle energy (:1749-1761); blocked level (equal filling): - (V V - U U) / 2 of the blocked quasiparticle.
TbcProblem tbc_problem_from(const HfbSolution& s, const ExtField& f, bool use_p) {
  TbcProblem pr;
  pr.nt = s.nt; pr.nb = s.nb; pr.id = s.id; pr.nz = s.nz; pr.nr = s.nr; pr.nl = s.nl; pr.ns = s.ns; pr.bz = s.bz; pr.bp = s.bp;
  pr.ir2c = f.mat.ir2c; pr.ir2m = f.mat.ir2m; pr.nxy = f.mat.elem.size(); pr.K = f.k; pr.beta_minus = f.beta_minus; pr.use_p = use_p;
  const int nbx = s.nb;
  std::vector<size_t> boff(nbx + 1, 0);
  for (int ib = 0; ib < nbx; ib++) boff[ib + 1] = boff[ib] + (size_t)s.id[ib] * s.id[ib];
  const bool hot = s.ft_active && s.temper > 1e-12;
  for (int it = 0; it < 2; it++) {
    pr.rho[it].assign(boff[nbx], 0.0);
    for (int ib = 0; ib < nbx; ib++) {
      const int nd = s.id[ib];
      double* R = pr.rho[it].data() + boff[ib];
      for (int kk = s.ka[it][ib]; kk < s.ka[it][ib] + s.kd[it][ib]; kk++) {
        const double* V = s.V[it].data() + s.Kpwi[it][kk];
        const double* U = s.U[it].data() + s.Kpwi[it][kk];
        const double fT = hot ? 0.5 * (1.0 - std::tanh(0.5 * s.E[it][s.Kqp[it][kk] - 1] / s.temper)) : 0.0;
        for (int n2 = 0; n2 < nd; n2++)
          for (int n1 = 0; n1 < nd; n1++) R[n1 + (size_t)n2 * nd] += V[n1] * (1.0 - fT) * V[n2] + U[n1] * fT * U[n2];
      }
      if (s.keyblo[it] && s.blo_block[it] == ib + 1 && s.kd[it][ib] > 0) {
        if (s.blok1k2d[it] <= 0) throw std::runtime_error("two-body-current field generator: no blocking candidate found");
        const double* V = s.V[it].data() + s.Kpwi[it][s.blok1k2d[it] - 1];
        const double* U = s.U[it].data() + s.Kpwi[it][s.blok1k2d[it] - 1];
        for (int n2 = 0; n2 < nd; n2++)
          for (int n1 = 0; n1 < nd; n1++) R[n1 + (size_t)n2 * nd] += 0.5 * (-V[n1] * V[n2] + U[n1] * U[n2]);
      }
    }
  }
  return pr;
}

TbcField generate_two_body_current_field(const HfbSolution& s, const FamBasis& b, const ExtField& f, bool use_p) {
  if (b.nb != 2 * s.nb) throw std::runtime_error("two-body-current field generator: basis and HFB solution do not match");
  return generate_two_body_current_field(tbc_problem_from(s, f, use_p));
}

// write_tbc (pnfam_storage.f90:488-559): same records, so that the reference and this library read each other's files
void write_tbc(const std::string& path, const FamBasis& b, const FamInput& in, const ExtField& f, const TwoBody& tb, const TbcField& fld) {
  // unique temporary name: the ranks of a sharded run that own points of the same operator may generate the same file
  const std::string tmp = path + ".tmp." + std::to_string((long)getpid()) + "." + std::to_string((unsigned long)omp_get_thread_num());
  {
    std::ofstream os(tmp, std::ios::binary | std::ios::trunc);
    if (!os) throw std::runtime_error("cannot write " + tmp);
    auto rec = [&](const void* p, size_t n) {
      const int32_t m = (int32_t)n;
      os.write(reinterpret_cast<const char*>(&m), 4);
      os.write(reinterpret_cast<const char*>(p), n);
      os.write(reinterpret_cast<const char*>(&m), 4);
    };
    auto rec_i = [&](std::initializer_list<int32_t> v) { std::vector<int32_t> a(v); rec(a.data(), a.size() * 4); };
    rec_i({1});                                       // VERSION_DATA (write_version, pnfam_storage.f90:139-160)
    rec("extf_2bc", 8);
    rec_i({1});                                       // use_hblas
    rec_i({in.two_body_current_mode});
    rec_i({in.two_body_current_usep ? 1 : 0});
    rec(in.two_body_current_lecs, 3 * sizeof(double));
    rec_i({b.nb, b.dqp, (int32_t)f.mat.elem.size(), -1});
    char label[80];
    std::memset(label, ' ', sizeof label);
    std::memcpy(label, f.label.data(), std::min<size_t>(f.label.size(), sizeof label));
    rec(label, sizeof label);
    rec_i({f.k});
    rec_i({f.rank});
    rec_i({f.beta_minus ? 1 : 0});
    rec_i({f.parity_even ? 1 : 0});
    int32_t u[7];
    for (int i = 0; i < 7; i++) u[i] = tb.u[i];
    rec(u + 1, 6 * 4);
    for (int o = 0; o < (in.two_body_current_usep ? 6 : 4); o++) rec(fld.c[o].data(), fld.c[o].size() * sizeof(double));
    rec_i({0});                                       // nxterms
    if (!os) throw std::runtime_error("error writing " + tmp);
  }
  if (std::rename(tmp.c_str(), path.c_str()) != 0) throw std::runtime_error("cannot rename " + tmp);
}

}  // namespace pnfam

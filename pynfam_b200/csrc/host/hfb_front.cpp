// HFB reconstruction front-end (CPU, once per nucleus).  See hfb_front.hpp.
//
// Follows, step by step, what the reference executes when pnfam_main.x calls HFBTHO with
// number_iterations = 0 (exes/pnfam/hfbtho_interface.f90:43-203):
//   read_data           hfbtho_io.f90:487-738          -> HelData::read
//   gaupol / optHFBTHO  hfbtho_solver.f90:3463-3542, 4237-4313 -> build_tables
//   coordinateLST       hfbtho_solver.f90:1215-1229    -> grid geometry
//   gamdel              hfbtho_solver.f90:5145-5300    -> gamdel
//   hfbdiag + ALambda   hfbtho_solver.f90:1441-2075    -> hfbdiag / alambda
//   DENSIT (rho only)   hfbtho_solver.f90:4318-4734    -> densit_rho
// Written from the algorithm's description; data layouts are our own.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include "hfb_front.hpp"

#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <limits>
#include <numeric>

namespace pnfam {

// ---------------------------------------------------------------------------------------------
// .hel reader
// ---------------------------------------------------------------------------------------------
HelData HelData::read(const std::string& path) {
  FortUnformatted f(path);
  FortRecord r;
  HelData h;
  if (!f.next(r)) throw std::runtime_error("empty .hel file");
  h.version = r.get<int32_t>();
  if (h.version != 5) throw std::runtime_error(".hel: unsupported VERSION_DATA (need 5)");
  bool set_neck = false, pairing_reg = false, coll_inertia = false;
  int switch_to_tho = 0;
  auto need = [&](FortRecord& rec) {
    if (!f.next(rec)) throw std::runtime_error(".hel: truncated file");
  };
  while (f.next(r)) {
    if (r.size() != 8) continue;  // not a keyword record (skip unknown payloads)
    std::string key = r.get_str(8);
    if (key == "Metadata") {
      need(r); h.Z = r.get<int32_t>(); h.N = r.get<int32_t>();
      need(r); coll_inertia = r.get<int32_t>() != 0; r.get<int32_t>(); pairing_reg = r.get<int32_t>() != 0;
      need(r); switch_to_tho = r.get<int32_t>(); r.get<int32_t>();
      h.set_temperature = r.get<int32_t>() != 0; set_neck = r.get<int32_t>() != 0;
    } else if (key == "SkyFunct") {
      need(r);
      r.get<int32_t>(); r.get<int32_t>();
      h.use_j2terms = r.get<int32_t>() != 0; r.get<int32_t>();
      h.finite_range = r.get<int32_t>() != 0;
      h.skyrme = r.get_str(30);
      while (!h.skyrme.empty() && h.skyrme.back() == ' ') h.skyrme.pop_back();
      need(r);
      { auto nm = r.get_vec<double>(9); h.rho_nm = nm[3]; }
      need(r);
      r.get_array(h.Crho, 2); r.get_array(h.Cdrho, 2); r.get_array(h.Ctau, 2); r.get_array(h.CrDr, 2);
      r.get_array(h.CrdJ, 2); r.get_array(h.CJ, 2); r.get_array(h.CpV0, 2); r.get_array(h.CpV1, 2);
      h.sigma = r.get<double>();
      need(r);
      h.hbzero = r.get<double>(); h.hb0 = r.get<double>(); h.hb0n = r.get<double>(); h.hb0p = r.get<double>();
    } else if (key == "HO-Basis") {
      need(r); h.b0 = r.get<double>(); h.bz = r.get<double>(); h.bp = r.get<double>();
      need(r);
      h.n00 = r.get<int32_t>(); h.nb = r.get<int32_t>(); h.nt = r.get<int32_t>();
      h.ngh = r.get<int32_t>(); h.ngl = r.get<int32_t>(); h.nleg = r.get<int32_t>();
      need(r);
      h.xh = r.get_vec<double>(h.ngh); h.xl = r.get_vec<double>(h.ngl);
      h.wh = r.get_vec<double>(h.ngh); h.wl = r.get_vec<double>(h.ngl);
    } else if (key == "QuantNum") {
      need(r);
      h.id = r.get_vec<int32_t>(h.nb);
      h.nr.resize(h.nt); h.nz.resize(h.nt); h.nl.resize(h.nt); h.ns.resize(h.nt);
      int ib = 0;
      for (int b = 0; b < h.nb; b++)
        for (int n = 0; n < h.id[b]; n++) {
          need(r);
          if (ib >= h.nt) throw std::runtime_error(".hel: more states than nt");
          h.nr[ib] = r.get<int32_t>(); h.nz[ib] = r.get<int32_t>();
          h.nl[ib] = r.get<int32_t>(); h.ns[ib] = r.get<int32_t>();
          ib++;
        }
    } else if (key == "Various.") {
      need(r);
      h.si = r.get<double>(); h.etot = r.get<double>(); r.get_array(h.rms, 3);
      h.bet = r.get<double>(); h.xmix = r.get<double>();
      need(r);
      h.pwi = r.get<double>(); r.get_array(h.del, 2); r.get_array(h.ept, 3); r.get_array(h.ala, 2);
      r.get_array(h.ala2, 2); r.get_array(h.alast, 2);
      need(r);
      r.get_array(h.tz, 2);
    } else if (key == "Constrai") {
      for (int i = 0; i < 5; i++) need(r);
      if (set_neck) for (int i = 0; i < 3; i++) need(r);
      need(r);
    } else if (key == "Densits.") {
      need(r); h.ro = r.get_vec<double>(r.size() / 8);
      need(r); h.aka = r.get_vec<double>(r.size() / 8);
    } else if (key == "FieldsN." || key == "FieldsP.") {
      int it = key == "FieldsN." ? 0 : 1;
      for (int i = 0; i < 11; i++) { need(r); h.fld[it][i] = r.get_vec<double>(r.size() / 8); }
    } else if (key == "Blocking") {
      need(r); h.bloall = r.get<int32_t>();
      need(r);
      size_t n = (size_t)(h.bloall + 1) * 2;
      h.bloblo = r.get_vec<int32_t>(n); h.blo123 = r.get_vec<int32_t>(n); h.blok1k2 = r.get_vec<int32_t>(n);
      h.blomax[0] = r.get<int32_t>(); h.blomax[1] = r.get<int32_t>();
      h.bloqpdif = r.get_vec<double>(n);
      h.has_blocking = true;
    } else if (key == "Blk-rest") {
      need(r);
      h.blocking_never_done[0] = r.get<int32_t>() != 0;
      h.blocking_never_done[1] = r.get<int32_t>() != 0;
    } else if (key == "HFBmatrX") {
      h.has_hfb_matrix = true;
      need(r); need(r);
    }
    // THObasis / Regular. / GognyVNN / CollMass / Temperat payload records are not 8 bytes long
    // in practice and are skipped by the size test above; features that need them are rejected
    // in HfbSolution::build.
  }
  (void)pairing_reg; (void)coll_inertia; (void)switch_to_tho;
  if (h.nt <= 0 || h.fld[0][0].empty() || h.fld[1][0].empty())
    throw std::runtime_error(".hel: missing basis or field records");
  return h;
}

// ---------------------------------------------------------------------------------------------
// hfbtho_NAMELIST.dat
// ---------------------------------------------------------------------------------------------
HfbInput HfbInput::read(const std::string& path) {
  Namelist nl = Namelist::parse_file(path);
  HfbInput in;
  in.n_shells = std::abs(nl.get_int("hfbtho_general", "number_of_shells", 10));
  in.proton_number = nl.get_int("hfbtho_general", "proton_number", 0);
  in.neutron_number = nl.get_int("hfbtho_general", "neutron_number", 0);
  in.type_of_calculation = nl.get_int("hfbtho_general", "type_of_calculation", 1);
  in.functional = nl.get_string("hfbtho_functional", "functional", "SLY4");
  in.user_pairing = nl.get_bool("hfbtho_pairing", "user_pairing", false);
  in.vpair_n = nl.get_double("hfbtho_pairing", "vpair_n", -300.0);
  in.vpair_p = nl.get_double("hfbtho_pairing", "vpair_p", -300.0);
  in.pairing_cutoff = nl.get_double("hfbtho_pairing", "pairing_cutoff", 60.0);
  in.pairing_feature = nl.get_double("hfbtho_pairing", "pairing_feature", 0.5);
  auto nb = nl.get_ints("hfbtho_blocking", "neutron_blocking", {0, 0, 0, 0, 0});
  auto pb = nl.get_ints("hfbtho_blocking", "proton_blocking", {0, 0, 0, 0, 0});
  for (int i = 0; i < 5; i++) { in.neutron_blocking[i] = nb[i]; in.proton_blocking[i] = pb[i]; }
  in.set_temperature = nl.get_bool("hfbtho_temperature", "set_temperature", false);
  in.temperature = nl.get_double("hfbtho_temperature", "temperature", 0.0);
  in.force_parity = nl.get_bool("hfbtho_debug", "force_parity", true);
  in.compatibility_hfodd = nl.get_bool("hfbtho_debug", "compatibility_hfodd", false);
  return in;
}

// ---------------------------------------------------------------------------------------------
// Symmetric eigen-solver: Householder reduction to tridiagonal form followed by the implicit QL
// algorithm (the classical tred2/tql2 pair), then an ascending sort.
// ---------------------------------------------------------------------------------------------
void sym_eig(int n, const double* a_lower, double* w, double* zz) {
  if (n <= 0) return;
  std::vector<double> e(n, 0.0);
  auto Z = [&](int i, int j) -> double& { return zz[(size_t)j * n + i]; };
  for (int j = 0; j < n; j++)
    for (int i = j; i < n; i++) { Z(i, j) = a_lower[(size_t)j * n + i]; Z(j, i) = Z(i, j); }
  if (n == 1) { w[0] = Z(0, 0); Z(0, 0) = 1.0; return; }
  double* d = w;
  // --- tridiagonalisation -------------------------------------------------------------------
  for (int i = n - 1; i >= 1; i--) {
    int l = i - 1;
    double h = 0.0, scale = 0.0;
    if (l > 0) {
      for (int k = 0; k <= l; k++) scale += std::fabs(Z(i, k));
      if (scale == 0.0) {
        e[i] = Z(i, l);
      } else {
        for (int k = 0; k <= l; k++) { Z(i, k) /= scale; h += Z(i, k) * Z(i, k); }
        double f = Z(i, l);
        double g = (f >= 0.0) ? -std::sqrt(h) : std::sqrt(h);
        e[i] = scale * g;
        h -= f * g;
        Z(i, l) = f - g;
        f = 0.0;
        for (int j = 0; j <= l; j++) {
          Z(j, i) = Z(i, j) / h;
          g = 0.0;
          for (int k = 0; k <= j; k++) g += Z(j, k) * Z(i, k);
          for (int k = j + 1; k <= l; k++) g += Z(k, j) * Z(i, k);
          e[j] = g / h;
          f += e[j] * Z(i, j);
        }
        double hh = f / (h + h);
        for (int j = 0; j <= l; j++) {
          f = Z(i, j);
          e[j] = g = e[j] - hh * f;
          for (int k = 0; k <= j; k++) Z(j, k) -= (f * e[k] + g * Z(i, k));
        }
      }
    } else {
      e[i] = Z(i, l);
    }
    d[i] = h;
  }
  d[0] = 0.0;
  e[0] = 0.0;
  for (int i = 0; i < n; i++) {
    int l = i - 1;
    if (d[i] != 0.0) {
      for (int j = 0; j <= l; j++) {
        double g = 0.0;
        for (int k = 0; k <= l; k++) g += Z(i, k) * Z(k, j);
        for (int k = 0; k <= l; k++) Z(k, j) -= g * Z(k, i);
      }
    }
    d[i] = Z(i, i);
    Z(i, i) = 1.0;
    for (int j = 0; j <= l; j++) Z(j, i) = Z(i, j) = 0.0;
  }
  // --- implicit QL ----------------------------------------------------------------------------
  for (int i = 1; i < n; i++) e[i - 1] = e[i];
  e[n - 1] = 0.0;
  const double eps = std::numeric_limits<double>::epsilon();
  for (int l = 0; l < n; l++) {
    int iter = 0, m;
    do {
      for (m = l; m < n - 1; m++) {
        double dd = std::fabs(d[m]) + std::fabs(d[m + 1]);
        if (std::fabs(e[m]) <= eps * dd) break;
      }
      if (m != l) {
        if (iter++ == 200) throw std::runtime_error("sym_eig: QL did not converge");
        double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
        double r = std::hypot(g, 1.0);
        g = d[m] - d[l] + e[l] / (g + (g >= 0.0 ? std::fabs(r) : -std::fabs(r)));
        double s = 1.0, c = 1.0, p = 0.0;
        int i;
        for (i = m - 1; i >= l; i--) {
          double f = s * e[i], b = c * e[i];
          e[i + 1] = (r = std::hypot(f, g));
          if (r == 0.0) { d[i + 1] -= p; e[m] = 0.0; break; }
          s = f / r; c = g / r;
          g = d[i + 1] - p;
          r = (d[i] - g) * s + 2.0 * c * b;
          d[i + 1] = g + (p = s * r);
          g = c * r - b;
          for (int k = 0; k < n; k++) {
            f = Z(k, i + 1);
            Z(k, i + 1) = s * Z(k, i) + c * f;
            Z(k, i) = c * Z(k, i) - s * f;
          }
        }
        if (r == 0.0 && i >= l) continue;
        d[l] -= p; e[l] = g; e[m] = 0.0;
      }
    } while (m != l);
  }
  // --- ascending sort -------------------------------------------------------------------------
  std::vector<int> ord(n);
  std::iota(ord.begin(), ord.end(), 0);
  std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return d[a] < d[b]; });
  std::vector<double> wd(d, d + n), zc((size_t)n * n);
  for (int j = 0; j < n; j++) {
    w[j] = wd[ord[j]];
    std::copy(zz + (size_t)ord[j] * n, zz + (size_t)(ord[j] + 1) * n, zc.begin() + (size_t)j * n);
  }
  std::copy(zc.begin(), zc.end(), zz);
}

// ---------------------------------------------------------------------------------------------
// Basis tables on the quadrature grid
// ---------------------------------------------------------------------------------------------
static void build_tables(HfbSolution& s, const HelData& h) {
  const int ngh = h.ngh, ngl = h.ngl, nghl = ngh * ngl, nt = h.nt;
  int nzm = 0, nrm = 0, nlm = 0;
  for (int i = 0; i < nt; i++) {
    nzm = std::max(nzm, h.nz[i]); nrm = std::max(nrm, h.nr[i]); nlm = std::max(nlm, h.nl[i]);
  }
  // factorial-type tables (hfbtho_math.f90:327-344): sq(n)=sqrt(n), sqi=1/sq, wfi(n)=1/sqrt(n!)
  const int ig = 170;
  std::vector<double> sq(ig + 1), sqi(ig + 1), wf(ig + 1), wfi(ig + 1);
  sq[0] = 0; sqi[0] = 1e30; wf[0] = 1; wfi[0] = 1;
  for (int i = 1; i <= ig; i++) {
    sq[i] = std::sqrt((double)i); sqi[i] = 1.0 / sq[i];
    wf[i] = sq[i] * wf[i - 1]; wfi[i] = 1.0 / wf[i];
  }
  const double pi = 4.0 * std::atan(1.0);
  // z direction: qh(n,ih), qh1(n,ih)  (hfbtho_solver.f90:3494-3503)
  const int nzd = std::max(nzm, 1) + 1;
  std::vector<double> qh((size_t)nzd * ngh), qh1((size_t)nzd * ngh);
  auto QH = [&](int n, int ih) -> double& { return qh[(size_t)ih * nzd + n]; };
  auto QH1 = [&](int n, int ih) -> double& { return qh1[(size_t)ih * nzd + n]; };
  const double w4pii = std::pow(pi, -0.25);
  for (int ih = 0; ih < ngh; ih++) {
    double z = h.xh[ih];
    double w0 = w4pii * std::exp(-0.5 * z * z);
    w0 = w0 * std::sqrt(h.wh[ih]);
    QH(0, ih) = w0; QH(1, ih) = sq[2] * w0 * z;
    QH1(0, ih) = -w0 * z; QH1(1, ih) = sq[2] * w0 * (1.0 - z * z);
    for (int n = 2; n <= nzm; n++) {
      QH(n, ih) = sqi[n] * (sq[2] * z * QH(n - 1, ih) - sq[n - 1] * QH(n - 2, ih));
      QH1(n, ih) = sq[n + n] * QH(n - 1, ih) - z * QH(n, ih);
    }
  }
  // perpendicular direction: ql(n,l,il), ql1(n,l,il)  (hfbtho_solver.f90:3529-3542)
  const int nrd = std::max(nrm, 1) + 1, nld = nlm + 1;
  std::vector<double> ql((size_t)nrd * nld * ngl), ql1((size_t)nrd * nld * ngl);
  auto QL = [&](int n, int l, int il) -> double& { return ql[((size_t)il * nld + l) * nrd + n]; };
  auto QL1 = [&](int n, int l, int il) -> double& { return ql1[((size_t)il * nld + l) * nrd + n]; };
  for (int il = 0; il < ngl; il++) {
    double x = h.xl[il];
    double w00 = sq[2] * std::exp(-0.5 * x);
    for (int l = 0; l <= nlm; l++) {
      double w0 = w00 * std::sqrt(0.5 * h.wl[il] * std::pow(x, l));
      QL(0, l, il) = wfi[l] * w0;
      QL(1, l, il) = (l + 1 - x) * wfi[l + 1] * w0;
      QL1(0, l, il) = (l - x) * wfi[l] * w0;
      QL1(1, l, il) = ((double)(l * l + l) - x * (double)(l + l + 3) + x * x) * wfi[l + 1] * w0;
      for (int n = 2; n <= nrm; n++) {
        double dsq = sq[n] * sq[n + l], d1 = (double)(n + n + l - 1) - x;
        double d2 = sq[n - 1] * sq[n - 1 + l], d3 = n + n + l - x, d4 = 2.0 * dsq;
        QL(n, l, il) = (d1 * QL(n - 1, l, il) - d2 * QL(n - 2, l, il)) / dsq;
        QL1(n, l, il) = d3 * QL(n, l, il) - d4 * QL(n - 1, l, il);
      }
    }
  }
  // grid geometry (coordinateLST, HO branch) and optHFBTHO tables
  s.y.resize(nghl); s.z.resize(nghl); s.wdcor.resize(nghl); s.wdcori.resize(nghl);
  for (int il = 0; il < ngl; il++)
    for (int ih = 0; ih < ngh; ih++) {
      int i = ih + il * ngh;
      s.z[i] = h.bz * h.xh[ih];
      s.wdcor[i] = pi * h.wh[ih] * h.wl[il] * h.bz * h.bp * h.bp;
      s.wdcori[i] = 1.0 / s.wdcor[i];
      double yi = std::sqrt(h.xl[il]) * h.bp;
      s.y[i] = 1.0 / yi;
    }
  s.qhla.assign((size_t)nt * nghl, 0.0); s.fi1r = s.qhla; s.fi1z = s.qhla; s.fi2d = s.qhla;
  const double bpi = 1.0 / h.bp, bpi2 = bpi * bpi, bzi = 1.0 / h.bz, bzi2 = bzi * bzi;
  // separable factors (consumed by the sum-factorised GPU kernels)
  s.sep_nzrows = nzm + 1;
  s.sep_z.assign((size_t)3 * s.sep_nzrows * ngh, 0.0);
  s.sep_r.assign((size_t)3 * nt * ngl, 0.0);
  for (int n = 0; n <= nzm; n++)
    for (int ih = 0; ih < ngh; ih++) {
      const double xh2 = h.xh[ih] * h.xh[ih];
      s.sep_z[((size_t)0 * s.sep_nzrows + n) * ngh + ih] = QH(n, ih);
      s.sep_z[((size_t)1 * s.sep_nzrows + n) * ngh + ih] = bzi * QH1(n, ih);
      s.sep_z[((size_t)2 * s.sep_nzrows + n) * ngh + ih] = (xh2 - (double)(n + n + 1)) * bzi2 * QH(n, ih);
    }
  for (int ja = 0; ja < nt; ja++) {
    const int nla = h.nl[ja], nra = h.nr[ja];
    const double sml2 = (double)(nla * nla), cnraa = nra + nra + nla + 1;
    for (int il = 0; il < ngl; il++) {
      const double v2 = 0.5 / h.xl[il], v4 = v2 * v2, qla = QL(nra, nla, il);
      s.sep_r[((size_t)0 * nt + ja) * ngl + il] = qla;
      s.sep_r[((size_t)1 * nt + ja) * ngl + il] = (2.0 * std::sqrt(h.xl[il]) * bpi) * (QL1(nra, nla, il) * v2);
      s.sep_r[((size_t)2 * nt + ja) * ngl + il] = 4.0 * (0.25 - cnraa * v2 + sml2 * v4) * h.xl[il] * bpi2 * qla;
    }
  }
  for (int ja = 0; ja < nt; ja++) {
    const int nla = h.nl[ja], nra = h.nr[ja], nza = h.nz[ja];
    const double sml2 = (double)(nla * nla), cnzaa = nza + nza + 1, cnraa = nra + nra + nla + 1;
    for (int il = 0; il < ngl; il++) {
      const double v2 = 0.5 / h.xl[il], v4 = v2 * v2;
      for (int ih = 0; ih < ngh; ih++) {
        const int ihil = ih + il * ngh;
        const double xh2 = h.xh[ih] * h.xh[ih];
        const double qha = QH(nza, ih), qla = QL(nra, nla, il), qhla = qha * qla;
        const double qhl1a = qha * QL1(nra, nla, il) * v2, qh1la = QH1(nza, ih) * qla;
        const size_t o = (size_t)ja * nghl + ihil;
        s.qhla[o] = qhla;
        s.fi1r[o] = (2.0 * std::sqrt(h.xl[il]) * bpi) * qhl1a;
        s.fi1z[o] = bzi * qh1la;
        s.fi2d[o] = ((xh2 - cnzaa) * bzi2 + 4.0 * (0.25 - cnraa * v2 + sml2 * v4) * h.xl[il] * bpi2) * qhla;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// gamdel: fields on the grid -> HFB matrix (packed lower triangle per block, row n1 >= col n2,
// running index as in the reference: n1 outer, n2 = 1..n1 inner).
// ---------------------------------------------------------------------------------------------
static void gamdel(HfbSolution& s, const HelData& h) {
  const int nghl = s.nghl;
  size_t nhh = 0;
  std::vector<size_t> off(s.nb);
  for (int ib = 0; ib < s.nb; ib++) { off[ib] = nhh; nhh += (size_t)s.id[ib] * (s.id[ib] + 1) / 2; }
  for (int it = 0; it < 2; it++) { s.hmat[it].assign(nhh, 0.0); s.dmat[it].assign(nhh, 0.0); }
  enum { V = 0, VHB, VR, VZ, VD, VS, VSFIZ, VSZFI, VSFIR, VSRFI, DV };
#pragma omp parallel for schedule(dynamic)
  for (int ib = 0; ib < s.nb; ib++) {
    const int nd = s.id[ib], im = s.ia[ib];
    std::vector<double> fiu(nd), fid(nd), fiur(nd), fidr(nd), fiuz(nd), fidz(nd), fiud2(nd), fidd2(nd);
    // Lambda of the spin-up / spin-down members of this Omega block (XLAM / XLAP)
    double xlam = 0, xlap = 0;
    {
      // Omega = Lambda + 1/2 (up) = Lambda' - 1/2 (down): take from the quantum numbers
      int om2 = 2 * s.nl[im] + s.ns[im];  // 2*Omega
      xlam = (om2 - 1) / 2; xlap = (om2 + 1) / 2;
    }
    const double xlam2 = xlam * xlam, xlap2 = xlap * xlap;
    double* hn = s.hmat[0].data() + off[ib];
    double* hp = s.hmat[1].data() + off[ib];
    double* dn = s.dmat[0].data() + off[ib];
    double* dp = s.dmat[1].data() + off[ib];
    for (int ihil = 0; ihil < nghl; ihil++) {
      const double y = s.y[ihil], xlamy = xlam * y, xlapy = xlap * y, xlampy = xlamy + xlapy;
      const double y2 = y * y, xlamy2 = xlam2 * y2, xlapy2 = xlap2 * y2;
      const double vn = h.fld[0][V][ihil], vrn = h.fld[0][VR][ihil], vzn = h.fld[0][VZ][ihil];
      const double vdn = h.fld[0][VD][ihil], vsn = h.fld[0][VS][ihil], vhbn = h.fld[0][VHB][ihil];
      const double vSRFIn = h.fld[0][VSRFI][ihil], vSFIRn = h.fld[0][VSFIR][ihil];
      const double vSFIZn = h.fld[0][VSFIZ][ihil], vSZFIn = h.fld[0][VSZFI][ihil];
      const double vp = h.fld[1][V][ihil], vrp = h.fld[1][VR][ihil], vzp = h.fld[1][VZ][ihil];
      const double vdp = h.fld[1][VD][ihil], vsp = h.fld[1][VS][ihil], vhbp = h.fld[1][VHB][ihil];
      const double vSRFIp = h.fld[1][VSRFI][ihil], vSFIRp = h.fld[1][VSFIR][ihil];
      const double vSFIZp = h.fld[1][VSFIZ][ihil], vSZFIp = h.fld[1][VSZFI][ihil];
      const double dvn = h.fld[0][DV][ihil], dvp = h.fld[1][DV][ihil];
      for (int n1 = 0; n1 < nd; n1++) {
        const int ja = im + n1, nsa = s.ns[ja];
        const double ssu = std::max(nsa, 0), ssd = std::max(-nsa, 0);
        const size_t o = (size_t)ja * nghl + ihil;
        const double q = s.qhla[o], f1r = s.fi1r[o], f1z = s.fi1z[o], f2d = s.fi2d[o];
        fiu[n1] = q * ssu; fiur[n1] = f1r * ssu; fiuz[n1] = f1z * ssu; fiud2[n1] = (f2d - xlamy2 * q) * ssu;
        fid[n1] = q * ssd; fidr[n1] = f1r * ssd; fidz[n1] = f1z * ssd; fidd2[n1] = (f2d - xlapy2 * q) * ssd;
      }
      size_t i = 0;
      for (int n1 = 0; n1 < nd; n1++) {
        const int nsa = s.ns[im + n1];
        const double FIUN1 = fiu[n1], FIURN1 = fiur[n1], FIUZN1 = fiuz[n1], FIUD2N1 = fiud2[n1];
        const double FIDN1 = fid[n1], FIDRN1 = fidr[n1], FIDZN1 = fidz[n1], FIDD2N1 = fidd2[n1];
        for (int n2 = 0; n2 <= n1; n2++, i++) {
          const int nsb = s.ns[im + n2];
          if (nsa + nsb != 0) {
            double vh, hbh, vdh, snr, snz, vsh, sfiz;
            if (nsb > 0) {  // up-up
              const double FIUN2 = fiu[n2], FIURN2 = fiur[n2], FIUD2N2 = fiud2[n2], FIUZN2 = fiuz[n2];
              vh = FIUN1 * FIUN2;
              hbh = vh * xlamy2 + FIURN1 * FIURN2 + FIUZN1 * FIUZN2;
              vdh = hbh + hbh + FIUN1 * FIUD2N2 + FIUN2 * FIUD2N1;
              snr = FIURN1 * FIUN2 + FIURN2 * FIUN1;
              snz = FIUZN1 * FIUN2 + FIUZN2 * FIUN1;
              vsh = snr * xlamy;
              sfiz = (vh + vh) * xlamy;
            } else {  // down-down
              const double FIDN2 = fid[n2], FIDRN2 = fidr[n2], FIDZN2 = fidz[n2], FIDD2N2 = fidd2[n2];
              vh = FIDN1 * FIDN2;
              hbh = vh * xlapy2 + FIDRN1 * FIDRN2 + FIDZN1 * FIDZN2;
              vdh = hbh + hbh + FIDN1 * FIDD2N2 + FIDN2 * FIDD2N1;
              snr = FIDRN1 * FIDN2 + FIDRN2 * FIDN1;
              snz = FIDZN1 * FIDN2 + FIDZN2 * FIDN1;
              vsh = -snr * xlapy;
              sfiz = -(vh + vh) * xlapy;
            }
            hn[i] = hn[i] + vSFIZn * sfiz + vh * vn + snr * vrn + snz * vzn + vdh * vdn + vsh * vsn + hbh * vhbn;
            hp[i] = hp[i] + vSFIZp * sfiz + vh * vp + snr * vrp + snz * vzp + vdh * vdp + vsh * vsp + hbh * vhbp;
            dn[i] = dn[i] + vh * dvn;
            dp[i] = dp[i] + vh * dvp;
          } else {
            double vsh, srfi, sfir, szfi;
            if (nsb > 0) {  // down-up
              const double FIUN2 = fiu[n2], FIURN2 = fiur[n2], FIUZN2 = fiuz[n2];
              const double fitw3 = -FIDZN1 * FIUN2, fitw4 = FIUZN2 * FIDN1;
              vsh = -FIDRN1 * FIUZN2 + FIURN2 * FIDZN1 + fitw3 * xlamy - fitw4 * xlapy;
              srfi = -FIDRN1 * FIUN2 + FIURN2 * FIDN1;
              sfir = FIDN1 * FIUN2 * xlampy;
              szfi = fitw3 + fitw4;
            } else {  // up-down
              const double FIDN2 = fid[n2], FIDRN2 = fidr[n2], FIDZN2 = fidz[n2];
              const double fitw3 = -FIDZN2 * FIUN1, fitw4 = FIUZN1 * FIDN2;
              vsh = FIURN1 * FIDZN2 - FIDRN2 * FIUZN1 - fitw4 * xlapy + fitw3 * xlamy;
              srfi = FIURN1 * FIDN2 - FIDRN2 * FIUN1;
              sfir = FIUN1 * FIDN2 * xlampy;
              szfi = fitw3 + fitw4;
            }
            hn[i] = hn[i] + vsh * vsn + vSRFIn * srfi + vSFIRn * sfir + vSZFIn * szfi;
            hp[i] = hp[i] + vsh * vsp + vSRFIp * srfi + vSFIRp * sfir + vSZFIp * szfi;
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// ALambda: Fermi level from the BCS-like reference spectrum (erhfb, drhfb)
// ---------------------------------------------------------------------------------------------
static void alambda(double& al, int kl, const std::vector<double>& erhfb, std::vector<double>& drhfb, double tz,
                    double cpv0, int blok1k2d, double temper, std::vector<double>* f_T) {
  const bool hot = f_T != nullptr && temper > 1e-12;
  if (cpv0 == 0.0) {
    int ntz = (int)(tz + 0.1); ntz /= 2;
    std::vector<double> d(erhfb.begin(), erhfb.begin() + kl);
    std::sort(d.begin(), d.end());
    if (ntz < kl) al = 0.5 * (d[ntz - 1] + d[ntz]);
    else al = d[ntz - 1] + 0.001;
    return;
  }
  double xinf = -1000.0, xsup = 1000.0;
  for (int lit = 1; lit <= 500; lit++) {
    double sn = 0, dez = 0, dfz = 0;
    for (int i = 0; i < kl; i++) {
      double vh = 0, dvh = 0, fT = 0, dfT = 0;
      const double y = erhfb[i] - al, a = y * y + drhfb[i] * drhfb[i], b = std::sqrt(a);
      if (hot) {
        fT = 0.5 * (1.0 - std::tanh(0.5 * b / temper));
        dfT = y / b / temper * fT * (1.0 - fT);
        (*f_T)[i] = fT;
      }
      if (b > 0) vh = 0.5 * (1.0 - y / b);
      if (vh < 1e-12) vh = 0;
      if ((vh - 1.0) > 1e-12) vh = 1.0;
      if (b > 0) dvh = 0.5 * drhfb[i] * drhfb[i] / (a * b);
      if (i + 1 == blok1k2d) { vh = 0.5; dvh = 0; }
      sn += 2.0 * vh + 2.0 * (1.0 - 2.0 * vh) * fT;
      dez += 2.0 * (1.0 - 2.0 * fT) * dvh;
      dfz += 2.0 * (1.0 - 2.0 * vh) * dfT;
    }
    dez += dfz;
    const double ez = sn - tz, absez = std::fabs(ez) / tz;
    if (ez < 0) xinf = std::max(xinf, al);
    else xsup = std::min(xsup, al);
    if (lit == 1) {
      if (absez <= 0.10) al = al - ez;
      else al = al - 0.10 * (ez >= 0 ? 1.0 : -1.0);
    } else {
      al = al - ez / (dez + 1e-20);
    }
    if (al < xinf || al > xsup) al = 0.5 * (xinf + xsup);
    if (absez <= 1e-10) return;
  }
}

// ---------------------------------------------------------------------------------------------
// hfbdiag: block diagonalisation with the particle-number (lambda) refinement loop
// ---------------------------------------------------------------------------------------------
static void hfbdiag(HfbSolution& s, const HelData& h, const HfbInput& in, int it, int iparenti, bool basis_hfodd) {
  const int nb = s.nb, nt = s.nt;
  const double cutoff_tol = 1e-6;  // hfbtho_variables.f90:253
  std::vector<size_t> off(nb), offuv(nb);
  size_t nhh = 0, nuv = 0;
  for (int ib = 0; ib < nb; ib++) {
    off[ib] = nhh; nhh += (size_t)s.id[ib] * (s.id[ib] + 1) / 2;
    offuv[ib] = nuv; nuv += (size_t)s.id[ib] * s.id[ib];
  }
  s.E[it].assign(nt, 0.0); s.U[it].assign(nuv, 0.0); s.V[it].assign(nuv, 0.0);
  s.ka[it].assign(nb, 0); s.kd[it].assign(nb, 0);
  s.Kqp[it].assign(nt, 0); s.Kpwi[it].assign(nt, 0); s.occ[it].assign(nt, 0.0);
  std::vector<std::vector<double>> evec(nb), eval(nb);
  std::vector<double> erhfb(nt), drhfb(nt);
  const double tz = (double)s.npr[it];
  const bool hot = s.ft_active && s.temper > 1e-12;
  const double sitest = std::min(0.10, h.si * 0.010);
  bool norm_to_improve = true;
  int inner = -1;
  double sumnz = 1.0;
  double ala = h.ala[it];
  // blocking state
  int keyblo = 0, ibiblo = 0, blocross = 0;
  bool never_done = h.blocking_never_done[it];
  const int nblo = h.bloall + 1;
  std::vector<double> hfb1;
  const int* nkblo = it == 0 ? in.neutron_blocking : in.proton_blocking;
  int blo123_1 = 0, bloblo_1 = 0;
  if (h.has_blocking) { bloblo_1 = h.bloblo[(size_t)it * nblo + 1]; blo123_1 = h.blo123[(size_t)it * nblo + 1]; }
  int blok1k2d = 0;
  const int blomax = iparenti == 0 ? 0 : h.blomax[it];
  while (norm_to_improve) {
    inner++;
    if (std::fabs(sumnz) < sitest || inner == 20) norm_to_improve = false;
    sumnz = 0.0;
    int kl = 0;
    const double al = ala;
    blok1k2d = 0;
    ibiblo = keyblo ? bloblo_1 : 0;
    // diagonalise all blocks
#pragma omp parallel for schedule(dynamic)
    for (int ib = 0; ib < nb; ib++) {
      const int nd = s.id[ib], nhfb = 2 * nd;
      std::vector<double> m((size_t)nhfb * nhfb, 0.0);
      const double* hh = s.hmat[it].data() + off[ib];
      const double* dd = s.dmat[it].data() + off[ib];
      size_t ibro = 0;
      for (int n1 = 0; n1 < nd; n1++) {
        const int nd1 = n1 + nd;
        for (int n2 = 0; n2 <= n1; n2++, ibro++) {
          const int nd2 = n2 + nd;
          const double hla = hh[ibro], dla = dd[ibro];
          m[(size_t)n2 * nhfb + n1] = hla;      // (n1,n2)
          m[(size_t)n1 * nhfb + nd2] = dla;     // (nd2,n1)
          m[(size_t)n2 * nhfb + nd1] = dla;     // (nd1,n2)
          m[(size_t)nd2 * nhfb + nd1] = -hla;   // (nd1,nd2)
        }
        m[(size_t)n1 * nhfb + n1] -= al;
        m[(size_t)nd1 * nhfb + nd1] += al;
      }
      eval[ib].resize(nhfb); evec[ib].resize((size_t)nhfb * nhfb);
      sym_eig(nhfb, m.data(), eval[ib].data(), evec[ib].data());
    }
    size_t i_uv = 0;
    int i_eqp = 0;
    for (int ib = 0; ib < nb; ib++) {
      const int nd = s.id[ib], nhfb = 2 * nd, i0 = s.ia[ib];
      const std::vector<double>& W = evec[ib];
      auto A = [&](int r, int c) { return W[(size_t)c * nhfb + r]; };  // 0-based
      // ---- blocking (hfbtho_solver.f90:1590-1680) --------------------------------------------
      if (inner == 0) {
        if (never_done) {
          if (iparenti != 0 && keyblo == 0 && nkblo[1] != 0) {
            // requested_blocked_level (hfbtho_solver.f90:7468-7506)
            const int omega = 2 * s.nl[i0] + s.ns[i0];
            if (nkblo[0] == omega) {
              for (int k = 0; k < nd; k++) {
                double s1 = 0; int iqn = i0;
                for (int na = 0; na < nd; na++) {
                  double s2 = std::max(s1, std::max(std::fabs(A(na, k + nd)), std::fabs(A(na + nd, k + nd))));
                  if (s2 > s1) { s1 = s2; iqn = na + i0; }
                }
                const int par = s.npar[iqn] == 1 ? +1 : -1;
                if (nkblo[1] != par) continue;
                if (nkblo[4] != s.nl[iqn]) continue;
                if (nkblo[3] != s.nz[iqn]) continue;
                if (nkblo[2] != s.nz[iqn] + 2 * s.nr[iqn] + s.nl[iqn]) continue;
                keyblo = 1; bloblo_1 = ib + 1; blo123_1 = k + 1;
                break;
              }
              ibiblo = keyblo ? bloblo_1 : 0;
            }
          }
        } else if (iparenti != 0 && keyblo == 0) {
          keyblo = 1;
          ibiblo = bloblo_1;
        }
      }
      int k0 = 0;
      if (ibiblo == ib + 1) {
        if (inner == 0) {
          k0 = blo123_1;
          hfb1.assign(nhfb, 0.0);
          for (int n2 = 0; n2 < nhfb; n2++) hfb1[n2] = A(n2, k0 - 1 + nd);
          blocross = std::min(blomax + 10, nd);
        }
        double s3 = 0;
        for (int n1 = 0; n1 < blocross; n1++) {
          double s1 = 0;
          for (int n2 = 0; n2 < nd; n2++) {
            s1 += std::fabs(hfb1[n2 + nd] * A(n2 + nd, n1 + nd));
            s1 += std::fabs(hfb1[n2] * A(n2, n1 + nd));
          }
          if (s1 > s3) { s3 = s1; k0 = n1 + 1; }
        }
        blo123_1 = k0;
        if (!norm_to_improve)
          for (int n1 = 0; n1 < nhfb; n1++) hfb1[n1] = A(n1, k0 - 1 + nd);
      }
      // ---- quasiparticles of the block --------------------------------------------------------
      const int kaib = kl;
      for (int k = 0; k < nd; k++) {
        const int ndk = k + nd;
        double pn = 0;
        for (int i = 0; i < nd; i++) { const double v = A(i + nd, ndk); pn += v * v; }
        if (k + 1 == k0) {
          for (int i = 0; i < nd; i++) {
            const double hla = A(i + nd, ndk) * A(i + nd, ndk), dla = A(i, ndk) * A(i, ndk);
            pn = pn - 0.5 * (hla - dla);
          }
        }
        const double eqpe = eval[ib][ndk], ela = eqpe * (1.0 - 2.0 * pn);
        const double enb = ela + al, ekb = std::sqrt(std::fabs(eqpe * eqpe - ela * ela));
        double expo = std::numeric_limits<double>::max();
        if (std::fabs(100.0 * (enb - s.pwi)) < std::log(std::numeric_limits<double>::max()))
          expo = std::exp(100.0 * (enb - s.pwi));
        bool lpr_pwi;
        if (basis_hfodd) lpr_pwi = enb <= s.pwi;
        else lpr_pwi = enb <= s.pwi || std::fabs(1.0 / (1.0 + expo)) > cutoff_tol;
        if (!norm_to_improve) {
          s.E[it][i_eqp] = eqpe;
          if (lpr_pwi) { s.Kqp[it][kl] = i_eqp + 1; s.Kpwi[it][kl] = (int)i_uv; }
          double* Uo = s.U[it].data() + i_uv;
          double* Vo = s.V[it].data() + i_uv;
          for (int n2 = 0; n2 < nd; n2++) { Uo[n2] = A(n2, ndk); Vo[n2] = A(n2 + nd, ndk); }
        }
        i_uv += nd;
        i_eqp++;
        const double fT = hot ? 0.5 * (1.0 - std::tanh(0.5 * eqpe / s.temper)) : 0.0;
        if (lpr_pwi) {
          if (k0 == k + 1) blok1k2d = kl + 1;
          erhfb[kl] = enb; drhfb[kl] = ekb; s.occ[it][kl] = pn;
          kl++;
          sumnz += 2.0 * pn + 2.0 * (1.0 - 2.0 * pn) * fT;
        }
      }
      if (!norm_to_improve) { s.ka[it][ib] = kaib; s.kd[it][ib] = kl - kaib; }
    }
    if (kl == 0) throw std::runtime_error("hfbdiag: no states below the pairing cut-off");
    if (iparenti != 0 && ibiblo == 0) throw std::runtime_error("hfbdiag: no blocking candidate found");
    sumnz -= tz;
    s.klmax[it] = kl;
    if (!norm_to_improve) s.ala[it] = al;
    double alnew = al;
    if (s.ft_active) s.fT_pwi[it].assign(kl, 0.0);
    alambda(alnew, kl, erhfb, drhfb, tz, s.CpV0[it], blok1k2d, s.temper, s.ft_active ? &s.fT_pwi[it] : nullptr);
    if (keyblo == 0) ala = alnew;
    else ala = ala + 0.50 * (alnew - ala);
  }
  s.inner[it] = inner;
  s.ala_out[it] = ala;
  s.keyblo[it] = keyblo;
  s.blo_block[it] = keyblo ? bloblo_1 : 0;
  s.blo_state[it] = keyblo ? blo123_1 : 0;
  s.blok1k2d[it] = blok1k2d;
  (void)offuv;
}

// ---------------------------------------------------------------------------------------------
// DENSIT restricted to rho(r) (all the FAM set-up needs, pnfam_interaction.f90:244-257)
// ---------------------------------------------------------------------------------------------
static void densit_rho(HfbSolution& s) {
  const int nghl = s.nghl;
  for (int it = 0; it < 2; it++) {
    std::vector<double> ro(nghl, 0.0);
    for (int ib = 0; ib < s.nb; ib++) {
      const int nd = s.id[ib], im = s.ia[ib];
      const int k1 = s.ka[it][ib], imen = s.kd[it][ib];
      if (imen <= 0) continue;
      const int k0 = (s.keyblo[it] && s.blo_block[it] == ib + 1) ? s.blo_state[it] : 0;
      // The reference indexes the ACTIVE list of the block with blo123d (DENSIT: "PNIK=OMPANK(JN+K0)",
      // "If(K.Ne.K0) Cycle"), i.e. it assumes every qp below the blocked one is inside the window.
      const int k0_active = (k0 >= 1 && k0 <= imen) ? k0 : 0;
      std::vector<double> tfiu(imen), tfid(imen), pfiu, pfid;
      if (s.ft_active) { pfiu.resize(imen); pfid.resize(imen); }
      for (int ihil = 0; ihil < nghl; ihil++) {
        std::fill(tfiu.begin(), tfiu.end(), 0.0);
        std::fill(tfid.begin(), tfid.end(), 0.0);
        std::fill(pfiu.begin(), pfiu.end(), 0.0);
        std::fill(pfid.begin(), pfid.end(), 0.0);
        double piu = 0, pid = 0;
        for (int i = 0; i < nd; i++) {
          const int ja = im + i;
          const double q = s.qhla[(size_t)ja * nghl + ihil];
          double* dst = s.ns[ja] > 0 ? tfiu.data() : tfid.data();
          for (int k = 0; k < imen; k++) dst[k] += q * s.V[it][(size_t)s.Kpwi[it][k1 + k] + i];
          if (s.ft_active) {                          // the U component of the quasiparticle (hfbtho_solver.f90:4484-4492)
            double* dsu = s.ns[ja] > 0 ? pfiu.data() : pfid.data();
            for (int k = 0; k < imen; k++) dsu[k] -= q * s.U[it][(size_t)s.Kpwi[it][k1 + k] + i];
          }
          if (k0_active) {
            const double pnik = s.U[it][(size_t)s.Kpwi[it][k1 + k0_active - 1] + i];
            if (s.ns[ja] > 0) piu += pnik * q; else pid += pnik * q;
          }
        }
        double t = 0;
        for (int k = 0; k < imen; k++) {
          double temp2 = tfiu[k] * tfiu[k] + tfid[k] * tfid[k];
          if (s.ft_active) {                          // TEMP2 of DENSIT, hfbtho_solver.f90:4535-4545
            const double f1k = s.fT_pwi[it][k1 + k], fk = 1.0 - f1k;
            temp2 = temp2 * fk + (pfiu[k] * pfiu[k] + pfid[k] * pfid[k]) * f1k;
          }
          t += temp2;
          if (k + 1 == k0_active) t -= 0.5 * (temp2 - (piu * piu + pid * pid));
        }
        ro[ihil] += t;
      }
    }
    double ssum = 0;
    for (int i = 0; i < nghl; i++) ssum += ro[i];
    const double sN = 2.0 * ssum;
    const double piu = 2.0 * (double)s.npr[it] / sN;
    s.ro[it].resize(nghl);
    for (int i = 0; i < nghl; i++) s.ro[it][i] = ro[i] * (s.wdcori[i] * piu);
  }
}

// DENSIT restricted to tau(r) and Delta rho(r): TEMP4 / TEMP5 (hfbtho_solver.f90:4535-4566) -- zero temperature, the
// finite-temperature branch (V part weighted with 1 - f_k, U part with f_k) and the equal-filling correction of the
// blocked quasiparticle (:4586-4596) -- with the weights and the particle-number rescaling of the end of the routine
// (:4697-4716).
void HfbSolution::kinetic_and_laplacian(std::vector<double> tau[2], std::vector<double> dro[2]) const {
  const HfbSolution& s = *this;
  const int nghl = s.nghl;
  for (int it = 0; it < 2; it++) {
    std::vector<double> ro(nghl, 0.0), ta(nghl, 0.0), dr(nghl, 0.0);
    for (int ib = 0; ib < s.nb; ib++) {
      const int nd = s.id[ib], im = s.ia[ib];
      const int k1 = s.ka[it][ib], imen = s.kd[it][ib];
      if (imen <= 0) continue;
      const int k0 = (s.keyblo[it] && s.blo_block[it] == ib + 1) ? s.blo_state[it] : 0;
      const int k0_active = (k0 >= 1 && k0 <= imen) ? k0 : 0;      // as in densit_rho above
      // LAPLUS = Omega + 1/2 of the block; Lambda of the spin-down (xlap) and spin-up (xlam) components
      const double xlap = (double)(s.nl[im] + (s.ns[im] + 1) / 2), xlam = xlap - 1.0;
      const double xlap2 = xlap * xlap, xlam2 = xlam * xlam;
      const bool hot = s.ft_active;
#pragma omp parallel
      {
        // V part: [0] phi, [1] d/dr, [2] d/dz, [3] second derivatives; u = spin up, d = spin down; p*: the U part (T > 0)
        std::vector<double> fu[4], fd[4], pu[4], pd[4];
        for (int c = 0; c < 4; c++) { fu[c].resize(imen); fd[c].resize(imen); if (hot) { pu[c].resize(imen); pd[c].resize(imen); } }
#pragma omp for schedule(static)
        for (int ihil = 0; ihil < nghl; ihil++) {
          for (int c = 0; c < 4; c++) {
            std::fill(fu[c].begin(), fu[c].end(), 0.0); std::fill(fd[c].begin(), fd[c].end(), 0.0);
            if (hot) { std::fill(pu[c].begin(), pu[c].end(), 0.0); std::fill(pd[c].begin(), pd[c].end(), 0.0); }
          }
          double bu[4] = {0, 0, 0, 0}, bd[4] = {0, 0, 0, 0};        // blocked level: PIU, PIUR, PIUZ, PIUD2 / PID...
          for (int i = 0; i < nd; i++) {
            const int ja = im + i;
            const double w[4] = {s.qhla[(size_t)ja * nghl + ihil], s.fi1r[(size_t)ja * nghl + ihil],
                                 s.fi1z[(size_t)ja * nghl + ihil], s.fi2d[(size_t)ja * nghl + ihil]};
            const bool up = s.ns[ja] > 0;
            for (int k = 0; k < imen; k++) {
              const double v = s.V[it][(size_t)s.Kpwi[it][k1 + k] + i];
              for (int c = 0; c < 4; c++) (up ? fu : fd)[c][k] += w[c] * v;
              if (hot) {
                const double u = s.U[it][(size_t)s.Kpwi[it][k1 + k] + i];
                for (int c = 0; c < 4; c++) (up ? pu : pd)[c][k] -= w[c] * u;
              }
            }
            if (k0_active) {
              const double pnik = s.U[it][(size_t)s.Kpwi[it][k1 + k0_active - 1] + i];
              for (int c = 0; c < 4; c++) (up ? bu : bd)[c] += pnik * w[c];
            }
          }
          const double y = s.y[ihil], y2 = y * y;
          double t2 = 0, t4 = 0, t5 = 0;
          for (int k = 0; k < imen; k++) {
            const double tw = fu[1][k] * fu[1][k] + fd[1][k] * fd[1][k] + fu[2][k] * fu[2][k] + fd[2][k] * fd[2][k];
            double e2 = fu[0][k] * fu[0][k] + fd[0][k] * fd[0][k];
            double e4 = xlam2 * y2 * fu[0][k] * fu[0][k] + xlap2 * y2 * fd[0][k] * fd[0][k] + tw;
            double e5 = fu[0][k] * fu[3][k] + fd[0][k] * fd[3][k] + tw;
            if (hot) {
              const double f1k = s.fT_pwi[it][k1 + k], fk = 1.0 - f1k;
              const double twp = pu[1][k] * pu[1][k] + pd[1][k] * pd[1][k] + pu[2][k] * pu[2][k] + pd[2][k] * pd[2][k];
              const double tw_t = tw * fk + twp * f1k;
              e2 = e2 * fk + (pu[0][k] * pu[0][k] + pd[0][k] * pd[0][k]) * f1k;
              e4 = (xlam2 * y2 * fu[0][k] * fu[0][k] + xlap2 * y2 * fd[0][k] * fd[0][k]) * fk +
                   (xlam2 * y2 * pu[0][k] * pu[0][k] + xlap2 * y2 * pd[0][k] * pd[0][k]) * f1k + tw_t;
              e5 = (fu[0][k] * fu[3][k] + fd[0][k] * fd[3][k]) * fk + (pu[0][k] * pu[3][k] + pd[0][k] * pd[3][k]) * f1k + tw_t;
            }
            t2 += e2; t4 += e4; t5 += e5;
            if (k + 1 == k0_active) {
              const double pw = bu[1] * bu[1] + bd[1] * bd[1] + bu[2] * bu[2] + bd[2] * bd[2];
              t2 -= 0.5 * (e2 - (bu[0] * bu[0] + bd[0] * bd[0]));
              t4 -= 0.5 * (e4 - (pw + xlam2 * y2 * bu[0] * bu[0] + xlap2 * y2 * bd[0] * bd[0]));
              t5 -= 0.5 * (e5 - (pw + bu[0] * bu[3] + bd[0] * bd[3]));
            }
          }
          ro[ihil] += t2; ta[ihil] += t4; dr[ihil] += t5;
        }
      }
    }
    double ssum = 0;
    for (int i = 0; i < nghl; i++) ssum += ro[i];
    const double piu = 2.0 * (double)s.npr[it] / (2.0 * ssum);
    tau[it].resize(nghl); dro[it].resize(nghl);
    for (int i = 0; i < nghl; i++) {
      tau[it][i] = ta[i] * (s.wdcori[i] * piu);
      dro[it][i] = dr[i] * (s.wdcori[i] * piu * 2.0);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Cache of the expensive stages.  Not part of the reference (which repeats the whole zero-iteration HFBTHO run in every
// pnfam_main.x launch, 4-18 s at 16 shells); SURVEY.md section 8f row 4.
unsigned long long hash_file(const std::string& path, unsigned long long seed) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return 0;
  unsigned long long h = seed ? seed : 1469598103934665603ull;
  unsigned char buf[1 << 16];
  size_t n;
  while ((n = std::fread(buf, 1, sizeof buf, f)) > 0)
    for (size_t i = 0; i < n; i++) { h ^= buf[i]; h *= 1099511628211ull; }
  std::fclose(f);
  return h ? h : 1;
}

namespace {
constexpr unsigned long long CACHE_MAGIC = 0x42323030484642ull + (3ull << 56);   // "B200HFB", format 3
struct CacheIO {
  FILE* f;
  bool ok = true, writing;
  CacheIO(FILE* f_, bool w) : f(f_), writing(w) {}
  void raw(void* p, size_t n) {
    if (!ok || n == 0) return;
    ok = (writing ? std::fwrite(p, 1, n, f) : std::fread(p, 1, n, f)) == n;
  }
  template <class T> void pod(T& x) { raw(&x, sizeof(T)); }
  template <class T> void vec(std::vector<T>& v) {
    unsigned long long n = v.size();
    pod(n);
    if (!ok) return;
    if (!writing) {
      if (n > (1ull << 32)) { ok = false; return; }
      v.resize((size_t)n);
    }
    raw(v.data(), (size_t)n * sizeof(T));
  }
};
// the fields written by gamdel / hfbdiag / densit_rho
void cache_fields(CacheIO& io, HfbSolution& s) {
  for (int it = 0; it < 2; it++) {
    io.vec(s.hmat[it]); io.vec(s.dmat[it]); io.vec(s.E[it]); io.vec(s.U[it]); io.vec(s.V[it]);
    io.vec(s.ka[it]); io.vec(s.kd[it]); io.vec(s.Kqp[it]); io.vec(s.Kpwi[it]); io.vec(s.occ[it]); io.vec(s.ro[it]);
    io.pod(s.ala[it]); io.pod(s.ala_out[it]); io.pod(s.inner[it]); io.pod(s.klmax[it]);
    io.pod(s.keyblo[it]); io.pod(s.blo_block[it]); io.pod(s.blo_state[it]); io.pod(s.blok1k2d[it]);
    io.vec(s.fT_pwi[it]);
  }
}
bool cache_load(const std::string& path, unsigned long long key, HfbSolution& s) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return false;
  CacheIO io(f, false);
  unsigned long long magic = 0, k = 0, tail = 0;
  int nt = 0, nghl = 0;
  io.pod(magic); io.pod(k); io.pod(nt); io.pod(nghl);
  bool good = io.ok && magic == CACHE_MAGIC && k == key && nt == s.nt && nghl == s.nghl;
  if (good) {
    cache_fields(io, s);
    io.pod(tail);
    good = io.ok && tail == (key ^ CACHE_MAGIC) && (int)s.E[0].size() == s.nt && (int)s.ro[0].size() == s.nghl;
    if (!good)                                      // truncated / foreign file: forget what was read, recompute
      for (int it = 0; it < 2; it++) {
        s.hmat[it].clear(); s.dmat[it].clear(); s.E[it].clear(); s.U[it].clear(); s.V[it].clear();
        s.ka[it].clear(); s.kd[it].clear(); s.Kqp[it].clear(); s.Kpwi[it].clear(); s.occ[it].clear(); s.ro[it].clear();
        s.ala[it] = s.ala_out[it] = 0; s.inner[it] = s.klmax[it] = 0;
        s.keyblo[it] = s.blo_block[it] = s.blo_state[it] = s.blok1k2d[it] = 0;
      }
  }
  std::fclose(f);
  return good;
}
void cache_save(const std::string& path, unsigned long long key, HfbSolution& s) {
  const std::string tmp = path + ".tmp" + std::to_string((long)getpid());
  FILE* f = std::fopen(tmp.c_str(), "wb");
  if (!f) return;                                   // read-only directory: no cache, no error
  CacheIO io(f, true);
  unsigned long long magic = CACHE_MAGIC, k = key, tail = key ^ CACHE_MAGIC;
  io.pod(magic); io.pod(k); io.pod(s.nt); io.pod(s.nghl);
  cache_fields(io, s);
  io.pod(tail);
  const bool ok = io.ok;
  std::fclose(f);
  if (!ok || std::rename(tmp.c_str(), path.c_str()) != 0) std::remove(tmp.c_str());
}
}  // namespace

HfbSolution HfbSolution::build(const HfbInput& in, const HelData& h, const std::string& cache_file, unsigned long long cache_key) {
  if (in.type_of_calculation < 0) throw std::runtime_error("Lipkin-Nogami (type_of_calculation<0) is not supported");
  if (h.finite_range) throw std::runtime_error("finite-range (Gogny) functionals are not supported");
  if (h.has_hfb_matrix) throw std::runtime_error(".hel with an HFBmatrX record is not supported");
  HfbSolution s;
  // finite temperature: switched off below 1e-10 MeV (hfbtho_interface.f90:157-158)
  s.ft_active = in.set_temperature && std::fabs(in.temperature) > 1e-10;
  s.temper = in.temperature;
  if (s.ft_active && in.temperature < 0) throw std::runtime_error("temperature out-of-bounds: T>=0");
  s.nb = h.nb; s.nt = h.nt; s.ngh = h.ngh; s.ngl = h.ngl; s.nghl = h.ngh * h.ngl; s.n_shells = h.n00;
  s.b0 = h.b0; s.bz = h.bz; s.bp = h.bp;
  s.id = h.id; s.nr = h.nr; s.nz = h.nz; s.nl = h.nl; s.ns = h.ns;
  s.ia.resize(s.nb);
  { int a = 0; for (int ib = 0; ib < s.nb; ib++) { s.ia[ib] = a; a += s.id[ib]; } }
  s.npar.resize(s.nt);
  for (int i = 0; i < s.nt; i++) s.npar[i] = 1 + ((s.nz[i] + s.nl[i]) % 2);
  if (std::abs(in.n_shells) != h.n00) throw std::runtime_error("number_of_shells differs between namelist and .hel");
  // particle numbers (hfbtho_solver.f90:330-352)
  int npr[2] = {in.neutron_number, in.proton_number};
  int iparenti[2] = {0, 0};
  for (int it = 0; it < 2; it++) {
    const int* nk = it == 0 ? in.neutron_blocking : in.proton_blocking;
    if (nk[0] != 0) {
      if (nk[1] == 0) throw std::runtime_error("automatic blocking-candidate search is not supported; request a specific level");
      if (npr[it] % 2 == 0) {
        if (nk[0] > 0) { npr[it] += 1; iparenti[it] = -1; }
        else { npr[it] -= 1; iparenti[it] = +1; }
      } else {
        iparenti[it] = 999;
      }
    }
  }
  s.npr[0] = npr[0]; s.npr[1] = npr[1]; s.npr[2] = npr[0] + npr[1];
  for (int it = 0; it < 2; it++) { s.CpV0[it] = h.CpV0[it]; s.CpV1[it] = h.CpV1[it]; }
  if (in.user_pairing) {
    s.CpV0[0] = in.vpair_n; s.CpV0[1] = in.vpair_p; s.CpV1[0] = s.CpV1[1] = in.pairing_feature;
  }
  s.rho_nm = h.rho_nm; s.hbzero = h.hbzero; s.use_j2terms = h.use_j2terms; s.pwi = h.pwi;
  s.hfb_cr0 = h.Crho[1]; s.hfb_crr = h.Cdrho[1]; s.hfb_cdrho = h.CrDr[1];
  s.hfb_ctau = h.Ctau[1]; s.hfb_ctj = h.CJ[1]; s.hfb_crdj = h.CrdJ[1];
  build_tables(s, h);
  auto t0 = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (getenv("PNFAM_B200_SETUP_TIMING")) {
      auto t1 = std::chrono::steady_clock::now();
      std::fprintf(stderr, "[setup]   %-20s %.3f s\n", what, std::chrono::duration<double>(t1 - t0).count());
      t0 = t1;
    }
  };
  lap("basis, tables");
  if (!cache_file.empty() && cache_key != 0 && cache_load(cache_file, cache_key, s)) {
    s.from_cache = true;
    lap("cache load");
    return s;
  }
  gamdel(s, h);
  lap("gamdel");
  // the blocking request is by |2*Omega| with the sign telling particle/hole
  HfbInput in2 = in;
  in2.neutron_blocking[0] = std::abs(in.neutron_blocking[0]);
  in2.proton_blocking[0] = std::abs(in.proton_blocking[0]);
  for (int it = 0; it < 2; it++) hfbdiag(s, h, in2, it, iparenti[it], in.compatibility_hfodd);
  lap("hfbdiag");
  densit_rho(s);
  lap("densit");
  if (!cache_file.empty() && cache_key != 0) {
    cache_save(cache_file, cache_key, s);
    lap("cache save");
  }
  return s;
}

}  // namespace pnfam

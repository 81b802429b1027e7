// Shared by the drop-in executables (pnfam_main.x, contour_main.x): the .dat text contract of the reference
// (exes/pnfam/pnfam_txtoutput.f90:96-253, parsed by pynfam/outputs/pnfam_parser.py:33-132), the device context of a
// Problem and the batched solve through the C ABI.
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <ctime>
#include <string>
#include <vector>

#include "../../../include/pnfam_b200.h"
#include "../host/problem.hpp"

using namespace pnfam;

namespace pnfam_driver {

struct Log {
  FILE* f = nullptr;
  bool to_stdout = true;
  void line(const std::string& s, bool comment = true) {
    std::string t = s;
    while (!t.empty() && t.back() == ' ') t.pop_back();
    const char* pre = comment ? "#" : " ";
    if (f) std::fprintf(f, "%s%s\n", pre, t.c_str());
    if (to_stdout) { std::fprintf(stdout, "%s%s\n", pre, t.c_str()); std::fflush(stdout); }
  }
  void fmt(const char* format, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, format);
    std::vsnprintf(buf, sizeof buf, format, ap);
    va_end(ap);
    line(buf);
  }
};

inline std::string center(const std::string& s, int w) {
  int l = (int)s.size();
  int lpad = (w - l) / 2 + ((w - l) % 2 + 2) % 2;
  if (w - l < 0) return s;
  return std::string(lpad, ' ') + s;
}
static const char* kElements[] = {"n", "H", "He", "Li", "Be", "B", "C", "N", "O", "F", "Ne", "Na", "Mg", "Al", "Si", "P", "S", "Cl", "Ar", "K",
    "Ca", "Sc", "Ti", "V", "Cr", "Mn", "Fe", "Co", "Ni", "Cu", "Zn", "Ga", "Ge", "As", "Se", "Br", "Kr", "Rb", "Sr", "Y", "Zr", "Nb",
    "Mo", "Tc", "Ru", "Rh", "Pd", "Ag", "Cd", "In", "Sn", "Sb", "Te", "I", "Xe", "Cs", "Ba", "La", "Ce", "Pr", "Nd", "Pm", "Sm", "Eu",
    "Gd", "Tb", "Dy", "Ho", "Er", "Tm", "Yb", "Lu", "Hf", "Ta", "W", "Re", "Os", "Ir", "Pt", "Au", "Hg", "Tl", "Pb", "Bi", "Po", "At",
    "Rn", "Fr", "Ra", "Ac", "Th", "Pa", "U", "Np", "Pu", "Am", "Cm", "Bk", "Cf", "Es", "Fm", "Md", "No", "Lr", "Rf", "Db", "Sg", "Bh",
    "Hs", "Mt", "Ds", "Rg", "Cn", "Nh", "Fl", "Mc", "Lv", "Ts", "Og"};
inline std::string element(int z) { return (z >= 0 && z < (int)(sizeof kElements / sizeof *kElements)) ? kElements[z] : "??"; }

inline void write_dat(const Problem& p, const std::string& namelist, const std::string& base, double wre, double wim, const double* strength,
               int iters, int conv, double si, const double* trace, int max_iter, double total_seconds, bool to_stdout) {
  const FamInput& in = p.in;
  const FamBasis& b = p.nuc->basis;
  const Interaction& x = p.inter;
  Log log;
  log.to_stdout = to_stdout;
  if (!base.empty()) log.f = std::fopen((base + ".dat").c_str(), "w");
  const int w = 51;
  const std::string bar(w, '='), dash(w, '-');
  int zf = 0, nf = 0;
  if (in.beta_type == "-") { zf = b.npr[1] + 1; nf = b.npr[0] - 1; } else { zf = b.npr[1] - 1; nf = b.npr[0] + 1; }
  log.line("");
  log.line(" " + bar);
  log.line(center("pnFAM", w));
  log.line(center("Charge-changing Finite Amplitude Method", w));
  log.line(center("Version: 2.00-b200", w));
  log.line(center("B200-native solver (CUDA sm_100a, FP64 tensor cores)", w));
  log.line("");
  {
    std::time_t t = std::time(nullptr);
    std::tm tm = *std::localtime(&t);
    char buf[128];
    std::snprintf(buf, sizeof buf, "Run date: %d/%d/%d %d:%02d:%02d", tm.tm_mon + 1, tm.tm_mday, tm.tm_year + 1900, tm.tm_hour, tm.tm_min, tm.tm_sec);
    log.line(center(buf, w));
  }
  log.line(" " + bar);
  {
    char buf[256];
    std::snprintf(buf, sizeof buf, "Operator: %s%s for K = %d", in.operator_name.c_str(), in.beta_type.c_str(), in.operator_k);
    log.line(center(buf, w));
    std::snprintf(buf, sizeof buf, "Energy: (%8.4f%s%8.4fi)", wre, wim < 0 ? " - " : " + ", std::fabs(wim));
    log.line(center(buf, w));
    std::snprintf(buf, sizeof buf, "Parent nucleus:   %d%s(N=%d, Z=%d)", b.npr[2], element(b.npr[1]).c_str(), b.npr[0], b.npr[1]);
    log.line(center(buf, w));
    std::snprintf(buf, sizeof buf, "Daughter nucleus: %d%s(N=%d, Z=%d)", b.npr[2], element(zf).c_str(), nf, zf);
    log.line(center(buf, w));
  }
  log.line(" " + bar);
  log.line("");
  log.line(" " + dash);
  log.line(center(" pnFAM solver details", w));
  log.line(" " + dash);
  log.fmt(" FAM input parameter file name:   '%s'", namelist.c_str());
  log.fmt(" FAM output file name base:       '%s'", in.fam_output_filename.c_str());
  log.line("");
  log.fmt(" Operator:                        %s%s  K= %d", in.operator_name.c_str(), in.beta_type.c_str(), in.operator_k);
  log.fmt(" Two-body current mode (use_p):   %d (%d)", in.two_body_current_mode, in.two_body_current_usep ? 1 : 0);
  log.fmt(" Compute cross-terms:             %s", in.compute_crossterms ? "Yes" : "No");
  if (in.compute_crossterms) {
    std::string l = "None";
    if (!p.g.empty()) { l.clear(); for (size_t i = 0; i < p.g.size(); i++) l += (i ? ", " : "") + p.g[i].label; }
    log.fmt("   Cross-terms computed:          %s", l.c_str());
  }
  log.fmt(" Re(EQRPA):                       %8.4f MeV", wre);
  log.fmt(" Im(EQRPA):                       %8.4f MeV", wim);
  log.line("");
  log.fmt(" Number shells:                   %d", b.n_shells);
  log.fmt(" Basis size:                      %d", b.dqp);
  log.fmt(" Number matrix blocks:            %d", b.nb);
  log.fmt(" Non-trivial HFB matrix elements: %zu", b.dmat);
  log.fmt(" Non-trivial FAM matrix elements: %zu", p.f.mat.elem.size());
  log.line("");
  log.fmt(" Maximum iterations:              %d", in.max_iter);
  log.fmt(" Broyden history size:            %d", x.skip_residual ? -1 : in.broyden_history_size);
  log.fmt(" Broyden mixing factor:           %4.2f", (double)0.7f);
  log.fmt(" Convergence limit:               %8.1E", in.convergence_epsilon);
  log.line("");
  for (const auto& n : p.notes) log.line(n);
  if (!p.notes.empty()) log.line("");
  log.line(" " + dash);
  log.line(center(" pnFAM Interaction", w));
  log.line(" " + dash);
  if (x.skip_residual) {
    log.line(" No residual interaction.");
  } else {
    for (const auto& n : x.notes) log.line(n);
    log.line("");
    log.fmt(" %s functional", x.name.c_str());
    log.line(" " + dash);
    log.fmt(" Crho[0] = %15.9f;   Cs[0]   = %15.9f", x.cr0, x.cs0);
    log.fmt(" Crho[r] = %15.9f;   Cs[r]   = %15.9f", x.crr, x.csr);
    log.fmt(" (sigma) = %15.9f;   (sigma) = %15.9f", x.sigma_r, x.sigma_s);
    log.fmt(" Cdr     = %15.9f;   Cds     = %15.9f", x.cdrho, x.cds);
    log.fmt(" Ctau    = %15.9f;   Cj      = %15.9f", x.ctau, x.cj);
    log.fmt(" CrdJ    = %15.9f;   Csdj    = %15.9f", x.crdj, x.csdj);
    log.fmt(" CtJ0    = %15.9f;   Cgs     = %15.9f", x.ctj0, x.cgs);
    log.fmt(" CtJ1    = %15.9f;   CT      = %15.9f", x.ctj1, x.ct);
    log.fmt(" CtJ2    = %15.9f;   CF      = %15.9f", x.ctj2, x.cf);
    log.line("");
    log.fmt(" Cpr[0]  = %15.9f;   Cps[0]  = %15.9f", x.cpair0, x.cspair0);
    log.fmt(" Cpr[r]  = %15.9f;   Cps[r]  = %15.9f;   (sigma) = %12.9f", x.cpairr, x.cspairr, x.sigma_pair);
  }
  log.line("");
  log.line(" " + dash);
  log.line(center(" Statistical pnFAM details", w));
  log.line(" " + dash);
  log.fmt(" Finite-temperature active: %s", b.ft_active ? "Yes" : "No");
  if (b.ft_active) {                       // pnfam_txtoutput.f90:179-186 (its first format ends before " MeV")
    auto f04 = [](double v) {              // Fortran F0.4: no leading zero
      char t[64]; std::snprintf(t, sizeof t, "%.4f", v);
      std::string o(t);
      if (o.rfind("0.", 0) == 0) o.erase(0, 1);
      else if (o.rfind("-0.", 0) == 0) o.erase(1, 1);
      return o;
    };
    log.fmt("   Temperature .............: %8.4f", b.ft_temp);
    const size_t ip = std::max_element(b.qp_fp.begin(), b.qp_fp.end()) - b.qp_fp.begin();
    const size_t in_ = std::max_element(b.qp_fn.begin(), b.qp_fn.end()) - b.qp_fn.begin();
    log.fmt("   Approx. max. f_p ........: %8.4f (Ep=%sMeV)", b.qp_fp[ip], f04(b.Ep[ip]).c_str());
    log.fmt("   Approx. max. f_n ........: %8.4f (En=%sMeV)", b.qp_fn[in_], f04(b.En[in_]).c_str());
  }
  log.fmt(" Odd-nucleus EFA active ..: %s", b.blo_active ? "Yes" : "No");
  if (b.blo_active) {
    const char* nm[2] = {"Neutron", "Proton"};
    for (int it = 0; it < 2; it++)
      if (b.blo_qp[it])
        log.fmt("   %s blocking: QP index %d (block = %d, state = %d), QP energy %.4f MeV", nm[it], b.blo_qp[it], b.blo_ib[it], b.blo_is[it],
                (it == 0 ? b.En : b.Ep)[b.blo_qp[it] - 1]);
  }
  log.line("");
  const std::string rule(72, '-');
  log.line(" " + rule);
  log.line("     i               si               Re(S)               Im(S)      Time");
  log.line(" " + rule);
  const char* lab = "N";
  for (int i = 0; i <= iters; i++) {
    const double* t = trace + (size_t)i * 4;
    if (i == 1) lab = x.skip_residual ? "N" : "L";
    if (i >= 2) lab = x.skip_residual ? "N" : (in.broyden_history_size == 0 ? "L" : "B");
    log.fmt(" %4d%s  %15.10f  %18.10f  %18.10f  %8.3f", i, lab, t[0], i == 0 ? 0.0 : t[1], i == 0 ? 0.0 : t[2], i == 0 ? 0.0 : t[3]);
  }
  log.line(" " + rule);
  if (conv) log.fmt("  *  FAM iteration converged after   %4d steps.  si = %12.5E", iters, si);
  else log.fmt("  *  FAM iteration interrupted after %4d steps.  si = %12.5E", iters, si);
  log.line(" " + rule);
  log.fmt(" Total CPU time = %12.3E minutes", total_seconds / 60.0);
  log.line("");
  log.line("");
  log.line(" Result (Energy [MeV], strength and cross-terms [MeV^-1]):");
  log.line(" " + std::string(78, '-'));
  {
    char buf[256];
    std::snprintf(buf, sizeof buf, " %10s%34s%34s", "", "Real", "Imag");
    log.line(buf, false);
    std::snprintf(buf, sizeof buf, " %-10s%34.19E%34.19E", "Energy", wre, wim);
    log.line(buf, false);
    std::snprintf(buf, sizeof buf, " %-10s%34.19E%34.19E", "Strength", strength[0], strength[1]);
    log.line(buf, false);
    for (size_t k = 0; k < p.g.size(); k++) {
      std::snprintf(buf, sizeof buf, " %-10s%34.19E%34.19E", p.g[k].label.c_str(), strength[2 * (k + 1)], strength[2 * (k + 1) + 1]);
      log.line(buf, false);
    }
    log.line("", false);
  }
  if (log.f) std::fclose(log.f);
  (void)max_iter;
}


// device context of a problem's nucleus + interaction (the arrays stay owned by the Problem)
inline pnfam_b200_ctx* make_context(const Problem& p, int device, char* err, int errlen) {
  const FamBasis& b = p.nuc->basis;
  const Interaction& x = p.inter;
  pnfam_b200_model m{};
  m.nb = b.nb; m.dqp = b.dqp; m.nghl = b.nghl; m.db = b.db.data(); m.num_spin_up = b.num_spin_up.data();
  m.wf = b.wf.data(); m.wfdr = b.wfdr.data(); m.wfdp = b.wfdp.data(); m.wfdz = b.wfdz.data(); m.wfd2_all = b.wfd2_all.data();
  m.wdcori = b.wdcori.data(); m.crho = x.crho.data(); m.cs = x.cs.data(); m.cpair = x.cpair.data(); m.cspair = x.cspair.data();
  m.cdrho = x.cdrho; m.ctau = x.ctau; m.ctj0 = x.ctj0; m.ctj1 = x.ctj1; m.ctj2 = x.ctj2; m.crdj = x.crdj; m.cds = x.cds;
  m.ct = x.ct; m.cj = x.cj; m.cgs = x.cgs; m.cf = x.cf; m.csdj = x.csdj;
  m.Ep = b.Ep.data(); m.En = b.En.data(); m.Up = b.Up.data(); m.Vp = b.Vp.data(); m.Un = b.Un.data(); m.Vn = b.Vn.data();
  m.qp_fp = b.statistical() ? b.qp_fp.data() : nullptr; m.qp_fn = b.statistical() ? b.qp_fn.data() : nullptr;
  m.ngh = b.ngh; m.ngl = b.ngl; m.sep_nzrows = b.sep_nzrows;
  m.sep_zrow = b.sep_zrow.data(); m.sep_z = b.sep_z.data(); m.sep_r = b.sep_r.data();
  pnfam_b200_ctx* ctx = nullptr;
  if (pnfam_b200_ctx_create(&m, device, &ctx, err, errlen) != 0) return nullptr;
  return ctx;
}

struct BatchResult {
  int nstr = 1, tstride = 0;
  std::vector<double> strength, si, trace;   // [P][nstr][2], [P], [P][tstride]
  std::vector<int32_t> iters, conv;
  pnfam_b200_stats stats{};
};

// all frequencies of the problem's operator in one batched call of the C ABI
inline int solve_batch(pnfam_b200_ctx* ctx, const Problem& p, const std::vector<double>& wre, const std::vector<double>& wim,
                       BatchResult& r, char* err, int errlen) {
  std::vector<const double*> gptr;
  for (const auto& g : p.g) gptr.push_back(g.mat.elem.data());
  pnfam_b200_operator op{};
  op.beta_minus = p.f.beta_minus ? 1 : 0; op.nxterms = (int)p.g.size(); op.f_ir2c = p.f.mat.ir2c.data();
  op.f_elem = p.f.mat.elem.data(); op.g_elem = gptr.data();
  pnfam_b200_solver_params prm{};
  prm.max_iter = p.in.max_iter; prm.broyden_history_size = p.in.broyden_history_size;
  prm.convergence_epsilon = p.in.convergence_epsilon;
  prm.quench_residual_int = p.inter.skip_residual ? 0.0 : p.in.quench_residual_int;
  prm.energy_shift_prot = p.in.energy_shift_prot; prm.energy_shift_neut = p.in.energy_shift_neut;
  const int P = (int)wre.size();
  r.nstr = 1 + op.nxterms; r.tstride = (prm.max_iter + 1) * 4;
  r.strength.assign((size_t)P * r.nstr * 2, 0.0); r.si.assign(P, 0.0); r.trace.assign((size_t)P * r.tstride, 0.0);
  r.iters.assign(P, 0); r.conv.assign(P, 0);
  return pnfam_b200_solve(ctx, &op, &prm, P, wre.data(), wim.data(), r.strength.data(), r.iters.data(), r.conv.data(), r.si.data(),
                          r.trace.data(), &r.stats, err, errlen);
}

}  // namespace pnfam_driver

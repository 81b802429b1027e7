// Host-side FAM set-up.  See fam_setup.hpp for the reference locations.
#include "fam_setup.hpp"

#include <algorithm>
#include <cctype>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <mutex>
#include <stdexcept>

namespace pnfam {

static std::string upper(std::string s) {
  for (auto& c : s) c = (char)std::toupper((unsigned char)c);
  return s;
}

// ---------------------------------------------------------------------------------------------
// pnFAM namelist
// ---------------------------------------------------------------------------------------------
FamInput FamInput::read(const std::string& path) {
  Namelist nl = Namelist::parse_file(path);
  FamInput in;
  in.namelist_path = path;
  in.fam_output_filename = nl.get_string("general", "fam_output_filename", "");
  in.print_stdout = nl.get_bool("general", "print_stdout", true);
  in.use_fam_storage = nl.get_int("general", "use_fam_storage", 0);
  in.real_eqrpa = nl.get_double("general", "real_eqrpa", 0.0);
  in.imag_eqrpa = nl.get_double("general", "imag_eqrpa", 0.5);
  in.beta_type = nl.get_string("ext_field", "beta_type", "-");
  in.operator_name = upper(nl.get_string("ext_field", "operator_name", "F"));
  in.operator_k = nl.get_int("ext_field", "operator_k", 0);
  in.compute_crossterms = nl.get_bool("ext_field", "compute_crossterms", false);
  in.two_body_current_mode = nl.get_int("ext_field", "two_body_current_mode", 0);
  in.two_body_current_usep = nl.get_bool("ext_field", "two_body_current_usep", false);
  auto lecs = nl.get_doubles("ext_field", "two_body_current_lecs", {-3.4, 5.4, 0.0});
  // the defaults are single-precision literals in the reference (pnfam_setup.f90:190-192)
  if (!nl.has("ext_field", "two_body_current_lecs")) { lecs[0] = (double)-3.4f; lecs[1] = (double)5.4f; }
  for (int i = 0; i < 3; i++) in.two_body_current_lecs[i] = lecs[i];
  in.max_iter = nl.get_int("solver", "max_iter", 200);
  in.broyden_history_size = nl.get_int("solver", "broyden_history_size", 50);
  in.convergence_epsilon = nl.get_double("solver", "convergence_epsilon", 1e-7);
  in.energy_shift_prot = nl.get_double("solver", "energy_shift_prot", 0.0);
  in.energy_shift_neut = nl.get_double("solver", "energy_shift_neut", 0.0);
  in.quench_residual_int = nl.get_double("solver", "quench_residual_int", 1.0);
  in.interaction_name = nl.get_string("interaction", "interaction_name", "NONE");
  in.require_self_consistency = nl.get_bool("interaction", "require_self_consistency", true);
  in.require_gauge_invariance = nl.get_bool("interaction", "require_gauge_invariance", false);
  in.force_j2_terms = nl.get_bool("interaction", "force_j2_terms", false);
  auto opt = [&](const char* k, bool& has, double& v) {
    has = nl.has("interaction", k);
    if (has) v = nl.get_double("interaction", k, 0.0);
  };
  opt("vpair0", in.has_vpair0, in.vpair0);
  opt("vpair1", in.has_vpair1, in.vpair1);
  opt("vpair_t0", in.has_vpair_t0, in.vpair_t0);
  opt("vpair_t1", in.has_vpair_t1, in.vpair_t1);
  const char* ov[8] = {"override_cs0", "override_csr", "override_cds", "override_ct",
                       "override_cf", "override_cgs", "override_cj", "override_csdj"};
  for (int i = 0; i < 8; i++) opt(ov[i], in.has_override[i], in.override_val[i]);
  return in;
}

// ---------------------------------------------------------------------------------------------
// Doubled, spin-sorted basis (hfbtho_solution.f90:137-395) + pairing-window erase (:402-459)
// ---------------------------------------------------------------------------------------------
FamBasis FamBasis::build(const HfbSolution& s) {
  FamBasis b;
  const int hb = s.nb, ht = s.nt, nghl = s.nghl;
  b.nb = 2 * hb; b.dqp = 2 * ht; b.nghl = nghl; b.n_shells = s.n_shells;
  for (int i = 0; i < 3; i++) b.npr[i] = s.npr[i];
  b.db.resize(b.nb);
  for (int i = 0; i < hb; i++) b.db[i] = b.db[i + hb] = s.id[i];
  b.isstart.resize(b.nb);
  { int a = 1; for (int i = 0; i < b.nb; i++) { b.isstart[i] = a; a += b.db[i]; } }
  size_t hmat = 0;
  for (int i = 0; i < hb; i++) hmat += (size_t)s.id[i] * s.id[i];
  b.dmat = 2 * hmat;
  const int N = b.dqp;
  std::vector<int> nr(N), nz(N), nl(N), ns(N), npar(N);
  for (int i = 0; i < ht; i++) {
    nr[i] = nr[i + ht] = s.nr[i]; nz[i] = nz[i + ht] = s.nz[i];
    nl[i] = s.nl[i]; nl[i + ht] = -s.nl[i];
    ns[i] = s.ns[i]; ns[i + ht] = -s.ns[i];
    npar[i] = npar[i + ht] = s.npar[i];
  }
  // spin sort inside each block: spin-up first (new_order, hfbtho_solution.f90:309-343)
  std::vector<int> order(N);
  b.num_spin_up.assign(b.nb, 0);
  {
    int im = 0;
    for (int ib = 0; ib < b.nb; ib++) {
      int q = 0;
      for (int ic = im; ic < im + b.db[ib]; ic++) if (ns[ic] > 0) order[im + q++] = ic;
      b.num_spin_up[ib] = q;
      for (int ic = im; ic < im + b.db[ib]; ic++) if (ns[ic] < 0) order[im + q++] = ic;
      im += b.db[ib];
    }
  }
  b.nr.resize(N); b.nz.resize(N); b.nl.resize(N); b.ns.resize(N); b.npar.resize(N);
  b.wf.resize((size_t)nghl * N); b.wfdr = b.wf; b.wfdz = b.wf; b.wfd2 = b.wf; b.wfdp = b.wf; b.wfd2_all = b.wf;
  b.y = s.y; b.z = s.z; b.wdcori = s.wdcori;
  for (int i = 0; i < N; i++) {
    const int src = order[i];
    b.nr[i] = nr[src]; b.nz[i] = nz[src]; b.nl[i] = nl[src]; b.ns[i] = ns[src]; b.npar[i] = npar[src];
    const int h = src < ht ? src : src - ht;
    // time-reversed partner: same spatial function, sign -1 if the ORIGINAL state had ns<0
    const double sgn = (src >= ht && s.ns[h] < 0) ? -1.0 : 1.0;
    const double* q = &s.qhla[(size_t)h * nghl];
    const double* f1r = &s.fi1r[(size_t)h * nghl];
    const double* f1z = &s.fi1z[(size_t)h * nghl];
    const double* f2d = &s.fi2d[(size_t)h * nghl];
    double* w = &b.wf[(size_t)i * nghl];
    double* wr = &b.wfdr[(size_t)i * nghl];
    double* wz = &b.wfdz[(size_t)i * nghl];
    double* w2 = &b.wfd2[(size_t)i * nghl];
    double* wp = &b.wfdp[(size_t)i * nghl];
    double* wa = &b.wfd2_all[(size_t)i * nghl];
    for (int r = 0; r < nghl; r++) {
      w[r] = sgn * q[r]; wr[r] = sgn * f1r[r]; wz[r] = sgn * f1z[r]; w2[r] = sgn * f2d[r];
      wp[r] = b.y[r] * b.nl[i] * w[r];
      wa[r] = w2[r] - b.y[r] * b.nl[i] * wp[r];
    }
  }
  // separable factors in the doubled, spin-sorted order
  b.ngh = s.ngh; b.ngl = s.ngl; b.sep_nzrows = s.sep_nzrows; b.sep_z = s.sep_z;
  b.sep_zrow = b.nz;
  b.sep_r.assign((size_t)4 * N * s.ngl, 0.0);
  for (int i = 0; i < N; i++) {
    const int src = order[i], h = src < ht ? src : src - ht;
    const double sgn = (src >= ht && s.ns[h] < 0) ? -1.0 : 1.0;
    for (int il = 0; il < s.ngl; il++) {
      const double yl = b.y[(size_t)il * s.ngh] * b.nl[i];   // Lambda / r
      const double r0 = sgn * s.sep_r[((size_t)0 * ht + h) * s.ngl + il];
      b.sep_r[((size_t)0 * N + i) * s.ngl + il] = r0;
      b.sep_r[((size_t)1 * N + i) * s.ngl + il] = sgn * s.sep_r[((size_t)1 * ht + h) * s.ngl + il];
      b.sep_r[((size_t)2 * N + i) * s.ngl + il] = yl * r0;
      b.sep_r[((size_t)3 * N + i) * s.ngl + il] = sgn * s.sep_r[((size_t)2 * ht + h) * s.ngl + il] - yl * yl * r0;
    }
  }
  // quasiparticles: E doubled; U doubled; V: first half = -V, second half = +V (:225-230)
  auto dbl_E = [&](const std::vector<double>& e) { std::vector<double> o(N); for (int i = 0; i < ht; i++) o[i] = o[i + ht] = e[i]; return o; };
  b.Ep = dbl_E(s.E[1]); b.En = dbl_E(s.E[0]);
  auto dbl_M = [&](const std::vector<double>& m, double s1, double s2) {
    std::vector<double> o(b.dmat);
    for (size_t i = 0; i < hmat; i++) { o[i] = s1 * m[i]; o[i + hmat] = s2 * m[i]; }
    return o;
  };
  b.Up = dbl_M(s.U[1], 1, 1); b.Un = dbl_M(s.U[0], 1, 1);
  b.Vp = dbl_M(s.V[1], -1, 1); b.Vn = dbl_M(s.V[0], -1, 1);
  // row permutation of U,V consistent with the spin sort (reorder_hfbmatrix_row :467-481)
  auto reorder_rows = [&](std::vector<double>& M) {
    std::vector<double> o(M.size());
    size_t im = 0; int ir = 0;
    for (int ib = 0; ib < b.nb; ib++) {
      const int d = b.db[ib];
      for (int ic = 0; ic < d; ic++) {
        for (int r = 0; r < d; r++) o[im + r] = M[im + (order[ir + r] - ir)];
        im += d;
      }
      ir += d;
    }
    M.swap(o);
  };
  reorder_rows(b.Un); reorder_rows(b.Vn); reorder_rows(b.Up); reorder_rows(b.Vp);
  // finite temperature: Fermi-Dirac occupations re-made from the doubled quasiparticle energies, indexed like E
  // (hfbtho_solution.f90:364-388); the pairing-window erase below zeroes them with E
  b.ft_active = s.ft_active; b.ft_temp = s.temper;
  if (b.ft_active) {
    b.qp_fn.resize(N); b.qp_fp.resize(N);
    for (int i = 0; i < N; i++) {
      b.qp_fn[i] = 0.5 * (1.0 - std::tanh(0.5 * b.En[i] / b.ft_temp));
      b.qp_fp[i] = 0.5 * (1.0 - std::tanh(0.5 * b.Ep[i] / b.ft_temp));
    }
  }
  // pairing window: zero E and the U,V columns of inactive quasiparticles (:402-459)
  for (int it = 0; it < 2; it++) {
    std::vector<char> active(N, 0);
    const int ncut = s.klmax[it];
    for (int k = 0; k < ncut; k++) {
      const int q = s.Kqp[it][k];  // 1-based among the ht qps
      if (q <= 0) throw std::runtime_error("inconsistent pairing-window bookkeeping");
      active[q - 1] = 1; active[q - 1 + ht] = 1;
    }
    std::vector<double>& E = it == 0 ? b.En : b.Ep;
    std::vector<double>& U = it == 0 ? b.Un : b.Up;
    std::vector<double>& V = it == 0 ? b.Vn : b.Vp;
    size_t im = 0; int iqp = 0;
    for (int ib = 0; ib < b.nb; ib++) {
      const int d = b.db[ib];
      for (int ic = 0; ic < d; ic++, iqp++, im += d) {
        if (!active[iqp]) {
          E[iqp] = 0;
          if (b.ft_active) (it == 0 ? b.qp_fn : b.qp_fp)[iqp] = 0;
          for (int r = 0; r < d; r++) { U[im + r] = 0; V[im + r] = 0; }
        }
      }
    }
  }
  b.rho_n = s.ro[0]; b.rho_p = s.ro[1];
  b.src = &s;
  // equal-filling blocking data (pnfam_setup.f90:323-362)
  for (int it = 0; it < 2; it++) {
    if (s.keyblo[it] != 0) {
      b.blo_active = true;
      b.blo_qp[it] = s.Kqp[it][s.blok1k2d[it] - 1];
      b.blo_ib[it] = s.blo_block[it]; b.blo_is[it] = s.blo_state[it];
    }
  }
  if (b.blo_active && b.ft_active)       // pnfam_setup.f90:166-169
    throw std::runtime_error("this code cannot handle T>0 and odd-Z/odd-N simultaneously.");
  if (b.blo_active) {
    b.qp_fn.assign(N, 0.0); b.qp_fp.assign(N, 0.0);
    auto trev = [&](int q) { return q <= N / 2 ? q + N / 2 : q - N / 2; };
    if (b.blo_qp[0]) { b.qp_fn[b.blo_qp[0] - 1] = 0.5; b.qp_fn[trev(b.blo_qp[0]) - 1] = 0.5; }
    if (b.blo_qp[1]) { b.qp_fp[b.blo_qp[1] - 1] = 0.5; b.qp_fp[trev(b.blo_qp[1]) - 1] = 0.5; }
  }
  for (int it = 0; it < 2; it++) { b.hfb_cpair[it] = s.CpV0[it]; b.hfb_alpha_pair[it] = s.CpV1[it]; }
  b.rho_nm = s.rho_nm; b.hbzero = s.hbzero;
  b.hfb_cr0 = s.hfb_cr0; b.hfb_crr = s.hfb_crr; b.hfb_cdrho = s.hfb_cdrho;
  b.hfb_ctau = s.hfb_ctau; b.hfb_ctj = s.hfb_ctj; b.hfb_crdj = s.hfb_crdj;
  b.hfb_use_j2terms = s.use_j2terms;
  return b;
}

// ---------------------------------------------------------------------------------------------
// Interaction (pnfam_interaction.f90:90-260, :267-715)
// ---------------------------------------------------------------------------------------------
namespace {
struct Skyrme {
  const char* key; const char* key2; const char* shown;
  double t0, t1, t2, t3, x0, x1, x2, x3, b4, b4p, alpha, tto, tte;
  bool j2;
};
// (t,x) tables of the built-in functionals, pnfam_interaction.f90:339-675.  w -> b4=b4p=w/2
// unless b4/b4p are given explicitly there.
const Skyrme kSkyrme[] = {
    {"SIII", "S3", "SIII", -1128.75, 395.0, -95.0, 14000.0, 0.45, 0.0, 0.0, 1.0, 60.0, 60.0, 1.0, 0, 0, false},
    {"SGII", "SG2", "SGII", -2645.0, 340.0, -41.9, 15595.0, 0.09, -0.0588, 1.425, 0.06044, 52.5, 52.5, 1.0 / 6.0, 0, 0, false},
    {"SKM*", "", "SkM*", -2645.0, 410.0, -135.0, 15595.0, 0.09, 0.0, 0.0, 0.0, 65.0, 65.0, 1.0 / 6.0, 0, 0, false},
    {"SKO", "", "SkO", -2103.653, 303.352, 791.674, 13553.252, -0.210701, -2.810752, -1.461595, -0.429881, 176.578, -198.7490, 0.25, 0, 0, false},
    {"SKOP", "SKO'", "SkO'", -2099.419, 301.531, 154.781, 13526.464, -0.029503, -1.325732, -2.323439, -0.147404, 143.895, -82.8888, 0.25, 0, 0, true},
    {"SLY4", "", "SLy4", -2488.913, 486.818, -546.395, 13777.0, 0.834, -0.344, -1.0, 1.354, 61.5, 61.5, 1.0 / 6.0, 0, 0, false},
    {"SLY4PUB", "", "SLy4 (published)", -2488.91, 486.820, -546.390, 13777.0, 0.834, -0.344, -1.0, 1.354, 61.5, 61.5, 1.0 / 6.0, 0, 0, false},
    {"SLY5", "", "SLy5 (HFBTHO)", -2483.45, 484.23, -556.69, 13757.0, 0.776, -0.317, -1.0, 1.263, 62.5, 62.5, 1.0 / 6.0, 0, 0, true},
    {"SLY5PUB", "", "SLy5 (published)", -2484.88, 483.13, -549.40, 13763.0, 0.778, -0.328, -1.0, 1.267, 63.0, 63.0, 1.0 / 6.0, 0, 0, true},
    {"T43", "", "T43", -2490.275, 494.608, -255.534, 13847.12, 0.698702, -0.781655, -0.646302, 1.135795, 153.103 / 2.0, 153.103 / 2.0, 1.0 / 6.0, -49.160, 196.868, true},
    {"SVMIN", "SV-MIN", "SV-min", -2112.248, 295.781, 142.268, 13988.567, 0.243886, -1.434926, -2.625899, 0.258070, 111.291 / 2.0, 45.93615, 0.255368, 0, 0, false},
    {"UNEDF0", "UNE0", "UNEDF0", -1883.68781034247695, 277.500212238931113, 608.430905591534383, 13901.9483446343256, 0.00974374651612197606, -1.77784394560870229, -1.67699034528797575, -0.38079041463310026, 125.161, -91.2604, 0.321955989588264435, 0, 0, false},
    {"UNEDF1", "UNE1", "UNEDF1", -2078.3280232556408, 239.400812041522045, 1575.11954189757989, 14263.646247077595, 0.0537569206858470316, -5.07723238187693671, -1.36650561394298609, -0.162491168089956339, 38.3680720616682009, 71.3165222295833985, 0.2700180115027076, 0, 0, false},
    {"UNEDF2", "UNE2", "UNEDF2", -1735.45685190850349, 262.814456235313969, 1183.31263243304011, 12293.2432941125371, 0.172247227624101162, -3.76692685046941023, -1.38349812677880979, 0.0528642727436126059, 25.658667730648304, 77.3003893702710059, 0.351455132555483607, -266.930662318713644, -240.140598215558754, true},
    {"UNEDF1HFB", "HFB1", "UNEDF1-HFB", -1666.40609652204898, 255.225795498529067, 1521.27603913447047, 12072.5070586946749, 0.146079699879160696, -4.82248258717499834, -1.35211044158474336, 0.00431668727902556615, 22.0338399999999979, 103.825096000000002, 0.37658031358391203, 0, 0, false},
};
}  // namespace

Interaction Interaction::build(const FamInput& in, const FamBasis& b, const std::string& rundir) {
  Interaction x;
  const int nghl = b.nghl;
  x.crho.assign(nghl, 0.0); x.cs = x.crho; x.cpair = x.crho; x.cspair = x.crho;
  std::string nm = upper(in.interaction_name);
  while (!nm.empty() && nm.front() == ' ') nm.erase(nm.begin());
  x.name = nm;
  if (nm == "NONE") { x.skip_residual = true; return x; }
  const bool from_file = nm.rfind("FILE:", 0) == 0;
  if (from_file) {
    // custom coupling constants (pnfam_interaction.f90:303-331): namelist &custom_interaction of the named file; none of
    // the built-in post-processing (J^2 terms, gauge invariance, pairing from vpair_*, overrides) applies (:126-129)
    std::string file = in.interaction_name;
    while (!file.empty() && file.front() == ' ') file.erase(file.begin());
    file = file.substr(5);
    while (!file.empty() && file.front() == ' ') file.erase(file.begin());
    while (!file.empty() && file.back() == ' ') file.pop_back();
    x.name = "FILE:" + file;
    const std::string path = (!file.empty() && file[0] == '/') ? file : rundir + "/" + file;
    {
      std::FILE* probe = std::fopen(path.c_str(), "r");
      if (!probe) throw std::runtime_error("could not open interaction file \"" + file + "\".");
      std::fclose(probe);
    }
    Namelist nl = Namelist::parse_file(path);
    if (nl.groups.find("custom_interaction") == nl.groups.end())
      throw std::runtime_error("could not read interaction file \"" + file + "\".");
    auto g = [&](const char* k) { return nl.get_double("custom_interaction", k, 0.0); };
    x.cr0 = g("cr0"); x.crr = g("crr"); x.sigma_r = g("sigma_r"); x.cdrho = g("cdrho"); x.ctau = g("ctau");
    x.ctj0 = g("ctj0"); x.ctj1 = g("ctj1"); x.ctj2 = g("ctj2"); x.crdj = g("crdj");
    x.cs0 = g("cs0"); x.csr = g("csr"); x.sigma_s = g("sigma_s"); x.cds = g("cds"); x.ct = g("ct"); x.cj = g("cj");
    x.csdj = g("csdj"); x.cf = g("cf"); x.cgs = g("cgs");
    x.cpair0 = g("cpair0"); x.cpairr = g("cpairr"); x.cspair0 = g("cspair0"); x.cspairr = g("cspairr"); x.sigma_pair = g("sigma_pair");
  }
  const Skyrme* sk = nullptr;
  if (!from_file) {
  for (const auto& e : kSkyrme) {
    if (nm == e.key || (e.key2[0] && nm == e.key2)) sk = &e;
  }
  if (nm == "UNE1HFB" || nm == "UNEDF1-HFB" || nm == "UNE1-HFB") sk = &kSkyrme[14];
  if (!sk) throw std::runtime_error("interaction \"" + in.interaction_name + "\" not found");
  x.name = sk->shown;
  const double t0 = sk->t0, t1 = sk->t1, t2 = sk->t2, t3 = sk->t3, x0 = sk->x0, x1 = sk->x1, x2 = sk->x2, x3 = sk->x3;
  const double b4p = sk->b4p, tto = sk->tto, tte = sk->tte;
  // (t,x) -> isovector couplings, pnfam_interaction.f90:679-697
  x.cr0 = -1.0 / 8.0 * t0 * (2.0 * x0 + 1.0);
  x.crr = -1.0 / 48.0 * t3 * (2.0 * x3 + 1.0);
  x.cs0 = -1.0 / 8.0 * t0;
  x.csr = -1.0 / 48.0 * t3;
  x.sigma_r = sk->alpha; x.sigma_s = sk->alpha;
  x.cdrho = 3.0 / 32.0 * t1 * (x1 + 0.5) + 1.0 / 32.0 * t2 * (x2 + 0.5);
  x.ctau = -1.0 / 8.0 * t1 * (x1 + 0.5) + 1.0 / 8.0 * t2 * (x2 + 0.5);
  x.ctj0 = 1.0 / 48.0 * (t1 - t2 + 10.0 * tte - 10.0 * tto);
  x.ctj1 = 1.0 / 32.0 * (t1 - t2 - 5.0 * tte + 5.0 * tto);
  x.ctj2 = 1.0 / 16.0 * (t1 - t2 + tte - tto);
  x.crdj = -0.5 * b4p;
  x.cds = 1.0 / 64.0 * (3.0 * t1 + t2 - 6.0 * tte - 2.0 * tto);
  x.ct = -1.0 / 16.0 * (t1 - t2 - 2.0 * tte + 2.0 * tto);
  x.cj = 1.0 / 8.0 * t1 * (x1 + 0.5) - 1.0 / 8.0 * t2 * (x2 + 0.5);
  x.cgs = -3.0 / 32.0 * (3.0 * tte + tto);
  x.cf = -3.0 / 8.0 * (tte - tto);
  x.csdj = -0.5 * b4p;
  x.ctj0 = 2 / 3.0 * x.ctj1;
  x.ctj2 = 2 * x.ctj1;
  // built-in functional post-processing, :118-239
  if (!sk->j2 && !in.force_j2_terms) { x.ctj0 = 0; x.ctj1 = 0; x.ctj2 = 0; }
  if (in.require_gauge_invariance) {
    x.cf = 2 * x.ctj1 - x.ctj2;
    x.ct = -2 * x.ctj1 + 0.5 * x.cf;
  }
  double vpair_t0, vpair_t1;
  if (!in.has_vpair_t0) vpair_t0 = in.has_vpair0 ? 2 * in.vpair0 : 0.0;
  else {
    if (in.has_vpair0) throw std::runtime_error("cannot set both vpair0 and vpair_t0");
    vpair_t0 = in.vpair_t0;
  }
  if (!in.has_vpair_t1) vpair_t1 = in.has_vpair1 ? in.vpair1 : 0.5 * (b.hfb_cpair[0] + b.hfb_cpair[1]);
  else {
    if (in.has_vpair1) throw std::runtime_error("cannot set both vpair1 and vpair_t1");
    vpair_t1 = in.vpair_t1;
  }
  x.sigma_pair = 1.0;
  x.cpair0 = vpair_t1 / 8.0;
  x.cspair0 = vpair_t0 / 8.0;
  x.cpairr = x.cpair0 * (-1 * b.hfb_alpha_pair[0] / std::pow(b.rho_nm, x.sigma_pair));
  x.cspairr = x.cspair0 * (-1 * b.hfb_alpha_pair[0] / std::pow(b.rho_nm, x.sigma_pair));
  double* ovt[8] = {&x.cs0, &x.csr, &x.cds, &x.ct, &x.cf, &x.cgs, &x.cj, &x.csdj};
  const char* ovn[8] = {"Cs0", "Csr", "Cds", "CT", "CF", "Cgs", "Cj", "Csdj"};
  for (int i = 0; i < 8; i++)
    if (in.has_override[i]) {
      *ovt[i] = in.override_val[i];
      x.notes.push_back(std::string(" [!] Coupling constant ") + ovn[i] + " has been overridden manually");
    }
  }   // built-in functional
  // self-consistency check against the HFBTHO couplings (:827-941), tolerance 5e-12
  if (in.require_self_consistency) {
    const double tol = 5.0e-12;
    auto bad = [&](double a, double h) { return std::fabs(h - a) > tol; };
    bool fail = bad(x.cr0, b.hfb_cr0) || bad(x.crr, b.hfb_crr) || bad(x.cdrho, b.hfb_cdrho) || bad(x.ctau, b.hfb_ctau) ||
                bad(x.crdj, b.hfb_crdj);
    if (b.hfb_use_j2terms) fail = fail || bad(2 * x.ctj1, b.hfb_ctj) || bad(x.ctj2, b.hfb_ctj);
    else fail = fail || std::fabs(x.ctj1) > tol;
    fail = fail || std::fabs(0.5 * (b.hfb_cpair[0] + b.hfb_cpair[1]) - 8 * x.cpair0) > tol;
    if (std::fabs(x.cpair0) > 0.0)
      fail = fail || std::fabs(b.hfb_alpha_pair[0] + x.cpairr * std::pow(b.rho_nm, x.sigma_pair) / x.cpair0) > tol;
    if (std::fabs(x.cspair0) > 0.0)
      fail = fail || std::fabs(b.hfb_alpha_pair[0] + x.cspairr * std::pow(b.rho_nm, x.sigma_pair) / x.cspair0) > tol;
    if (fail) throw std::runtime_error("FAM couplings are not self-consistent with the HFBTHO functional");
  }
  // gauge invariance (check_gauge_invariance :777-822), tolerance 1e-10
  if (in.require_gauge_invariance) {
    const double tol = 1.0e-10;
    if (std::fabs(x.ctau + x.cj) > tol || std::fabs(x.crdj - x.csdj) > tol || std::fabs(3 * x.ctj0 + x.ct + 2 * x.cf) > tol ||
        std::fabs(4 * x.ctj1 + 2 * x.ct - x.cf) > tol || std::fabs(2 * x.ctj2 + 2 * x.ct + x.cf) > tol)
      throw std::runtime_error("gauge invariance was requested, but is not fulfilled.");
  }
  for (int r = 0; r < nghl; r++) {
    const double rho = b.rho_n[r] + b.rho_p[r];
    x.crho[r] = x.cr0 + x.crr * std::pow(rho, x.sigma_r);
    x.cs[r] = x.cs0 + x.csr * std::pow(rho, x.sigma_s);
    x.cpair[r] = x.cpair0 + x.cpairr * std::pow(rho, x.sigma_pair);
    x.cspair[r] = x.cspair0 + x.cspairr * std::pow(rho, x.sigma_pair);
  }
  return x;
}

// ---------------------------------------------------------------------------------------------
// External fields
// ---------------------------------------------------------------------------------------------
static void init_fam_mapping(const FamBasis& b, ExtField& op) {
  const int nb = b.nb;
  bool pty_blocks = true;
  {
    int ip = 0;
    for (int i = 0; i < nb && pty_blocks; i++) {
      for (int k = ip; k < ip + b.db[i]; k++) if (b.npar[k] != b.npar[ip]) { pty_blocks = false; break; }
      ip += b.db[i];
    }
  }
  std::vector<int> ib1, ib2;
  int ip1 = 0;
  for (int i1 = 0; i1 < nb; i1++) {
    int ip2 = 0;
    for (int i2 = 0; i2 < nb; i2++) {
      const bool kmatch = (2 * b.nl[ip1] + b.ns[ip1] - 2 * b.nl[ip2] - b.ns[ip2]) == 2 * op.k;
      const bool pmatch = !pty_blocks || ((b.npar[ip1] == b.npar[ip2]) == op.parity_even);
      if (kmatch && pmatch) { ib1.push_back(i1); ib2.push_back(i2); break; }
      ip2 += b.db[i2];
    }
    ip1 += b.db[i1];
  }
  if (ib1.empty()) throw std::runtime_error("The K and parity are outside the model space");
  size_t n = 0;
  for (size_t i = 0; i < ib1.size(); i++) n += (size_t)b.db[ib1[i]] * b.db[ib2[i]];
  op.mat.init(nb, n);
  for (size_t i = 0; i < ib1.size(); i++) { op.mat.ir2c[ib1[i]] = ib2[i] + 1; op.mat.ic2r[ib2[i]] = ib1[i] + 1; }
  int ip = 1;
  for (int i = 0; i < nb; i++)
    if (op.mat.ir2c[i] > 0) {
      op.mat.ir2m[i] = ip; op.mat.ic2m[op.mat.ir2c[i] - 1] = ip;
      ip += b.db[i] * b.db[op.mat.ir2c[i] - 1];
    }
}

// I(1,1,1,1;r) of Behrens & Buehring for a uniform charge distribution (pnfam_extfield.f90:1683-1719)
static std::vector<double> i1111(const FamBasis& b) {
  const int A = b.npr[2];
  const double R = 1.2 * std::pow((double)A, 1.0 / 3.0);
  std::vector<double> I(b.nghl);
  for (int i = 0; i < b.nghl; i++) {
    const double r = 1.0 / b.y[i];
    const double rad = std::pow(r * r + b.z[i] * b.z[i], 0.5);
    if (rad <= R) I[i] = 1.5 * (1.0 - 0.2 * rad * rad / (R * R));
    else I[i] = 1.5 * (R / rad - 0.2 * (R * R * R) / (rad * rad * rad));
  }
  return I;
}

// ---- closed-form two-body-current factors (nuclear matter + local density approximation) ----------------
namespace {
constexpr double TB_HBARC = 197.3269718, TB_FPI = 92.4, TB_MN = 939.0, TB_MPI = 138.04, TB_GA = 1.27;   // pnfam_constants.f90:44-67
const double TB_PI = 3.14159265358979323846264338327950288;
inline double tb_caux() { return (TB_HBARC * TB_HBARC * TB_HBARC) / (2.0 * TB_MN * TB_FPI * TB_FPI); }
// tbc_nmlda_I0: the P = p = 0 limit of the nuclear-matter integrals (pnfam_extfield.f90:976-990)
inline double tb_i0(double kf) {
  const double m = TB_MPI / TB_HBARC, kf2 = kf * kf, kf3 = kf * kf * kf;
  return 1.0 - 3.0 * m * m / kf2 + 3.0 * m * m * m / kf3 * std::atan(kf / m);
}
inline double kf_snm(double rho) { return std::pow(1.5 * TB_PI * TB_PI * rho, 1.0 / 3.0); }    // lda_kf_snm
inline double kf_asnm(double rho) { return std::pow(3.0 * TB_PI * TB_PI * rho, 1.0 / 3.0); }   // lda_kf_asnm
}  // namespace

// GT: rho_fac of ext_field_operator (pnfam_extfield.f90:157-182): contact term + tbc_nmlda_da1 at Q = 0 (:1063-1128)
std::vector<double> tbc_gt_rho_fac(const FamBasis& b, const TwoBody& tb) {
  const double caux = tb_caux(), c3 = tb.lecs[0], c4 = tb.lecs[1], cd = tb.lecs[2] * (-0.25);
  std::vector<double> rf(b.nghl);
  for (int r = 0; r < b.nghl; r++) rf[r] = (caux * 2.0 * cd) * (b.rho_n[r] + b.rho_p[r]);
  if (tb.u[2] == 2 || tb.u[2] == 3) {
    const bool snm = tb.u[2] == 2;
    const double caux0 = (TB_HBARC * TB_HBARC * TB_HBARC) / (TB_MN * TB_FPI * TB_FPI);
    double caux3 = -1.0 / 3.0 * c3;
    if (tb.use_p) caux3 = caux3 + 1.0 / 12.0;
    const double caux4 = 1.0 / 3.0 * (c4 + 0.25);
    std::vector<double> da1(b.nghl, 0.0);
    for (int it = 1; it <= 3; it++) {
      if (snm ? it != 3 : it == 3) continue;
      for (int r = 0; r < b.nghl; r++) {
        const double ri = it == 1 ? b.rho_n[r] : (it == 2 ? b.rho_p[r] : b.rho_p[r] + b.rho_n[r]);
        const double ki = it == 3 ? kf_snm(ri) : kf_asnm(ri);
        const double I1 = tb_i0(ki), I2 = I1;
        da1[r] = da1[r] + caux0 * ri * (caux4 * (3.0 * I2 - I1) + caux3 * I1);
      }
    }
    for (int r = 0; r < b.nghl; r++) rf[r] = rf[r] + da1[r];
  } else if (tb.u[2] == 4 || tb.u[2] == 5) {
    const std::vector<double> ex = tbc_dme_exc(b, tb);
    for (int r = 0; r < b.nghl; r++) rf[r] = rf[r] + ex[r];
  }
  return rf;
}

// forbidden_2bc_rsL (pnfam_extfield.f90:1146-1193): the RS* fields are weighted with (1 - this)
std::vector<double> tbc_rsl_correction(const FamBasis& b, const TwoBody& tb, bool snm) {
  const double ca = 2 * tb_caux(), c3 = tb.lecs[0], c4 = tb.lecs[1], cd = tb.lecs[2] * (-0.25);
  const double ch = 1.0 / 3.0 * (2 * c4 - c3 + 0.5);
  std::vector<double> out(b.nghl);
  for (int r = 0; r < b.nghl; r++) {
    if (snm) {
      const double rho = b.rho_n[r] + b.rho_p[r];
      out[r] = ca * rho * (cd + ch * tb_i0(kf_snm(rho)));
    } else {
      const double rn = b.rho_n[r], rp = b.rho_p[r];
      out[r] = ca * (rn + rp) * cd + ca * ch * rn * tb_i0(kf_asnm(rn)) + ca * ch * rp * tb_i0(kf_asnm(rp));
    }
  }
  return out;
}

// forbidden_2bc_P (pnfam_extfield.f90:1339-1368)
std::vector<double> tbc_p_correction(const FamBasis& b) {
  const double P = (double)(std::sqrt(1.2f) * 1.30465f / 2.0f);   // `sqrt(1.2)*1.30465 / 2.0`: default-real (single precision) arithmetic in the source
  const double m = TB_MPI / TB_HBARC;
  std::vector<double> rf(b.nghl);
  for (int r = 0; r < b.nghl; r++) {
    const double kf = kf_snm(b.rho_n[r] + b.rho_p[r]);
    const double lg = std::log((m * m + (P + kf) * (P + kf)) / (m * m + (P - kf) * (P - kf)));
    const double iout = 1.0 / (P * P * P) * (P * m * m * kf - 0.25 * m * m * (m * m + P * P + kf * kf) * lg);
    rf[r] = TB_GA * TB_GA * TB_HBARC * TB_MN / (4 * TB_PI * TB_PI * TB_FPI * TB_FPI) * iout;
  }
  return rf;
}

// forbidden_2bc_PS0 (pnfam_extfield.f90:1370-1393)
std::vector<double> tbc_ps0_correction(const FamBasis& b) {
  const double P = (double)(std::sqrt(1.2f) * 1.30465f / 2.0f);
  const double m = TB_MPI / TB_HBARC;
  std::vector<double> rf(b.nghl);
  for (int r = 0; r < b.nghl; r++) {
    const double kf = kf_snm(b.rho_n[r] + b.rho_p[r]);
    const double lg = std::log((m * m + (P + kf) * (P + kf)) / (m * m + (P - kf) * (P - kf)));
    const double t1 = (m * m + P * P) * (m * m + P * P) + kf * kf * (kf * kf - 2 * P * P + 2 * m * m);
    const double iout = 1.0 / (P * P * P) * (-P * kf * (kf * kf + m * m + P * P) + 0.25 * (t1 * lg));
    rf[r] = -TB_HBARC * TB_MN / (8 * TB_PI * TB_PI * TB_FPI * TB_FPI) * iout;
  }
  return rf;
}

void FamBasis::need_tau_d2rho() const {
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  if (!tau0.empty()) return;
  if (!src) throw std::runtime_error("density-matrix-expansion currents: no HFB solution attached to the basis");
  std::vector<double> tau[2], dro[2];
  src->kinetic_and_laplacian(tau, dro);
  d2rho0.resize(nghl);
  for (int i = 0; i < nghl; i++) d2rho0[i] = dro[0][i] + dro[1][i];
  std::vector<double> t(nghl);
  for (int i = 0; i < nghl; i++) t[i] = tau[0][i] + tau[1][i];
  tau0.swap(t);
}

// ---- density-matrix expansion (pnfam_extfield.f90:1195-1330, 1380-1545): the U integrals, with the reference's cut-offs
namespace {
struct DmeU {
  double m;
  DmeU() : m(TB_MPI / TB_HBARC) {}
  double lg(double kf) const { return std::log(1.0 + 4.0 * kf * kf / (m * m)); }
  double u00(double kf) const {
    if (kf < 0.001) return 1.0 / (m * m);
    return 1.5 / (kf * kf) * (1 - m / kf * std::atan(2 * kf / m) + m * m / (4.0 * kf * kf) * lg(kf));
  }
  double u02(double kf) const {
    if (kf < 0.00005) return 0.0;
    return 0.75 / (kf * kf) * (lg(kf) - 4.0 * kf * kf / (4.0 * kf * kf + m * m));
  }
  double u02_kf2(double kf) const {
    if (kf < 0.001) return 6.0 / (m * m * m * m);
    return 0.75 / (kf * kf * kf * kf) * (lg(kf) - 4.0 * kf * kf / (4.0 * kf * kf + m * m));
  }
  double u22(double kf) const {
    return -1.0 * kf / (3.0 * TB_PI * TB_PI) * (1.0 + 8.0 * kf * kf / (4.0 * kf * kf + m * m) - m * m / (4.0 * kf * kf) * lg(kf));
  }
  double u10(double kf) const {
    if (kf < 0.001) return 3.0 / (m * m);
    return 1.5 / (kf * kf) * (1.0 - m * m / (4.0 * kf * kf) * lg(kf));
  }
  double u12(double kf) const {
    if (kf < 0.0001) return 0.0;
    const double d = 4.0 * kf * kf + m * m;
    return 3.0 / d * ((4.0 * kf * kf - m * m) / d + (1.0 + m * m / (4.0 * kf * kf)) * lg(kf));
  }
  double u12_kf2(double kf) const {
    if (kf < 0.001) return 30.0 / (m * m * m * m) + (-224.0 / std::pow(m, 6)) * kf * kf;
    const double d = 4.0 * kf * kf + m * m;
    return 3.0 / (kf * kf * d) * ((4.0 * kf * kf - m * m) / d + (1.0 + m * m / (4.0 * kf * kf)) * lg(kf));
  }
  double u24(double kf) const {
    const double d = 4.0 * kf * kf + m * m;
    return kf * kf * (80.0 * std::pow(kf, 4) + 40.0 * kf * kf * m * m - 3.0 * std::pow(m, 4)) / (d * d * d) + 0.75 * lg(kf);
  }
  double u24_kf4(double kf) const {
    if (kf < 0.001) return 35.0 / (3.0 * std::pow(m, 4)) + (-112.0 / std::pow(m, 6)) * kf * kf;
    return 1.0 / (6.0 * kf * kf * kf * kf) * u24(kf);
  }
};
}  // namespace

// dme_exc (pnfam_extfield.f90:1271-1307)
std::vector<double> tbc_dme_exc(const FamBasis& b, const TwoBody& tb) {
  b.need_tau_d2rho();
  const DmeU U;
  const double m = U.m, c3 = tb.lecs[0], c4 = tb.lecs[1], cd = tb.lecs[2] * (-0.25), d1 = cd * 0.50, d2 = cd * 0.25;
  const double ca = 2 * tb_caux(), ch = 1.0 / 3.0 * (2 * c4 - c3 + 0.5), cc = -d1 + 2 * d2;
  std::vector<double> rf(b.nghl);
  for (int r = 0; r < b.nghl; r++) {
    const double rho = b.rho_n[r] + b.rho_p[r], kf = kf_snm(rho);
    const double rf1 = ca * (ch * (1 - m * m * U.u00(kf) - m * m * U.u02(kf) * 0.1) + cc) * rho;
    const double rf2 = ca * ch * m * m / 6 * U.u02_kf2(kf) * (0.25 * b.d2rho0[r] - b.tau0[r]);
    rf[r] = rf1 - rf2;
  }
  return rf;
}

// dme_vector (pnfam_extfield.f90:1621-1662)
std::vector<double> tbc_dme_vector(const FamBasis& b) {
  b.need_tau_d2rho();
  const DmeU U;
  std::vector<double> rf(b.nghl);
  for (int r = 0; r < b.nghl; r++) {
    const double rho = b.rho_n[r] + b.rho_p[r], kf = kf_snm(rho);
    rf[r] = TB_GA * TB_GA * TB_MN * TB_HBARC / (4.0 * TB_FPI * TB_FPI) *
            ((U.u10(kf) + 0.1 * U.u12(kf)) * rho + U.u22(kf) - U.u24(kf) * kf / (15.0 * TB_PI * TB_PI) +
             (U.u12_kf2(kf) / 6.0 - U.u24_kf4(kf)) * (0.25 * b.d2rho0[r] - b.tau0[r]));
  }
  return rf;
}

// dme_axial_2 (pnfam_extfield.f90:1573-1602)
std::vector<double> tbc_dme_axial(const FamBasis& b) {
  b.need_tau_d2rho();
  const DmeU U;
  std::vector<double> rf(b.nghl);
  for (int r = 0; r < b.nghl; r++) {
    const double rho = b.rho_n[r] + b.rho_p[r], kf = kf_snm(rho);
    rf[r] = TB_HBARC * TB_MN / (6.0 * TB_FPI * TB_FPI) *
            ((U.u10(kf) + 0.1 * U.u12(kf)) * rho + U.u12_kf2(kf) / 6.0 * (0.25 * b.d2rho0[r] - b.tau0[r]));
  }
  return rf;
}

ExtField make_external_field(const FamBasis& b, const std::string& beta_type, const std::string& label_in, int K,
                             const std::vector<double>* rho_fac, const TwoBody* tb) {
  ExtField op;
  op.label = upper(label_in);
  const std::string& L = op.label;
  if (beta_type == "-") op.beta_minus = true;
  else if (beta_type == "+") op.beta_minus = false;
  else throw std::runtime_error("Unknown operator beta_type=" + beta_type);
  if (L == "F" || L == "GT") op.parity_even = true;
  else if (L == "RS0" || L == "RS1" || L == "RS2" || L == "R" || L == "P" || L == "PS0" || L == "RS0I" || L == "RI" || L == "RS1I")
    op.parity_even = false;
  else throw std::runtime_error("Unknown operator \"" + L + "\" in extfield.");
  op.k = K;
  int kmax = 0;
  if (L == "F" || L == "RS0" || L == "PS0" || L == "RS0I") { op.rank = 0; kmax = 0; }
  else if (L == "RS2") { op.rank = 2; kmax = 2; }
  else { op.rank = 1; kmax = 1; }
  if (std::abs(K) > kmax) throw std::runtime_error("K out of range for operator \"" + L + "\" in extfield.");
  init_fam_mapping(b, op);

  const int nghl = b.nghl;
  std::vector<double> r(nghl), ifun;
  for (int i = 0; i < nghl; i++) r[i] = 1.0 / b.y[i];
  const bool useI = (L == "RS0I" || L == "RI" || L == "RS1I");
  if (useI) ifun = i1111(b);
  // two-body-current weight on wf_1 (pnfam_extfield.f90:189-197, 349-366, 383-557, 562-607): RS* fields carry
  // (1 - correction) when the 4th mode digit is >= 2 (3: asymmetric nuclear matter); P and PS0 carry 1 + correction
  // (1BC + 2BC) or the correction alone (2BC only): digit 1 = nuclear matter, 2 = density-matrix expansion.
  std::vector<double> w1;
  if (tb && tb->active()) {
    const bool rs = (L == "RS0" || L == "RS1" || L == "RS2" || L == "RS0I" || L == "RS1I");
    if (rs && tb->u[4] >= 2) {
      w1 = tbc_rsl_correction(b, *tb, tb->u[4] != 3);
      for (auto& x : w1) x = 1.0 - x;
    } else if ((L == "P" && tb->u[5] != 0) || (L == "PS0" && tb->u[6] != 0)) {
      const int dg = L == "P" ? tb->u[5] : tb->u[6];
      if (dg == 1 || dg == 2) {
        if (dg == 1) w1 = L == "P" ? tbc_p_correction(b) : tbc_ps0_correction(b);
        else w1 = L == "P" ? tbc_dme_vector(b) : tbc_dme_axial(b);
        if (tb->u[1] == 1) for (auto& x : w1) x = 1 + x;
      }
    }
  }
  const bool useW1 = !w1.empty();
  // weight functions applied to wf_1 * (...) * wf_2 products
  auto dot = [&](const double* a, const double* c) {
    double s = 0;
    if (useW1) for (int i = 0; i < nghl; i++) s += (a[i] * w1[i]) * c[i];
    else for (int i = 0; i < nghl; i++) s += a[i] * c[i];
    return s;
  };
  auto dotw = [&](const double* a, const std::vector<double>& w, const double* c) {
    double s = 0;
    if (useI && useW1) for (int i = 0; i < nghl; i++) s += ((ifun[i] * a[i]) * w1[i]) * (w[i] * c[i]);
    else if (useI) for (int i = 0; i < nghl; i++) s += (ifun[i] * a[i]) * (w[i] * c[i]);
    else if (useW1) for (int i = 0; i < nghl; i++) s += (a[i] * w1[i]) * (w[i] * c[i]);
    else for (int i = 0; i < nghl; i++) s += a[i] * (w[i] * c[i]);
    return s;
  };
  const double sq2 = std::sqrt(2.0), sq3 = std::sqrt(3.0), sq32 = std::sqrt(3.0 / 2.0);
  // base label without the trailing I (same selection rules, I-function folded into dotw)
  std::string B = L;
  if (useI) B = L.substr(0, L.size() - 1);
  // every matrix element is an independent grid sum (same order of additions whatever the thread count): block rows in
  // parallel, largest first
  std::vector<size_t> ipt0(b.nb, 0);
  std::vector<int> rows;
  {
    size_t acc = 0;
    for (int ibx1 = 0; ibx1 < b.nb; ibx1++) {
      ipt0[ibx1] = acc;
      const int ibx2 = op.mat.ir2c[ibx1] - 1;
      if (ibx2 < 0) continue;
      acc += (size_t)b.db[ibx1] * b.db[ibx2];
      rows.push_back(ibx1);
    }
    std::stable_sort(rows.begin(), rows.end(), [&](int x, int y) {
      return (size_t)b.db[x] * b.db[op.mat.ir2c[x] - 1] > (size_t)b.db[y] * b.db[op.mat.ir2c[y] - 1];
    });
  }
#pragma omp parallel for schedule(dynamic, 1)
  for (int ir = 0; ir < (int)rows.size(); ir++) {
    const int ibx1 = rows[ir];
    const int ibx2 = op.mat.ir2c[ibx1] - 1;
    size_t ipt = ipt0[ibx1];
    const int nd1 = b.db[ibx1], nd2 = b.db[ibx2];
    for (int i2 = 0; i2 < nd2; i2++) {
      const int ix2 = i2 + b.isstart[ibx2] - 1;
      const double* wf2 = &b.wf[(size_t)ix2 * nghl];
      const double* dr2 = &b.wfdr[(size_t)ix2 * nghl];
      const double* dz2 = &b.wfdz[(size_t)ix2 * nghl];
      const int xl2 = b.nl[ix2], xs2 = b.ns[ix2];
      for (int i1 = 0; i1 < nd1; i1++, ipt++) {
        const int ix1 = i1 + b.isstart[ibx1] - 1;
        const double* wf1 = &b.wf[(size_t)ix1 * nghl];
        const int xl1 = b.nl[ix1], xs1 = b.ns[ix1];
        double me = 0.0;
        if (B == "F") {
          if (xl1 == xl2 && xs1 == xs2) me = dot(wf1, wf2);
        } else if (B == "GT") {
          auto dotg = [&](const double* a, const double* cc) {
            if (!rho_fac) return dot(a, cc);
            double s = 0;
            for (int i = 0; i < nghl; i++) s += a[i] * ((*rho_fac)[i] * cc[i]);
            return s;
          };
          if (K == 0) { if (xl1 == xl2 && xs1 == xs2) me = xs1 * dotg(wf1, wf2); }
          else { if (xl1 == xl2 && xs1 == xs2 + 2 * K) me = -K * sq2 * dotg(wf1, wf2); }
        } else if (B == "R") {
          if (K == 0) { if (xl1 == xl2 && xs1 == xs2) me = dotw(wf1, b.z, wf2); }
          else { if (xl1 == xl2 + K && xs1 == xs2) me = -K / sq2 * dotw(wf1, r, wf2); }
        } else if (B == "P") {
          if (K == 0) { if (xl1 == xl2 && xs1 == xs2) me = -dot(wf1, dz2); }
          else if (xl1 == xl2 + K && xs1 == xs2) me = K / sq2 * (dot(wf1, dr2) - K * xl2 * dotw(wf1, b.y, wf2));
        } else if (B == "RS0") {
          if (xl1 == xl2 && xs1 == xs2) me = -xs1 * dotw(wf1, b.z, wf2);
          else if (xl1 == xl2 + 1 && xs1 == xs2 - 2) me = -dotw(wf1, r, wf2);
          else if (xl1 == xl2 - 1 && xs1 == xs2 + 2) me = -dotw(wf1, r, wf2);
        } else if (B == "RS1") {
          if (K == 0) {
            if (xl1 == xl2 - 1 && xs1 == xs2 + 2) me = sq32 * dotw(wf1, r, wf2);
            else if (xl1 == xl2 + 1 && xs1 == xs2 - 2) me = -sq32 * dotw(wf1, r, wf2);
          } else {
            if (xl1 == xl2 && xs1 == xs2 + 2 * K) me = sq3 * dotw(wf1, b.z, wf2);
            else if (xl1 == xl2 + K && xs1 == xs2) me = -xs1 * sq3 / 2.0 * dotw(wf1, r, wf2);
          }
        } else if (B == "RS2") {
          if (K == 0) {
            if (xl1 == xl2 && xs1 == xs2) me = xs1 * sq2 * dotw(wf1, b.z, wf2);
            else if (xl1 == xl2 + 1 && xs1 == xs2 - 2) me = -1 / sq2 * dotw(wf1, r, wf2);
            else if (xl1 == xl2 - 1 && xs1 == xs2 + 2) me = -1 / sq2 * dotw(wf1, r, wf2);
          } else if (std::abs(K) == 1) {
            if (xl1 == xl2 && xs1 == xs2 + 2 * K) me = -K * sq3 * dotw(wf1, b.z, wf2);
            else if (xl1 == xl2 + K && xs1 == xs2) me = -K * xs1 * sq3 / 2.0 * dotw(wf1, r, wf2);
          } else {
            if (xl1 == xl2 + K / 2 && xs1 == xs2 + K) me = sq3 * dotw(wf1, r, wf2);
          }
        } else if (B == "PS0") {
          if (xl1 == xl2 && xs1 == xs2) me = -xs1 * dot(wf1, dz2);
          else if (xl1 == xl2 - 1 && xs1 == xs2 + 2) me = -dot(wf1, dr2) - xl2 * dotw(wf1, b.y, wf2);
          else if (xl1 == xl2 + 1 && xs1 == xs2 - 2) me = -dot(wf1, dr2) + xl2 * dotw(wf1, b.y, wf2);
        }
        op.mat.elem[ipt] = me;
      }
    }
  }
  return op;
}

std::vector<ExtField> make_crossterms(const FamBasis& b, const ExtField& op, const TwoBody* tb) {
  // the cross-term field R is always the one-body one (setup_crossterms, pnfam_extfield.f90:918-921)
  return make_crossterms(op, [&](const std::string& beta, const std::string& l, int k) {
    return make_external_field(b, beta, l, k, nullptr, l == "R" ? nullptr : tb);
  });
}

std::vector<ExtField> make_crossterms(const ExtField& op, const FieldProvider& field) {
  std::vector<ExtField> out;
  if (op.parity_even) return out;
  const std::string& L = op.label;
  const std::string beta = op.beta_minus ? "-" : "+";
  std::vector<std::string> labels;
  if (L == "R" || L == "P" || L == "RS1" || L == "RI" || L == "RS1I") labels = {"R", "RS1", "P", "RI", "RS1I"};
  else if (L == "RS0" || L == "PS0" || L == "RS0I") labels = {"RS0", "PS0", "RS0I"};
  for (const auto& l : labels) {
    ExtField g = field(beta, l, op.k);
    g.label = L + "x" + l;
    out.push_back(std::move(g));
  }
  return out;
}

bool read_tbc(const std::string& path, const FamBasis& b, const FamInput& in, ExtField& f, std::string& why, bool direct_only) {
  std::ifstream probe(path, std::ios::binary);
  if (!probe) { why = "cannot open " + path; return false; }
  probe.close();
  FortUnformatted fu(path);
  FortRecord r;
  const size_t nxy = f.mat.elem.size();
  const double c3new = in.two_body_current_lecs[0], c4new = in.two_body_current_lecs[1] + 0.25;
  bool found = false;
  while (fu.next(r)) {
    if (r.size() != 8 || r.get_str(8) != "extf_2bc") continue;
    auto need = [&]() { if (!fu.next(r)) throw std::runtime_error(".tbc: truncated file"); };
    need(); const int use_hblas = r.get<int32_t>();
    need(); /* mode */
    need(); const int usep = r.get<int32_t>();
    need(); /* lecs */
    need();
    const int nb_r = r.get<int32_t>(), dqp_r = r.get<int32_t>(), nxy_r = r.get<int32_t>(), nxy12_r = r.get<int32_t>();
    need();
    std::string label = r.get_str(std::min<size_t>(r.size(), 80));
    while (!label.empty() && (label.back() == ' ' || label.back() == 0)) label.pop_back();
    need(); const int k_r = r.get<int32_t>();
    need(); /* rank */
    need(); const int bm_r = r.get<int32_t>();
    need(); const int pe_r = r.get<int32_t>();
    need(); /* use_2bc */
    if (use_hblas != 1 || (usep == 0 && in.two_body_current_usep) || nb_r != b.nb || dqp_r != b.dqp || (size_t)nxy_r != nxy ||
        nxy12_r > 0 || label != f.label || k_r != f.k || bm_r != (f.beta_minus ? 1 : 0) || pe_r != (f.parity_even ? 1 : 0)) {
      why = "the file '" + path + "' is incompatible with the current calculation";
      return false;
    }
    std::vector<double> tot(nxy, 0.0);
    const double fac[4] = {c3new, c3new, c4new, c4new};
    for (int a = 0; a < 4 + (usep ? 2 : 0); a++) {
      need();
      if (r.size() != nxy * 8) throw std::runtime_error(".tbc: unexpected record size");
      std::vector<double> v = r.get_vec<double>(nxy);
      const double s = a < 4 ? fac[a] : 1.0;
      if (direct_only && (a & 1)) continue;                    // c3e, c4e, cpe: exchange parts
      for (size_t i = 0; i < nxy; i++) tot[i] += v[i] * s;
    }
    f.mat.elem = tot;
    found = true;
    break;
  }
  if (!found) why = "no extf_2bc record in " + path;
  return found;
}

void apply_two_body_current_gt(const std::string& tbc_path, const FamBasis& b, const FamInput& in, const TwoBody& tb, ExtField& f,
                               const HfbSolution* hfb, std::vector<std::string>* notes) {
  // setup_extfield (pnfam_solver.f90:586-646): F = [-GT_1body if 1BC+2BC] + GT[rho_fac] (+ Yukawa part from <name>.tbc
  // for the full-FAM mode).  rho_fac: contact term, plus the nuclear-matter exchange term in the LDA modes.
  std::vector<double> tmp(f.mat.elem.size(), 0.0);
  if (tb.u[1] == 1) for (size_t i = 0; i < tmp.size(); i++) tmp[i] = -f.mat.elem[i];   // Park sign convention: -sigma tau + 2BC
  const std::vector<double> rho_fac = tbc_gt_rho_fac(b, tb);
  ExtField contact = make_external_field(b, f.beta_minus ? "-" : "+", f.label, f.k, &rho_fac);
  for (size_t i = 0; i < tmp.size(); i++) tmp[i] += contact.mat.elem[i];
  const bool direct_only = tb.u[2] == 5;    // DME exchange term + the direct part of the FAM field
  if (tb.u[2] != 1 && tb.u[2] != 5) {       // no FAM part: F = -sigma tau + sigma tau f(rho)
    f.mat.elem = tmp;
    return;
  }
  std::string why;
  if (!read_tbc(tbc_path, b, in, f, why, direct_only)) {
    // "Starting calculation from scratch..." (fam_io(-1) failed, pnfam_solver.f90:622-640): compute and cache
    if (!hfb || getenv("PNFAM_B200_NO_TBC_GENERATOR"))
      throw std::runtime_error(why + " (the two-body-current field generator is switched off: provide the .tbc file)");
    // the reference's log lines (pnfam_storage.f90:690-702, pnfam_solver.f90:623-632)
    if (notes) {
      notes->push_back("  * " + why + ": starting calculation from scratch...");
      notes->push_back("  * Calculating 2BC matrix elements...");
    }
    const auto t_gen = std::chrono::steady_clock::now();
    const TbcField fld = generate_two_body_current_field(*hfb, b, f, in.two_body_current_usep);
    if (notes) {
      char buf[96];
      std::snprintf(buf, sizeof buf, "  * Two-body current wall time (min): %12.3E",
                    std::chrono::duration<double>(std::chrono::steady_clock::now() - t_gen).count() / 60.0);
      notes->push_back(buf);
    }
    const double c3 = in.two_body_current_lecs[0], c4 = in.two_body_current_lecs[1] + 0.25;
    const double fac[6] = {c3, c3, c4, c4, 1.0, 1.0};            // summed like read_tbc does: the next set-up, which reads
    std::fill(f.mat.elem.begin(), f.mat.elem.end(), 0.0);        // the cached file, gets bit-identical elements
    for (int a = 0; a < (in.two_body_current_usep ? 6 : 4); a++) {
      if (direct_only && (a & 1)) continue;
      for (size_t i = 0; i < tmp.size(); i++) f.mat.elem[i] += fld.c[a][i] * fac[a];
    }
    try { write_tbc(tbc_path, b, in, f, tb, fld); } catch (const std::exception&) { /* a read-only run directory: keep going */ }
  }
  else if (notes) notes->push_back("  * Reading from extfield file: " + tbc_path.substr(tbc_path.find_last_of('/') + 1));
  for (size_t i = 0; i < tmp.size(); i++) f.mat.elem[i] += tmp[i];
}

}  // namespace pnfam

// C ABI of the host set-up library (libpnfam_host.so): declared in include/pnfam_b200.h.
#include <algorithm>
#include <cstring>
#include <map>
#include <string>

#include <omp.h>

#include "../../../include/pnfam_b200.h"
#include "problem.hpp"

using namespace pnfam;

struct pnfam_problem {
  std::unique_ptr<Problem> p;
  std::map<std::string, std::vector<double>> f64_cache;
  std::map<std::string, std::vector<int32_t>> i32_cache;
};

static void set_err(char* err, int errlen, const std::string& s) {
  if (err && errlen > 0) {
    std::strncpy(err, s.c_str(), (size_t)errlen - 1);
    err[errlen - 1] = 0;
  }
}

extern "C" {

int pnfam_problem_create(const char* rundir, const char* namelist_file, pnfam_problem** out, char* err, int errlen) {
  try {
    auto h = new pnfam_problem;
    h->p = Problem::load(rundir ? rundir : ".", namelist_file ? namelist_file : "pnfam_NAMELIST.dat");
    *out = h;
    return 0;
  } catch (const std::exception& e) {
    set_err(err, errlen, e.what());
    return 1;
  }
}

int pnfam_problem_create_shared(const pnfam_problem* nucleus_of, const char* rundir, const char* namelist_file,
                                pnfam_problem** out, char* err, int errlen) {
  try {
    auto h = new pnfam_problem;
    h->p = Problem::load(rundir ? rundir : ".", namelist_file, nucleus_of->p->nuc);
    *out = h;
    return 0;
  } catch (const std::exception& e) {
    set_err(err, errlen, e.what());
    return 1;
  }
}

void pnfam_problem_destroy(pnfam_problem* p) { delete p; }

int pnfam_problem_scalar(const pnfam_problem* h, const char* name, double* out) {
  const Problem& p = *h->p;
  const FamBasis& b = p.nuc->basis;
  const Interaction& x = p.inter;
  const FamInput& in = p.in;
  const std::string n(name);
  std::map<std::string, double> m = {
      {"nb", b.nb}, {"dqp", b.dqp}, {"nghl", b.nghl}, {"ngh", b.ngh}, {"ngl", b.ngl}, {"sep_nzrows", b.sep_nzrows}, {"n_shells", b.n_shells}, {"dmat", (double)b.dmat},
      {"npr_n", b.npr[0]}, {"npr_p", b.npr[1]}, {"nxy", (double)p.f.mat.elem.size()}, {"nxterms", (double)p.g.size()},
      {"beta_minus", p.f.beta_minus}, {"blo_active", b.blo_active}, {"ft_active", b.ft_active}, {"ft_temp", b.ft_temp},
      {"statistical", b.statistical()},
      {"blo_qp_n", b.blo_qp[0]}, {"blo_qp_p", b.blo_qp[1]},
      {"skip_residual", x.skip_residual},
      {"cdrho", x.cdrho}, {"ctau", x.ctau}, {"ctj0", x.ctj0}, {"ctj1", x.ctj1}, {"ctj2", x.ctj2}, {"crdj", x.crdj},
      {"cds", x.cds}, {"ct", x.ct}, {"cj", x.cj}, {"cgs", x.cgs}, {"cf", x.cf}, {"csdj", x.csdj},
      {"cr0", x.cr0}, {"crr", x.crr}, {"cs0", x.cs0}, {"csr", x.csr}, {"sigma_r", x.sigma_r}, {"sigma_s", x.sigma_s}, {"sigma_pair", x.sigma_pair},
      {"cpair0", x.cpair0}, {"cpairr", x.cpairr}, {"cspair0", x.cspair0}, {"cspairr", x.cspairr},
      {"real_eqrpa", in.real_eqrpa}, {"imag_eqrpa", in.imag_eqrpa}, {"max_iter", in.max_iter},
      {"broyden_history_size", in.broyden_history_size}, {"convergence_epsilon", in.convergence_epsilon},
      {"quench_residual_int", x.skip_residual ? 0.0 : in.quench_residual_int},
      {"energy_shift_prot", in.energy_shift_prot}, {"energy_shift_neut", in.energy_shift_neut},
      {"ala_n", p.nuc->hfb.ala[0]}, {"ala_p", p.nuc->hfb.ala[1]},
      {"inner_n", p.nuc->hfb.inner[0]}, {"inner_p", p.nuc->hfb.inner[1]},
      {"setup_seconds", p.setup_seconds}, {"hfb_bz", p.nuc->hfb.bz}, {"hfb_bp", p.nuc->hfb.bp}, {"hfb_nt", p.nuc->hfb.nt}, {"hfb_nb", p.nuc->hfb.nb},
  };
  auto it = m.find(n);
  if (it == m.end()) return 1;
  *out = it->second;
  return 0;
}

int pnfam_problem_array_f64(pnfam_problem* h, const char* name, const double** ptr, int64_t* n) {
  const Problem& p = *h->p;
  const FamBasis& b = p.nuc->basis;
  const HfbSolution& s = p.nuc->hfb;
  const Interaction& x = p.inter;
  const std::string nm(name);
  const std::vector<double>* v = nullptr;
  std::map<std::string, const std::vector<double>*> m = {
      {"wf", &b.wf}, {"wfdr", &b.wfdr}, {"wfdz", &b.wfdz}, {"wfd2", &b.wfd2}, {"wfdp", &b.wfdp}, {"wfd2_all", &b.wfd2_all},
      {"y", &b.y}, {"z", &b.z}, {"sep_z", &b.sep_z}, {"sep_r", &b.sep_r}, {"wdcori", &b.wdcori}, {"Ep", &b.Ep}, {"En", &b.En},
      {"Up", &b.Up}, {"Vp", &b.Vp}, {"Un", &b.Un}, {"Vn", &b.Vn}, {"rho_n", &b.rho_n}, {"rho_p", &b.rho_p},
      {"qp_fn", &b.qp_fn}, {"qp_fp", &b.qp_fp},
      {"crho", &x.crho}, {"cs", &x.cs}, {"cpair", &x.cpair}, {"cspair", &x.cspair},
      {"f_elem", &p.f.mat.elem},
      {"hfb_hmat_n", &s.hmat[0]}, {"hfb_hmat_p", &s.hmat[1]}, {"hfb_dmat_n", &s.dmat[0]}, {"hfb_dmat_p", &s.dmat[1]},
      {"hfb_E_n", &s.E[0]}, {"hfb_E_p", &s.E[1]}, {"hfb_U_n", &s.U[0]}, {"hfb_V_n", &s.V[0]},
      {"hfb_U_p", &s.U[1]}, {"hfb_V_p", &s.V[1]}, {"hel_ro", &p.nuc->hel.ro},
  };
  auto it = m.find(nm);
  if (it != m.end()) v = it->second;
  if (!v && nm == "hfb_rk") {
    // HFBTHO's HO-basis density matrix rk(nqx, 2 nbx) = pnFAM's rmat (hfbtho_solution.f90:57): column ib = block ib of the
    // neutrons, column nbx + ib = of the protons, id(ib)^2 elements each, column-major; nqx = the largest id^2
    auto& c = h->f64_cache["hfb_rk"];
    if (c.empty()) {
      const TbcProblem pr = tbc_problem_from(s, p.f, false);
      size_t nqx = 0;
      for (int d : s.id) nqx = std::max(nqx, (size_t)d * d);
      c.assign(nqx * 2 * s.nb, 0.0);
      for (int it = 0; it < 2; it++) {
        size_t off = 0;
        for (int ib = 0; ib < s.nb; ib++) {
          const size_t d2 = (size_t)s.id[ib] * s.id[ib];
          for (size_t i = 0; i < d2; i++) c[(size_t)(it * s.nb + ib) * nqx + i] = 2.0 * pr.rho[it][off + i];
          off += d2;
        }
      }
    }
    v = &c;
  }
  if (!v && nm.rfind("g_elem_", 0) == 0) {
    size_t i = (size_t)std::stoi(nm.substr(7));
    if (i < p.g.size()) v = &p.g[i].mat.elem;
  }
  if (!v) return 1;
  *ptr = v->data();
  *n = (int64_t)v->size();
  return 0;
}

int pnfam_problem_array_i32(pnfam_problem* h, const char* name, const int32_t** ptr, int64_t* n) {
  const Problem& p = *h->p;
  const FamBasis& b = p.nuc->basis;
  const std::string nm(name);
  const std::vector<int>* v = nullptr;
  std::map<std::string, const std::vector<int>*> m = {
      {"db", &b.db}, {"isstart", &b.isstart}, {"nr", &b.nr}, {"nz", &b.nz}, {"nl", &b.nl}, {"ns", &b.ns},
      {"npar", &b.npar}, {"num_spin_up", &b.num_spin_up}, {"sep_zrow", &b.sep_zrow},
      {"f_ir2c", &p.f.mat.ir2c}, {"f_ic2r", &p.f.mat.ic2r}, {"f_ir2m", &p.f.mat.ir2m}, {"f_ic2m", &p.f.mat.ic2m},
      {"hfb_id", &p.nuc->hfb.id}, {"hfb_nz", &p.nuc->hfb.nz}, {"hfb_nr", &p.nuc->hfb.nr}, {"hfb_nl", &p.nuc->hfb.nl},
      {"hfb_ns", &p.nuc->hfb.ns},
  };
  auto it = m.find(nm);
  if (it != m.end()) v = it->second;
  if (!v && nm.rfind("g_ir2c_", 0) == 0) {
    size_t i = (size_t)std::stoi(nm.substr(7));
    if (i < p.g.size()) v = &p.g[i].mat.ir2c;
  }
  if (!v) return 1;
  static_assert(sizeof(int) == sizeof(int32_t), "int must be 32-bit");
  *ptr = reinterpret_cast<const int32_t*>(v->data());
  *n = (int64_t)v->size();
  return 0;
}

int pnfam_problem_label(const pnfam_problem* h, int which, char* out, int outlen) {
  const Problem& p = *h->p;
  std::string s;
  if (which == 0) s = p.f.label;
  else if (which >= 1 && (size_t)which <= p.g.size()) s = p.g[which - 1].label;
  else if (which == -1) s = p.inter.name;
  else if (which == -2) s = p.in.fam_output_filename;
  else return 1;
  set_err(out, outlen, s);
  return 0;
}

// effective_2bc_extfield with plain arrays (see include/pnfam_b200.h)
int pnfam_host_effective_2bc_extfield(int32_t ntx, int32_t nbx, const int32_t* id, const int32_t* hnz, const int32_t* hnr,
                                      const int32_t* hnl, const int32_t* hns, double bz, double bp, const double* rmat,
                                      int64_t ld_rmat, const int32_t* ir2c, const int32_t* ir2m, int64_t nxy, int32_t k,
                                      int32_t beta_minus, int32_t use_p, int32_t spin_sorted, double* c3d, double* c3e,
                                      double* c4d, double* c4e, double* cpd, double* cpe, char* err, int errlen) {
  try {
    if (ntx <= 0 || nbx <= 0 || !id || !hnz || !hnr || !hnl || !hns || !rmat || !ir2c || !ir2m || nxy <= 0)
      throw std::runtime_error("pnfam_host_effective_2bc_extfield: missing argument");
    TbcProblem pr;
    pr.nt = ntx; pr.nb = nbx;
    pr.id.assign(id, id + nbx);
    pr.nz.assign(hnz, hnz + ntx); pr.nr.assign(hnr, hnr + ntx); pr.nl.assign(hnl, hnl + ntx); pr.ns.assign(hns, hns + ntx);
    pr.bz = bz; pr.bp = bp;
    int tot = 0;
    for (int ib = 0; ib < nbx; ib++) {
      if (id[ib] <= 0 || (int64_t)id[ib] * id[ib] > ld_rmat) throw std::runtime_error("pnfam_host_effective_2bc_extfield: block larger than ld_rmat");
      tot += id[ib];
    }
    if (tot != ntx) throw std::runtime_error("pnfam_host_effective_2bc_extfield: sum of id differs from ntx");
    for (int it = 0; it < 2; it++)
      for (int ib = 0; ib < nbx; ib++) {
        const double* col = rmat + (size_t)(it * nbx + ib) * ld_rmat;
        for (int64_t i = 0; i < (int64_t)id[ib] * id[ib]; i++) pr.rho[it].push_back(0.5 * col[i]);      // rho_db = rmat / 2
      }
    pr.ir2c.assign(ir2c, ir2c + 2 * nbx); pr.ir2m.assign(ir2m, ir2m + 2 * nbx);
    pr.nxy = (size_t)nxy; pr.K = k; pr.beta_minus = beta_minus != 0; pr.use_p = use_p != 0; pr.spin_sorted = spin_sorted != 0;
    const TbcField f = generate_two_body_current_field(pr);
    double* out[6] = {c3d, c3e, c4d, c4e, cpd, cpe};
    for (int o = 0; o < 6; o++)
      if (out[o]) std::memcpy(out[o], f.c[o].data(), (size_t)nxy * sizeof(double));
    return 0;
  } catch (const std::exception& e) {
    set_err(err, errlen, e.what());
    return 1;
  }
}

// number of OpenMP threads of the host set-up (HFB reconstruction, tables, external fields); n <= 0: leave unchanged.
// Returns the previous setting.
int pnfam_host_set_threads(int n) {
  const int prev = omp_get_max_threads();
  if (n > 0) omp_set_num_threads(n);
  return prev;
}

}  // extern "C"

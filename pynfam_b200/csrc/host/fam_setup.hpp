// Host-side FAM set-up: everything pnfam_main.x prepares before entering the iteration
// (exes/pnfam/pnfam_solver.f90:44-62): the time-reversal-doubled, spin-sorted basis
// (hfbtho_solution.f90:137-459), the residual-interaction couplings (pnfam_interaction.f90:90-260),
// the external-field matrices and their block maps (pnfam_extfield.f90:37-108,110-622,763-840,
// 882-949) and the pnFAM namelist (pnfam_setup.f90:98-108,178-250).
#pragma once
#include <functional>
#include <array>
#include <string>
#include <vector>

#include "hfb_front.hpp"

namespace pnfam {

// Block-sparse matrix with at most one non-zero block per block row/column
// (pnfam_type_blockmatrix.f90:15-26).  Blocks are stored in increasing block-row order,
// column-major inside a block; ir2c/ic2r are 1-based partner indices (0 = none), ir2m/ic2m are
// 1-based offsets into elem.
struct BlockMatrix {
  std::vector<int> ir2c, ic2r, ir2m, ic2m;
  std::vector<double> elem;
  void init(int nb, size_t n) {
    ir2c.assign(nb, 0); ic2r.assign(nb, 0); ir2m.assign(nb, 0); ic2m.assign(nb, 0);
    elem.assign(n, 0.0);
  }
};

// The pnFAM namelist (&general &ext_field &interaction &solver), defaults of pnfam_setup.f90:178-210.
struct FamInput {
  std::string namelist_path = "pnfam_NAMELIST.dat";
  // general
  std::string fam_output_filename;
  bool print_stdout = true;
  int use_fam_storage = 0;
  double real_eqrpa = 0.0, imag_eqrpa = 0.5;
  // ext_field
  std::string beta_type = "-", operator_name = "F";
  int operator_k = 0;
  bool compute_crossterms = false;
  int two_body_current_mode = 0;
  bool two_body_current_usep = false;
  double two_body_current_lecs[3] = {-3.4, 5.4, 0.0};
  // solver
  int max_iter = 200, broyden_history_size = 50;
  double convergence_epsilon = 1e-7, energy_shift_prot = 0, energy_shift_neut = 0, quench_residual_int = 1.0;
  // interaction
  std::string interaction_name = "NONE";
  bool require_self_consistency = true, require_gauge_invariance = false, force_j2_terms = false;
  bool has_vpair0 = false, has_vpair1 = false, has_vpair_t0 = false, has_vpair_t1 = false;
  double vpair0 = 0, vpair1 = 0, vpair_t0 = 0, vpair_t1 = 0;
  // overrides: cs0 csr cds ct cf cgs cj csdj
  bool has_override[8] = {false, false, false, false, false, false, false, false};
  double override_val[8] = {0, 0, 0, 0, 0, 0, 0, 0};

  static FamInput read(const std::string& path);
};

// Doubled (explicit time-reversed partners), spin-sorted single-particle basis + HFB solution.
struct FamBasis {
  int nb = 0, dqp = 0, nghl = 0, n_shells = 0;
  size_t dmat = 0;
  int npr[3] = {0, 0, 0};
  std::vector<int> db, isstart;          // isstart 1-based like the reference
  std::vector<int> nr, nz, nl, ns, npar, num_spin_up;
  std::vector<double> wf, wfdr, wfdz, wfd2, wfdp, wfd2_all;  // (nghl, dqp) column-major
  std::vector<double> y, z, wdcori;
  // separable factors of the tables above (HO basis): wf = Z0 R0, wfdr = Z0 R1, wfdp = Z0 R2, wfdz = Z1 R0,
  // wfd2_all = Z2 R0 + Z0 R3 with Z*[sep_zrow[state]][ih], R*[state][il] and ihil = ih + il*ngh
  int ngh = 0, ngl = 0, sep_nzrows = 0;
  std::vector<int> sep_zrow;             // [dqp]
  std::vector<double> sep_z;             // [3][sep_nzrows][ngh]
  std::vector<double> sep_r;             // [4][dqp][ngl]
  std::vector<double> Ep, En, Up, Vp, Un, Vn;
  std::vector<double> rho_n, rho_p;      // coordinate-space densities (normalised)
  // isoscalar kinetic density and Laplacian of the density (tau_n + tau_p, Delta rho_n + Delta rho_p) for the
  // density-matrix-expansion two-body currents; computed from `src` on first use
  const HfbSolution* src = nullptr;
  mutable std::vector<double> tau0, d2rho0;
  void need_tau_d2rho() const;
  bool blo_active = false;
  int blo_qp[2] = {0, 0};                // 1-based overall qp index of the blocked level (n, p)
  int blo_ib[2] = {0, 0}, blo_is[2] = {0, 0};
  std::vector<double> qp_fn, qp_fp;      // quasiparticle occupations: equal filling (blo_active) or Fermi-Dirac (ft_active)
  bool ft_active = false;                // finite-temperature HFB solution (pnfam_setup.f90:323-331)
  double ft_temp = 0;                    // MeV
  bool statistical() const { return blo_active || ft_active; }   // P,Q quadrants and T factors in the FAM
  // functional data handed over from HFBTHO
  double hfb_cpair[2] = {0, 0}, hfb_alpha_pair[2] = {0, 0}, rho_nm = 0.16, hbzero = 0;
  double hfb_cr0 = 0, hfb_crr = 0, hfb_cdrho = 0, hfb_ctau = 0, hfb_ctj = 0, hfb_crdj = 0;
  bool hfb_use_j2terms = false;

  static FamBasis build(const HfbSolution& s);
};

// Residual interaction couplings (pnfam_interaction.f90).
struct Interaction {
  bool skip_residual = false;
  std::string name;
  double cr0 = 0, crr = 0, cs0 = 0, csr = 0, sigma_r = 0, sigma_s = 0;
  double cpair0 = 0, cpairr = 0, cspair0 = 0, cspairr = 0, sigma_pair = 1;
  double cdrho = 0, ctau = 0, ctj0 = 0, ctj1 = 0, ctj2 = 0, crdj = 0, cds = 0, ct = 0, cj = 0, cgs = 0, cf = 0, csdj = 0;
  std::vector<double> crho, cs, cpair, cspair;  // (nghl)
  std::vector<std::string> notes;               // "[!] ..." lines for the log

  // rundir: where a `FILE:<name>` interaction file is looked up (the reference opens it relative to its working directory)
  static Interaction build(const FamInput& in, const FamBasis& b, const std::string& rundir = ".");
};

struct ExtField {
  std::string label;
  bool beta_minus = true, parity_even = true;
  int k = 0, rank = 0;
  BlockMatrix mat;
};

// rho_fac (optional, nghl): weight folded into the GT operator, sigma tau f(rho) -- the contact part of the
// two-body current (pnfam_extfield.f90:167-190)
// Two-body-current request of one field (set_use_2bc, pnfam_extfield.f90:690-720): u[1..6] = the six digits of
// two_body_current_mode (1: 1 = 1BC+2BC, 2 = 2BC only; 2: 1 full FAM, 2 SNM+LDA, 3 ASNM+LDA, 4/5 DME; 3: 1 = Gamma;
// 4: GT current active, >= 2 also corrects RS*; 5: P current; 6: PS0 current), all zero = one-body field.
struct TwoBody {
  int u[7] = {0, 0, 0, 0, 0, 0, 0};
  double lecs[3] = {0, 0, 0};      // c3, c4, cd of the namelist
  bool use_p = false;
  bool active() const { return u[1] != 0; }
};
// closed-form density-dependent factors of the nuclear-matter / LDA two-body currents (pnfam_extfield.f90:976-1410)
std::vector<double> tbc_gt_rho_fac(const FamBasis& b, const TwoBody& tb);   // GT: contact + SNM / ASNM exchange term
std::vector<double> tbc_rsl_correction(const FamBasis& b, const TwoBody& tb, bool snm);
std::vector<double> tbc_p_correction(const FamBasis& b);
std::vector<double> tbc_ps0_correction(const FamBasis& b);
// density-matrix-expansion variants (pnfam_extfield.f90:1195-1330, 1380-1545, 1573-1662): exchange term of the GT current
// (2nd mode digit 4, 5), vector current of P (5th digit 2), axial charge of PS0 (6th digit 2)
std::vector<double> tbc_dme_exc(const FamBasis& b, const TwoBody& tb);
std::vector<double> tbc_dme_vector(const FamBasis& b);
std::vector<double> tbc_dme_axial(const FamBasis& b);
ExtField make_external_field(const FamBasis& b, const std::string& beta_type, const std::string& label, int k,
                             const std::vector<double>* rho_fac = nullptr, const TwoBody* tb = nullptr);
std::vector<ExtField> make_crossterms(const FamBasis& b, const ExtField& op, const TwoBody* tb = nullptr);
// the same with the fields taken from a provider (beta type, label, K) -> field, e.g. a per-nucleus cache: the
// operators of one J^pi group share their cross-term fields
using FieldProvider = std::function<ExtField(const std::string&, const std::string&, int)>;
std::vector<ExtField> make_crossterms(const ExtField& op, const FieldProvider& field);
// Read the Yukawa part of a two-body-current field from <name>.tbc (pnfam_storage.f90:562-727).
// On success f.mat.elem holds c3/c4-weighted gamma (direct + exchange) exactly as read_tbc leaves it.
// direct_only: the direct (Hartree) part alone, calc_totgamdel(..., 'd') -- mode digit 2 = 5 (DME exchange + FAM direct)
bool read_tbc(const std::string& path, const FamBasis& b, const FamInput& in, ExtField& f, std::string& why, bool direct_only = false);
// Full two-body-current GT field of mode i1 i2=1 i3=1 i4>0 (pnfam_solver.f90:596-652):
//   F = [-GT_1body if i1==1] + GT[contact rho_fac] + Yukawa part from <name>.tbc
//   hfb: the undoubled HFB solution; when the file is absent or does not fit the calculation the Yukawa part is computed
//   (generate_two_body_current_field) and cached in <name>.tbc like the reference does; nullptr = read only
void apply_two_body_current_gt(const std::string& tbc_path, const FamBasis& b, const FamInput& in, const TwoBody& tb, ExtField& f,
                               const HfbSolution* hfb = nullptr, std::vector<std::string>* notes = nullptr);

// The Yukawa part of the full-FAM two-body-current GT field (effective_2bc_extfield, pnfam_extfield_2bc.f90:26-465;
// csrc/host/tbc_generator.cpp), low-energy constants stripped as in the .tbc file: c[0..5] = c3 direct, c3 exchange,
// c4 direct, c4 exchange, momentum direct, momentum exchange (the last two zero without use_p), each of the size and
// element order of f.mat.elem.  Gamma part; even and blocked nuclei, zero and finite temperature.
struct TbcField { std::vector<double> c[6]; };
// The contraction described by plain arrays -- what effective_2bc_extfield takes from the modules hfb_solution, hfbtho and
// type_blockmatrix: the undoubled HFBTHO basis (blocks id, quantum numbers of every state, oscillator lengths), the
// HO-basis density matrices rho = rmat / 2 of neutrons and protons (packed block by block, column-major id x id), the block
// structure of the operator in the doubled basis (1-based ir2c, ir2m), K, beta type.  spin_sorted: rows and columns of
// every block in the spin-sorted order of the USE_HBLAS = 1 build (the shipped one), else the original order.
struct TbcProblem {
  int nt = 0, nb = 0;
  std::vector<int> id, nz, nr, nl, ns;
  double bz = 0, bp = 0;
  std::vector<double> rho[2];
  std::vector<int> ir2c, ir2m;
  size_t nxy = 0;
  int K = 0;
  bool beta_minus = true, use_p = false, spin_sorted = true;
};
TbcField generate_two_body_current_field(const TbcProblem& pr);
TbcProblem tbc_problem_from(const HfbSolution& s, const ExtField& f, bool use_p);
TbcField generate_two_body_current_field(const HfbSolution& s, const FamBasis& b, const ExtField& f, bool use_p);
// <name>.tbc in the reference's record layout (write_tbc, pnfam_storage.f90:488-559; written atomically)
void write_tbc(const std::string& path, const FamBasis& b, const FamInput& in, const ExtField& f, const TwoBody& tb, const TbcField& fld);

}  // namespace pnfam

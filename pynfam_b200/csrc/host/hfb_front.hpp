// Host front-end: rebuild the HFB solution (U, V, E, rho(r), basis tables) that the FAM iteration
// consumes, from the two files pnfam_main.x reads in its working directory:
//   hfbtho_NAMELIST.dat + hfbtho_output.hel      (exes/pnfam/hfbtho_interface.f90:33,
//                                                 hfbtho_io.f90:209)
// The reference does this by running HFBTHO with zero iterations (hfbtho_interface.f90:20-226,
// hfbtho_solver.f90:262-539).  This is a one-off per nucleus, CPU-side, and is NOT the hot path;
// it exists here because the image has no Fortran compiler to reuse the reference's own setup.
#pragma once
#include <string>
#include <vector>

#include "fortio.hpp"

namespace pnfam {

// Contents of hfbtho_output.hel (record layout: hfbtho_io.f90:745-897).
struct HelData {
  int version = 0;
  int Z = -1, N = -1;
  bool set_temperature = false;
  // SkyFunct
  bool use_j2terms = false, finite_range = false;
  std::string skyrme;
  double rho_nm = 0, sigma = 0, hbzero = 0, hb0 = 0, hb0n = 0, hb0p = 0;
  double Crho[2], Cdrho[2], Ctau[2], CrDr[2], CrdJ[2], CJ[2], CpV0[2], CpV1[2];
  // HO-Basis
  double b0 = 0, bz = 0, bp = 0;
  int n00 = 0, nb = 0, nt = 0, ngh = 0, ngl = 0, nleg = 0;
  std::vector<double> xh, xl, wh, wl;
  // QuantNum
  std::vector<int> id, nr, nz, nl, ns;
  // Various.
  double si = 1, etot = 0, bet = 0, xmix = 0, pwi = 0;
  double rms[3], del[2], ept[3], ala[2], ala2[2], alast[2], tz[2];
  // Densits., FieldsN., FieldsP.
  std::vector<double> ro, aka;          // (nghl,2)
  std::vector<double> fld[2][11];       // v, vhb, vr, vz, vd, vs, vSFIZ, vSZFI, vSFIR, vSRFI, dv
  // Blocking
  bool has_blocking = false;
  int bloall = 0;
  std::vector<int> bloblo, blo123, blok1k2;  // (0:bloall-1? see reader) stored (bloall,2) col-major
  int blomax[2] = {0, 0};
  std::vector<double> bloqpdif;
  bool blocking_never_done[2] = {true, true};
  bool has_hfb_matrix = false;

  static HelData read(const std::string& path);
};

// hfbtho_NAMELIST.dat keys the zero-iteration reconstruction consumes.
struct HfbInput {
  int n_shells = 0, proton_number = 0, neutron_number = 0, type_of_calculation = 1;
  std::string functional;
  bool user_pairing = false;
  double vpair_n = 0, vpair_p = 0, pairing_cutoff = 60, pairing_feature = 0.5;
  int neutron_blocking[5] = {0, 0, 0, 0, 0}, proton_blocking[5] = {0, 0, 0, 0, 0};
  bool set_temperature = false;
  double temperature = 0;
  bool force_parity = true, compatibility_hfodd = false;

  static HfbInput read(const std::string& path);
};

// The HFB solution in HFBTHO's own (undoubled, K>0) basis.
struct HfbSolution {
  // basis
  int nb = 0, nt = 0, ngh = 0, ngl = 0, nghl = 0, n_shells = 0;
  double b0 = 0, bz = 0, bp = 0;
  std::vector<int> id, ia, nr, nz, nl, ns, npar;
  std::vector<double> y, z, wdcor, wdcori;          // (nghl): 1/r, z, weights
  std::vector<double> qhla, fi1r, fi1z, fi2d;       // [state][nghl]
  // separable factors of the same tables (ihil = ih + il*ngh):  qhla = Z0 R0, fi1r = Z0 R1, fi1z = Z1 R0,
  // fi2d = Z2 R0 + Z0 R3c;  Z*[nz][ih] depend on n_z only, R*[state][il] on (n_r, Lambda)
  int sep_nzrows = 0;
  std::vector<double> sep_z;                        // [3][sep_nzrows][ngh]
  std::vector<double> sep_r;                        // [3][nt][ngl]: R0, R1, R3c
  int npr[3] = {0, 0, 0};                           // N, Z, A of the (possibly odd) nucleus
  // per isospin it = 0 (n), 1 (p)
  std::vector<double> hmat[2], dmat[2];             // packed lower triangles per block (gamdel)
  std::vector<double> E[2], U[2], V[2];             // E (nt); U,V blocks [block][qp k][basis n] packed
  std::vector<int> ka[2], kd[2], Kqp[2], Kpwi[2];   // pairing-window bookkeeping (hfbdiag)
  std::vector<double> occ[2];                       // uk(kl)
  double ala[2] = {0, 0};                           // Fermi levels used for the final diagonalisation
  double ala_out[2] = {0, 0};                       // after the last ALambda
  int inner[2] = {0, 0};
  int klmax[2] = {0, 0};
  // blocking (odd nuclei, equal-filling approximation)
  int keyblo[2] = {0, 0}, blo_block[2] = {0, 0}, blo_state[2] = {0, 0}, blok1k2d[2] = {0, 0};
  // finite temperature (hfbtho_solver.f90:1749-1761, 1987-2022): Fermi-Dirac occupations of the quasiparticles inside
  // the pairing window as the last ALambda call left them (the values DENSIT weighs the densities with)
  bool ft_active = false;
  double temper = 0;
  std::vector<double> fT_pwi[2];
  // densities
  std::vector<double> ro[2];                        // (nghl) normalised rho_n, rho_p
  // kinetic density tau and Laplacian of rho, same normalisation as DENSIT leaves them (hfbtho_solver.f90:4545-4549,
  // 4697-4716); only the density-matrix-expansion two-body currents need them: computed on demand
  void kinetic_and_laplacian(std::vector<double> tau[2], std::vector<double> dro[2]) const;
  // functional info carried for the FAM interaction set-up
  double CpV0[2], CpV1[2], rho_nm = 0.16, hbzero = 0;
  double hfb_cr0 = 0, hfb_crr = 0, hfb_cdrho = 0, hfb_ctau = 0, hfb_ctj = 0, hfb_crdj = 0;
  bool use_j2terms = false;
  double pwi = 0;

  // cache_file (optional): where the results of the expensive stages (gamdel, hfbdiag, DENSIT: E, U, V, pairing-window
  // bookkeeping, Fermi levels, blocking indices, rho) are kept between processes; cache_key = hash of the two input
  // files.  A file with a matching key is loaded instead of recomputing (bit-identical doubles); otherwise the
  // solution is computed and the file (re)written atomically.
  static HfbSolution build(const HfbInput& in, const HelData& hel, const std::string& cache_file = "",
                           unsigned long long cache_key = 0);
  bool from_cache = false;
};

// FNV-1a hash of the bytes of a file, continued from `seed` (0 = start); 0 if the file cannot be read.
unsigned long long hash_file(const std::string& path, unsigned long long seed = 0);

// Symmetric eigen-solver (Householder tridiagonalisation + implicit QL), ascending eigenvalues.
// a: n x n column-major, lower triangle referenced; on exit z (n x n col-major) = eigenvectors.
void sym_eig(int n, const double* a_lower, double* w, double* z);

}  // namespace pnfam

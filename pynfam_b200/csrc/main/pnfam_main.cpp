// pnfam_main.x -- drop-in for the reference's executable (exes/pnfam/pnfam_prog.f90:8-18):
//   pnfam_main.x [namelist]        cwd holds hfbtho_NAMELIST.dat + hfbtho_output.hel (+ <name>.tbc)
// writes <fam_output_filename>.dat in the text contract that pynfam parses
// (pynfam/outputs/pnfam_parser.py:33-132, exes/pnfam/pnfam_txtoutput.f90:230-253) and keeps stderr silent on
// success (pynfam treats any stderr text as a failed task, pynfam/fortran/fortran_utils.py:255).
// The host set-up is libpnfam_host; the iteration runs on the GPU through the C ABI of libpnfam_b200.
//
// Extension (absent from the reference): an optional namelist group
//   &b200_batch  real_eqrpa = r1, r2, ...  imag_eqrpa = i1, i2, ... /
// solves all listed frequencies of the operator in ONE batched call and writes <name>_<k>.dat (k = 0..)
// next to <name>.dat (which then holds the first point), so that one launch amortises the HFB
// reconstruction that the reference repeats for every point.
#include "driver_common.hpp"

using namespace pnfam;
using namespace pnfam_driver;

int main(int argc, char** argv) {
  std::string namelist = "pnfam_NAMELIST.dat";
  if (argc == 2) namelist = argv[1];
  if (argc > 2) {
    std::printf(" More than one command-line argument! Aborting.\n");
    return 0;  // the reference `stop`s with exit code 0; failure is detected from the missing result table
  }
  auto t0 = std::chrono::steady_clock::now();
  std::unique_ptr<Problem> prob;
  try {
    prob = Problem::load(".", namelist);
  } catch (const std::exception& e) {
    std::printf("# ERROR: %s\n", e.what());
    return 0;
  }
  const Problem& p = *prob;
  // optional batch of frequencies
  std::vector<double> wre = {p.in.real_eqrpa}, wim = {p.in.imag_eqrpa};
  try {
    Namelist nl = Namelist::parse_file(p.in.namelist_path);
    if (nl.has("b200_batch", "real_eqrpa") && nl.has("b200_batch", "imag_eqrpa")) {
      const auto& r = nl.raw("b200_batch", "real_eqrpa");
      const auto& i = nl.raw("b200_batch", "imag_eqrpa");
      if (r.size() == i.size() && !r.empty()) {
        wre.clear(); wim.clear();
        for (size_t k = 0; k < r.size(); k++) { wre.push_back(std::atof(r[k].c_str())); wim.push_back(std::atof(i[k].c_str())); }
      }
    }
  } catch (...) {
  }
  char err[1024] = {0};
  pnfam_b200_ctx* ctx = make_context(p, 0, err, sizeof err);
  if (!ctx) {
    std::printf("# ERROR: %s\n", err);
    return 0;
  }
  BatchResult r;
  if (solve_batch(ctx, p, wre, wim, r, err, sizeof err) != 0) {
    std::printf("# ERROR: %s\n", err);
    pnfam_b200_ctx_destroy(ctx);
    return 0;
  }
  const double total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  const std::string base = p.in.fam_output_filename;
  const int P = (int)wre.size();
  for (int k = 0; k < P; k++) {
    std::string out = base;
    if (P > 1 && k > 0 && !base.empty()) out = base + "_" + std::to_string(k);
    const bool so = (base.empty() || p.in.print_stdout) && k == 0;
    write_dat(p, namelist, out, wre[k], wim[k], &r.strength[(size_t)k * r.nstr * 2], r.iters[k], r.conv[k], r.si[k],
              &r.trace[(size_t)k * r.tstride], p.in.max_iter, total / P, so);
  }
  pnfam_b200_ctx_destroy(ctx);
  return 0;
}

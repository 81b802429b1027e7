// contour_main.x -- drop-in for the reference's contour executable (exes/pnfam/contour_prog.f90:9-52,
// contour_setup.f90): solve the FAM equations of one or more operators on a set of complex energies.
//   contour_main.x [contour-namelist]      default pnfam_CONTOUR.dat; cwd as for pnfam_main.x
// Namelists (contour_setup.f90:19-21, defaults :68-79): &ctr_general fam_mode, fam_input_filename /
// &ctr_extfield operator_groups, operator_active / &str_parameters energy_start, energy_step, nr_points, half_width /.
// Outputs, as the reference: one summary "<OP><beta>K<k>.out" per operator (write_ctr_output_header / _point,
// contour_setup.f90:195-231) and, when the pnfam namelist names an output file, one "<OP><beta>K<k>_<iiiiii>.dat"
// per point (apply_task_values, :262-275).
//
// Where the reference loops over tasks calling pnfam_solve (HFB reconstruction, set-up and a serial iteration per
// task), this program sets the nucleus up once, uploads it once, and solves ALL points of an operator in one batched
// GPU call.  fam_mode = 'STR' is the straight line of the reference; 'CONTOUR' (marked "not implemented" in
// README_ctr.md) is provided as the circle pynfam integrates on (pynfam/strength/contour.py:212-283) through
//   &contour_parameters energy_min, energy_max, nr_points, theta_init, shift_imag, max_height /
// with equally spaced theta (the Gauss-Legendre variant is driven from pynfam_b200/strength.py).
#include <algorithm>
#include <cctype>
#include <complex>

#include "driver_common.hpp"

using namespace pnfam;
using namespace pnfam_driver;

namespace {

struct Task { std::string op; int k; };

std::string upper(std::string s) {
  for (auto& c : s) c = (char)std::toupper((unsigned char)c);
  return s;
}
std::string txtr(const std::string& s, int n) { return (int)s.size() >= n ? s : std::string(n - s.size(), ' ') + s; }
int fsign(int a, int b) { return b >= 0 ? std::abs(a) : -std::abs(a); }   // Fortran sign(a, b)

// contour_setup.f90:130-186 setup_operators
std::vector<Task> operator_list(const std::vector<int>& active, const FamInput& in) {
  std::vector<Task> t;
  const int ks = in.operator_k;
  auto add = [&](const char* op, int k) { t.push_back({op, fsign(k, ks)}); };
  for (size_t i = 0; i < active.size() && i < 5; i++) {
    if (!active[i]) continue;
    switch (i) {
      case 0: add("F", 0); break;
      case 1: add("GT", 0); add("GT", 1); break;
      case 2: add("RS0", 0); add("PS0", 0); break;
      case 3: add("R", 0); add("R", 1); add("P", 0); add("P", 1); add("RS1", 0); add("RS1", 1); break;
      case 4: add("RS2", 0); add("RS2", 1); add("RS2", 2); break;
    }
  }
  if (t.empty()) t.push_back({in.operator_name, in.operator_k});
  return t;
}

}  // namespace

int main(int argc, char** argv) {
  std::string ctr_file = "pnfam_CONTOUR.dat";
  if (argc == 2) ctr_file = argv[1];
  if (argc > 2) {
    std::printf(" More than one command-line argument! Aborting.\n");
    return 0;
  }
  auto t0 = std::chrono::steady_clock::now();
  std::string mode = "STR", fam_input = "pnfam_NAMELIST.dat";
  std::vector<int> active(5, 0);
  std::vector<double> wre, wim;
  double half_width = 0.5;
  try {
    Namelist nl = Namelist::parse_file(ctr_file);
    mode = upper(nl.get_string("ctr_general", "fam_mode", "STR"));
    fam_input = nl.get_string("ctr_general", "fam_input_filename", "pnfam_NAMELIST.dat");
    active = nl.get_ints("ctr_extfield", "operator_active", {0, 0, 0, 0, 0});
    if (mode == "STR") {
      // contour_setup.f90:114-128 setup_energy_contour
      const double e0 = nl.get_double("str_parameters", "energy_start", 0.0), de = nl.get_double("str_parameters", "energy_step", 0.2);
      const int n = nl.get_int("str_parameters", "nr_points", 5);
      half_width = nl.get_double("str_parameters", "half_width", 0.5);
      for (int i = 0; i < n; i++) { wre.push_back(e0 + i * de); wim.push_back(half_width); }
    } else if (mode == "CONTOUR") {
      const double emin = nl.get_double("contour_parameters", "energy_min", 0.0), emax = nl.get_double("contour_parameters", "energy_max", 10.0);
      const int n = nl.get_int("contour_parameters", "nr_points", 60);
      const double pi = 3.14159265358979323846;
      const double th0 = nl.get_double("contour_parameters", "theta_init", pi), shift = nl.get_double("contour_parameters", "shift_imag", 0.0);
      const double hmax = nl.get_double("contour_parameters", "max_height", 30.0);
      const double r0 = 0.5 * (emax + emin), r = emax - r0;
      half_width = 0.0;
      for (int i = 0; i < n; i++) {
        const double th = n > 1 ? th0 + 2.0 * pi * i / (n - 1) : th0;
        const std::complex<double> z = r > hmax ? std::complex<double>(r0 + r * std::cos(th), hmax * std::sin(th))
                                                : r0 + r * std::exp(std::complex<double>(0.0, th));
        wre.push_back(z.real()); wim.push_back(z.imag() + shift);
      }
    } else {
      std::printf(" ERROR: fam_mode '%s' is not available (STR, CONTOUR).\n", mode.c_str());
      return 0;
    }
  } catch (const std::exception& e) {
    std::printf(" ERROR: problem reading contour namelist: %s\n", e.what());
    return 0;
  }
  if (wre.empty()) {
    std::printf(" ERROR: the contour has no points.\n");
    return 0;
  }

  FamInput base;
  std::shared_ptr<Nucleus> nuc;
  try {
    base = FamInput::read(fam_input);
    nuc = Nucleus::load(".");
  } catch (const std::exception& e) {
    std::printf("# ERROR: %s\n", e.what());
    return 0;
  }
  const std::vector<Task> tasks = operator_list(active, base);
  pnfam_b200_ctx* ctx = nullptr;
  char err[1024] = {0};
  for (const Task& tk : tasks) {
    auto top = std::chrono::steady_clock::now();
    FamInput in = base;
    in.operator_name = tk.op; in.operator_k = tk.k;
    const std::string opname = tk.op + in.beta_type + "K" + std::to_string(tk.k);
    std::unique_ptr<Problem> prob;
    try {
      prob = Problem::build(".", in, nuc);
    } catch (const std::exception& e) {
      std::printf("# ERROR (%s): %s\n", opname.c_str(), e.what());
      continue;
    }
    const Problem& p = *prob;
    if (!ctx) ctx = make_context(p, 0, err, sizeof err);   // nucleus + interaction: the same for every operator
    if (!ctx) {
      std::printf("# ERROR: %s\n", err);
      return 0;
    }
    BatchResult r;
    if (solve_batch(ctx, p, wre, wim, r, err, sizeof err) != 0) {
      std::printf("# ERROR (%s): %s\n", opname.c_str(), err);
      continue;
    }
    const int P = (int)wre.size();
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - top).count();
    // per-point .dat files
    if (!base.fam_output_filename.empty()) {
      for (int k = 0; k < P; k++) {
        char nm[256];
        std::snprintf(nm, sizeof nm, "%s_%06d", opname.c_str(), k);
        Problem& pm = *prob;
        pm.in.fam_output_filename = nm;
        write_dat(pm, fam_input, nm, wre[k], wim[k], &r.strength[(size_t)k * r.nstr * 2], r.iters[k], r.conv[k], r.si[k],
                  &r.trace[(size_t)k * r.tstride], p.in.max_iter, secs / P, false);
      }
    }
    // summary file (contour_setup.f90:195-231)
    FILE* f = std::fopen((opname + ".out").c_str(), "w");
    if (!f) {
      std::printf("# ERROR: cannot open %s.out\n", opname.c_str());
      continue;
    }
    std::fprintf(f, "# pnFAM code version: 2.00-b200\n");
    std::fprintf(f, "# pnFAM code commit:  B200-native solver (batched contour)\n");
    std::fprintf(f, "# Contour input file name: %s\n", ctr_file.c_str());
    std::fprintf(f, "# Residual interaction: %s\n", in.interaction_name.c_str());
    std::fprintf(f, "# Operator: %s with K = %d\n", tk.op.c_str(), tk.k);
    {
      char hw[64];
      std::snprintf(hw, sizeof hw, "%.4f", half_width);           // Fortran F0.4 drops the leading zero
      std::string s = hw;
      if (s.rfind("0.", 0) == 0) s = s.substr(1);
      else if (s.rfind("-0.", 0) == 0) s = "-" + s.substr(2);
      std::fprintf(f, "# Gamma (half-width): %s\n", s.c_str());
    }
    std::fprintf(f, "#\n");
    std::string h = "#" + txtr("Conv", 5) + txtr("Re(EQRPA)", 29) + txtr("Re(Strength)", 34) + txtr("Im(Strength)", 34);
    for (const auto& g : p.g) h += txtr("Re(" + g.label + ")", 34) + txtr("Im(" + g.label + ")", 34);
    std::fprintf(f, "%s\n", h.c_str());
    for (int k = 0; k < P; k++) {
      std::fprintf(f, "%5d%30.19f", r.conv[k] ? 1 : 0, wre[k]);
      for (int j = 0; j < r.nstr; j++)
        std::fprintf(f, "%34.19E%34.19E", r.strength[((size_t)k * r.nstr + j) * 2], r.strength[((size_t)k * r.nstr + j) * 2 + 1]);
      std::fprintf(f, "\n");
    }
    std::fclose(f);
    if (base.print_stdout)
      std::printf("# %-10s %4d points, %6lld iterations, %8.3f s\n", opname.c_str(), P, (long long)r.stats.iterations, secs);
  }
  if (ctx) pnfam_b200_ctx_destroy(ctx);
  if (base.print_stdout)
    std::printf("# Total time = %.3E minutes\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / 60.0);
  return 0;
}

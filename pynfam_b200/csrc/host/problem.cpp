#include "problem.hpp"

#include <cctype>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>

namespace pnfam {

std::shared_ptr<Nucleus> Nucleus::load(const std::string& rundir) {
  auto n = std::make_shared<Nucleus>();
  const std::string d = rundir.empty() ? std::string(".") : rundir;
  // file names are hard-coded in the reference (hfbtho_interface.f90:33, hfbtho_io.f90:209)
  auto t0 = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (getenv("PNFAM_B200_SETUP_TIMING")) {
      auto t1 = std::chrono::steady_clock::now();
      std::fprintf(stderr, "[setup] %-22s %.3f s\n", what, std::chrono::duration<double>(t1 - t0).count());
      t0 = t1;
    }
  };
  n->hfb_in = HfbInput::read(d + "/hfbtho_NAMELIST.dat");
  n->hel = HelData::read(d + "/hfbtho_output.hel");
  lap("read files");
  // cache of the reconstructed solution next to the input files (PNFAM_B200_CACHE_DIR: elsewhere; PNFAM_B200_NO_CACHE:
  // off), keyed by the bytes of both files: pynfam launches one pnfam_main.x per omega point in directories holding
  // copies of the same two files, and every rank of a sharded run sets up the same nucleus
  std::string cache_file;
  unsigned long long key = 0;
  if (!getenv("PNFAM_B200_NO_CACHE")) {
    key = hash_file(d + "/hfbtho_output.hel", hash_file(d + "/hfbtho_NAMELIST.dat"));
    const char* cd = getenv("PNFAM_B200_CACHE_DIR");
    char name[64];
    std::snprintf(name, sizeof name, "/.pnfam_b200_hfb_%016llx.cache", key);
    cache_file = (cd && *cd ? std::string(cd) : d) + name;
  }
  n->hfb = HfbSolution::build(n->hfb_in, n->hel, cache_file, key);
  lap(n->hfb.from_cache ? "HFB solution (cache)" : "HFB reconstruction");
  n->basis = FamBasis::build(n->hfb);
  lap("FAM basis and tables");
  return n;
}

const ExtField& Nucleus::field(const std::string& beta, const std::string& label, int k) {
  const std::string key = beta + "|" + label + "|" + std::to_string(k);
  auto it = fields.find(key);
  if (it == fields.end()) it = fields.emplace(key, make_external_field(basis, beta, label, k)).first;
  return it->second;
}

static int digit(int v, int p) {
  for (int i = 0; i < p; i++) v /= 10;
  return v % 10;
}

std::unique_ptr<Problem> Problem::load(const std::string& rundir, const std::string& namelist,
                                       std::shared_ptr<Nucleus> nuc) {
  const std::string d = rundir.empty() ? std::string(".") : rundir;
  std::string nml = namelist;
  if (!nml.empty() && nml[0] != '/') nml = d + "/" + nml;
  return build(rundir, FamInput::read(nml), nuc);
}

std::unique_ptr<Problem> Problem::build(const std::string& rundir, const FamInput& input, std::shared_ptr<Nucleus> nuc) {
  auto t0 = std::chrono::steady_clock::now();
  auto p = std::make_unique<Problem>();
  const std::string d = rundir.empty() ? std::string(".") : rundir;
  p->in = input;
  p->nuc = nuc ? nuc : Nucleus::load(d);
  FamBasis& b = p->nuc->basis;
  p->inter = Interaction::build(p->in, b, d);
  const FamInput& in = p->in;
  // external field (setup_extfield, pnfam_solver.f90:556-654)
  const int mode = in.two_body_current_mode;
  TwoBody tb;
  if (mode != 0) {
    int* u = tb.u;
    u[1] = digit(mode, 5); u[2] = digit(mode, 4); u[3] = digit(mode, 3);
    u[4] = digit(mode, 2); u[5] = digit(mode, 1); u[6] = digit(mode, 0);
    if (u[1] < 1 || u[1] > 2 || u[2] < 1 || u[2] > 5 || u[3] < 1 || u[3] > 3 || (u[1] != 1 && u[3] != 1))
      throw std::runtime_error("Invalid value supplied for two_body_current_mode.");
    if (u[3] != 1) throw std::runtime_error("This two_body_current_mode is not yet operational.");
    for (int i = 0; i < 3; i++) tb.lecs[i] = in.two_body_current_lecs[i];
    tb.use_p = in.two_body_current_usep;
  }
  Nucleus& nucl = *p->nuc;
  // fields with a two-body-current weight depend on the mode: they bypass the per-nucleus cache of one-body fields
  auto field_of = [&](const std::string& beta, const std::string& l, int k, bool with_2bc) -> ExtField {
    if (mode == 0 || !with_2bc) return nucl.field(beta, l, k);
    return make_external_field(b, beta, l, k, nullptr, &tb);
  };
  {
    // only these operators take the mode as the MAIN field (pnfam_solver.f90:574-578); every cross-term field except R
    // takes it (setup_crossterms, pnfam_extfield.f90:910-944)
    std::string L = in.operator_name;
    for (auto& ch : L) ch = (char)std::toupper((unsigned char)ch);
    const bool main_2bc = L == "P" || L == "PS0" || L == "RS0" || L == "RS1" || L == "RS2" || L == "RS0I" || L == "RS1I";
    p->f = field_of(in.beta_type, in.operator_name, in.operator_k, main_2bc);     // a copy: the GT correction below edits it
  }
  if (in.compute_crossterms)
    p->g = make_crossterms(p->f, [&](const std::string& beta, const std::string& l, int k) { return field_of(beta, l, k, l != "R"); });
  if (mode != 0 && p->f.label == "GT" && tb.u[4] != 0)
    apply_two_body_current_gt(d + "/" + in.fam_output_filename + ".tbc", b, in, tb, p->f, &p->nuc->hfb, &p->notes);
  p->setup_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return p;
}

}  // namespace pnfam

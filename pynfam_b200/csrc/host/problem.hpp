// One pnFAM problem as pnfam_main.x sees it: a working directory holding hfbtho_NAMELIST.dat +
// hfbtho_output.hel and one pnFAM namelist -> basis, HFB solution, interaction, external field(s).
#pragma once
#include <map>
#include <memory>
#include <string>

#include "fam_setup.hpp"

namespace pnfam {

struct Nucleus {  // everything that depends only on the HFB files (shared by all operators / omegas)
  HfbInput hfb_in;
  HelData hel;
  HfbSolution hfb;
  FamBasis basis;
  // external fields already built for this nucleus, keyed by beta type / label / K (shared by the operators of a
  // contour run: R, P, RS1 ... all need the same five cross-term fields)
  std::map<std::string, ExtField> fields;
  const ExtField& field(const std::string& beta, const std::string& label, int k);
  static std::shared_ptr<Nucleus> load(const std::string& rundir);
};

struct Problem {
  std::shared_ptr<Nucleus> nuc;
  FamInput in;
  Interaction inter;
  ExtField f;
  std::vector<ExtField> g;
  double setup_seconds = 0;
  std::vector<std::string> notes;       // set-up messages of the reference's log (two-body-current file read / computed)
  static std::unique_ptr<Problem> load(const std::string& rundir, const std::string& namelist,
                                       std::shared_ptr<Nucleus> nuc = nullptr);
  // the same from an already parsed (and possibly overridden) namelist -- what the contour driver does per task
  // (contour_setup.f90:262-275 apply_task_values)
  static std::unique_ptr<Problem> build(const std::string& rundir, const FamInput& in, std::shared_ptr<Nucleus> nuc = nullptr);
};

}  // namespace pnfam

#include "fortio.hpp"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <sstream>

namespace pnfam {

static std::string lower(std::string s) {
  for (auto& c : s) c = (char)std::tolower((unsigned char)c);
  return s;
}
static std::string trim(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && std::isspace((unsigned char)s[a])) a++;
  while (b > a && std::isspace((unsigned char)s[b - 1])) b--;
  return s.substr(a, b - a);
}

// Split "v1, v2 , 'a,b'" at commas outside quotes; drops empty trailing tokens.
static std::vector<std::string> split_values(const std::string& v) {
  std::vector<std::string> out;
  std::string cur;
  char q = 0;
  for (char c : v) {
    if (q) {
      cur.push_back(c);
      if (c == q) q = 0;
    } else if (c == '\'' || c == '"') {
      q = c;
      cur.push_back(c);
    } else if (c == ',') {
      out.push_back(trim(cur));
      cur.clear();
    } else {
      cur.push_back(c);
    }
  }
  if (!trim(cur).empty()) out.push_back(trim(cur));
  // also split whitespace-separated numeric lists ("1 2 3")
  std::vector<std::string> out2;
  for (auto& t : out) {
    if (t.empty()) continue;
    if (t[0] == '\'' || t[0] == '"') {
      out2.push_back(t);
      continue;
    }
    std::istringstream is(t);
    std::string w;
    while (is >> w) out2.push_back(w);
  }
  return out2;
}

Namelist Namelist::parse_file(const std::string& path) {
  std::ifstream in(path);
  if (!in) throw std::runtime_error("cannot open namelist file " + path);
  Namelist nl;
  std::string line, group, lastkey;
  while (std::getline(in, line)) {
    // strip comments (outside quotes)
    {
      char q = 0;
      for (size_t i = 0; i < line.size(); i++) {
        char c = line[i];
        if (q) {
          if (c == q) q = 0;
        } else if (c == '\'' || c == '"') {
          q = c;
        } else if (c == '!') {
          line.resize(i);
          break;
        }
      }
    }
    std::string t = trim(line);
    if (t.empty()) continue;
    if (t[0] == '&' || t[0] == '$') {
      std::string rest = trim(t.substr(1));
      size_t sp = rest.find_first_of(" \t");
      group = lower(rest.substr(0, sp));
      nl.groups[group];
      lastkey.clear();
      if (sp == std::string::npos) continue;
      t = trim(rest.substr(sp));
      if (t.empty()) continue;
    }
    if (group.empty()) continue;
    // a group may end on the same line as values
    bool end_group = false;
    if (t == "/" || lower(t) == "&end" || lower(t) == "$end") {
      group.clear();
      continue;
    }
    if (!t.empty() && t.back() == '/') {      // a terminating slash outside quotes
      char q = 0;
      for (size_t i = 0; i + 1 < t.size(); i++) {
        if (q) { if (t[i] == q) q = 0; }
        else if (t[i] == '\'' || t[i] == '"') q = t[i];
      }
      if (!q) {
        end_group = true;
        t = trim(t.substr(0, t.size() - 1));
      }
    }
    // Several "key = value" assignments may share a line (the namelists shipped with HFBTHO and the reference's own
    // test inputs do); lines without '=' continue the value list of the previous key.
    std::vector<size_t> eqs;                 // positions of '=' outside quotes
    {
      char q = 0;
      for (size_t i = 0; i < t.size(); i++) {
        char c = t[i];
        if (q) {
          if (c == q) q = 0;
        } else if (c == '\'' || c == '"') {
          q = c;
        } else if (c == '=') {
          eqs.push_back(i);
        }
      }
    }
    if (eqs.empty()) {
      if (!lastkey.empty()) {
        auto more = split_values(t);
        auto& dst = nl.groups[group][lastkey];
        dst.insert(dst.end(), more.begin(), more.end());
      }
    } else {
      // start of the name in front of every '=': skip blanks, an optional (section), then the identifier
      std::vector<size_t> starts;
      for (size_t eq : eqs) {
        size_t i = eq;
        while (i > 0 && std::isspace((unsigned char)t[i - 1])) i--;
        if (i > 0 && t[i - 1] == ')') {
          while (i > 0 && t[i - 1] != '(') i--;
          if (i > 0) i--;
          while (i > 0 && std::isspace((unsigned char)t[i - 1])) i--;
        }
        while (i > 0 && (std::isalnum((unsigned char)t[i - 1]) || t[i - 1] == '_' || t[i - 1] == '%')) i--;
        starts.push_back(i);
      }
      // text in front of the first name continues the previous key
      if (starts[0] > 0 && !lastkey.empty()) {
        std::string head = trim(t.substr(0, starts[0]));
        if (!head.empty() && head != ",") {
          auto more = split_values(head);
          auto& dst = nl.groups[group][lastkey];
          dst.insert(dst.end(), more.begin(), more.end());
        }
      }
      for (size_t k = 0; k < eqs.size(); k++) {
        std::string key = lower(trim(t.substr(starts[k], eqs[k] - starts[k])));
        // drop array-section syntax like key(1:3)
        size_t par = key.find('(');
        if (par != std::string::npos) key = trim(key.substr(0, par));
        const size_t vend = k + 1 < eqs.size() ? starts[k + 1] : t.size();
        std::string val = trim(t.substr(eqs[k] + 1, vend - eqs[k] - 1));
        // the comma that separates this assignment from the next one is not a (null) value
        if (k + 1 < eqs.size() && !val.empty() && val.back() == ',') val = trim(val.substr(0, val.size() - 1));
        nl.groups[group][key] = split_values(val);
        lastkey = key;
      }
    }
    if (end_group) group.clear();
  }
  return nl;
}

static double parse_real(const std::string& s0) {
  std::string s = lower(s0);
  std::replace(s.begin(), s.end(), 'd', 'e');
  char* end = nullptr;
  double v = std::strtod(s.c_str(), &end);
  if (end == s.c_str()) throw std::runtime_error("namelist: bad real '" + s0 + "'");
  return v;
}
static bool parse_bool(const std::string& s0) {
  std::string s = lower(s0);
  if (s == ".true." || s == "t" || s == ".t." || s == "true") return true;
  if (s == ".false." || s == "f" || s == ".f." || s == "false") return false;
  throw std::runtime_error("namelist: bad logical '" + s0 + "'");
}

double Namelist::get_double(const std::string& g, const std::string& k, double dflt) const {
  return has(g, k) ? parse_real(raw(g, k)[0]) : dflt;
}
int Namelist::get_int(const std::string& g, const std::string& k, int dflt) const {
  return has(g, k) ? (int)std::lround(parse_real(raw(g, k)[0])) : dflt;
}
bool Namelist::get_bool(const std::string& g, const std::string& k, bool dflt) const {
  return has(g, k) ? parse_bool(raw(g, k)[0]) : dflt;
}
std::string Namelist::get_string(const std::string& g, const std::string& k, const std::string& dflt) const {
  if (!has(g, k)) return dflt;
  std::string s = raw(g, k)[0];
  if (s.size() >= 2 && (s[0] == '\'' || s[0] == '"')) s = s.substr(1, s.size() - 2);
  return trim(s);
}
std::vector<double> Namelist::get_doubles(const std::string& g, const std::string& k, std::vector<double> dflt) const {
  if (!has(g, k)) return dflt;
  const auto& r = raw(g, k);
  for (size_t i = 0; i < r.size() && i < dflt.size(); i++) dflt[i] = parse_real(r[i]);
  return dflt;
}
std::vector<int> Namelist::get_ints(const std::string& g, const std::string& k, std::vector<int> dflt) const {
  if (!has(g, k)) return dflt;
  const auto& r = raw(g, k);
  for (size_t i = 0; i < r.size() && i < dflt.size(); i++) dflt[i] = (int)std::lround(parse_real(r[i]));
  return dflt;
}

}  // namespace pnfam

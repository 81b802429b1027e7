// Host-side I/O helpers: Fortran sequential-unformatted records and Fortran namelists.
//
// The reference reads its inputs with the Fortran runtime (READ(unit) ..., READ(unit,nml=...));
// this is the equivalent for the two file kinds on the pnFAM path:
//   * hfbtho_output.hel  -- sequential unformatted, little-endian, 4-byte record markers
//                           (reference reader: hfbtho_io.f90:487-738, writer :745-897)
//   * *.in / hfbtho_NAMELIST.dat -- namelists (pnfam_setup.f90:98-108, :214-250;
//                           hfbtho_variables.f90:360-379)
#pragma once
#include <cstdint>
#include <cstring>
#include <fstream>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace pnfam {

struct FortRecord {
  std::vector<char> bytes;
  size_t pos = 0;
  template <class T> T get() {
    if (pos + sizeof(T) > bytes.size()) throw std::runtime_error("fortran record underflow");
    T v;
    std::memcpy(&v, bytes.data() + pos, sizeof(T));
    pos += sizeof(T);
    return v;
  }
  template <class T> void get_array(T* dst, size_t n) {
    if (pos + n * sizeof(T) > bytes.size()) throw std::runtime_error("fortran record underflow");
    std::memcpy(dst, bytes.data() + pos, n * sizeof(T));
    pos += n * sizeof(T);
  }
  template <class T> std::vector<T> get_vec(size_t n) {
    std::vector<T> v(n);
    get_array(v.data(), n);
    return v;
  }
  std::string get_str(size_t n) {
    std::string s(bytes.data() + pos, bytes.data() + pos + n);
    pos += n;
    return s;
  }
  size_t size() const { return bytes.size(); }
};

class FortUnformatted {
 public:
  explicit FortUnformatted(const std::string& path) : in_(path, std::ios::binary) {
    if (!in_) throw std::runtime_error("cannot open " + path);
  }
  // Returns false at end of file.
  bool next(FortRecord& rec) {
    int32_t n = 0;
    in_.read(reinterpret_cast<char*>(&n), 4);
    if (!in_ || in_.gcount() != 4) return false;
    if (n < 0) throw std::runtime_error("fortran record: negative length (split records unsupported)");
    rec.bytes.resize(n);
    rec.pos = 0;
    in_.read(rec.bytes.data(), n);
    int32_t m = 0;
    in_.read(reinterpret_cast<char*>(&m), 4);
    if (!in_ || m != n) throw std::runtime_error("fortran record: corrupt markers");
    return true;
  }

 private:
  std::ifstream in_;
};

// ---------------------------------------------------------------------------------------------
// Namelists: {group -> {key -> list of raw value tokens}}; keys and groups lower-cased.
// A blank value ("key = ,") yields an empty token list (Fortran keeps the default).
// ---------------------------------------------------------------------------------------------
struct Namelist {
  std::map<std::string, std::map<std::string, std::vector<std::string>>> groups;

  static Namelist parse_file(const std::string& path);

  bool has(const std::string& g, const std::string& k) const {
    auto it = groups.find(g);
    if (it == groups.end()) return false;
    auto jt = it->second.find(k);
    return jt != it->second.end() && !jt->second.empty();
  }
  const std::vector<std::string>& raw(const std::string& g, const std::string& k) const {
    return groups.at(g).at(k);
  }
  double get_double(const std::string& g, const std::string& k, double dflt) const;
  int get_int(const std::string& g, const std::string& k, int dflt) const;
  bool get_bool(const std::string& g, const std::string& k, bool dflt) const;
  std::string get_string(const std::string& g, const std::string& k, const std::string& dflt) const;
  std::vector<double> get_doubles(const std::string& g, const std::string& k, std::vector<double> dflt) const;
  std::vector<int> get_ints(const std::string& g, const std::string& k, std::vector<int> dflt) const;
};

}  // namespace pnfam

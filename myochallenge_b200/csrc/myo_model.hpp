// Host-side model: MuJoCo 2.1.0 MJB reader (product path; the checker-side twin is oracle/mjb.py).
// Replaces mujoco_py.load_model_from_mjb as reached from MyoSuite BaseV0.__init__ with the
// `model_path` kwargs of /root/reference/src/envs/__init__.py:17,29,44,62.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace myo {

enum DType { DT_F64 = 0, DT_I32 = 1, DT_U8 = 2, DT_F32 = 3, DT_CHAR = 4 };

struct Array {
  std::string name;
  int dtype = DT_F64;
  int rows = 0, cols = 1;
  std::vector<uint8_t> bytes;
  template <class T> T* as() { return reinterpret_cast<T*>(bytes.data()); }
  template <class T> const T* as() const { return reinterpret_cast<const T*>(bytes.data()); }
  size_t count() const { return (size_t)rows * (size_t)cols; }
};

struct Model {
  std::map<std::string, int> sizes;        // nq, nv, ...
  std::vector<std::string> size_order;
  std::map<std::string, double> opt;       // mjOption + mjStatistic by name
  std::vector<Array> arrays;               // MJMODEL_POINTERS order
  std::map<std::string, int> index;        // name -> arrays index
  std::vector<std::string> id2name_cache;  // storage for returned c-strings

  int sz(const char* n) const { auto it = sizes.find(n); return it == sizes.end() ? 0 : it->second; }
  Array* arr(const std::string& n) { auto it = index.find(n); return it == index.end() ? nullptr : &arrays[it->second]; }
  const Array* arr(const std::string& n) const { auto it = index.find(n); return it == index.end() ? nullptr : &arrays[it->second]; }
  const double* d(const char* n) const { return arr(n)->as<double>(); }
  const int* i(const char* n) const { return arr(n)->as<int>(); }
  const uint8_t* b(const char* n) const { return arr(n)->as<uint8_t>(); }
  std::string name_of(const char* group, int id) const;
  int name2id(const char* group, const char* name) const;
};

// returns "" on success, else an error message; status receives a myo_status code
std::string load_mjb(const uint8_t* raw, size_t len, Model& out, int& status);
void set_error(const std::string& msg);

}  // namespace myo

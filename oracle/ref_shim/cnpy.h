// Stand-in for cnpy (https://github.com/rogersce/cnpy, unpinned in the reference: SRC/CMakeLists.txt:51,78),
// used ONLY to compile the reference's own MPPI sources (oracle/refbuild.py).  TEST INFRASTRUCTURE.
// Backed by the repo's stored-zip/npy reader; implements the three calls the reference makes:
// cnpy::npz_load, npz_t::operator[], NpyArray::data<T>() (PI/neural_net_model.cu:82-89, PI/costs.cu:195-216).
#ifndef REF_SHIM_CNPY_H_
#define REF_SHIM_CNPY_H_
#include <map>
#include <string>
#include <vector>
#include "../../include/autorally_control/path_integral/npz_io.h"

namespace cnpy {
struct NpyArray {
  std::vector<size_t> shape;
  size_t word_size = 0;
  std::vector<unsigned char> bytes;
  template <class T> T *data() { return reinterpret_cast<T *>(bytes.data()); }
};
typedef std::map<std::string, NpyArray> npz_t;
inline npz_t npz_load(const std::string &path) {
  npz_t out;
  autorally_control::npz::Archive a = autorally_control::npz::load(path);
  for (auto &kv : a) {
    NpyArray n;
    n.shape = kv.second.shape;
    n.word_size = kv.second.word_size;
    n.bytes = kv.second.bytes;
    out[kv.first] = n;
  }
  return out;
}
}  // namespace cnpy
#endif

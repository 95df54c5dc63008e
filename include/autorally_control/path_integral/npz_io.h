// npz_io.h -- minimal reader/writer for numpy .npz archives (stored, i.e. uncompressed, members;
// npy format 1.0/2.0; little-endian f4/f8/i4/i8).  Replaces the cnpy dependency of the reference
// (cnpy::npz_load, PI/neural_net_model.cu:80, PI/generalized_linear.cu:99, PI/costs.cu:195).
// Entries are located through the zip central directory, so archives written by current numpy
// (zip64 local headers) and by cnpy both load.
#ifndef MPPI_NPZ_IO_H_
#define MPPI_NPZ_IO_H_
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace autorally_control {
namespace npz {

struct Array {
  std::vector<size_t> shape;
  char kind = 'f';        // 'f' float, 'i' int, 'u' unsigned
  size_t word_size = 0;   // bytes per element
  bool fortran_order = false;
  std::vector<unsigned char> bytes;
  size_t num_vals() const { size_t n = 1; for (size_t s : shape) n *= s; return n; }
  template <class T> const T *data() const { return reinterpret_cast<const T *>(bytes.data()); }
  // element i converted to double regardless of the stored type
  double at(size_t i) const {
    if (kind == 'f' && word_size == 8) return data<double>()[i];
    if (kind == 'f' && word_size == 4) return data<float>()[i];
    if (kind == 'i' && word_size == 8) return (double)data<int64_t>()[i];
    if (kind == 'i' && word_size == 4) return (double)data<int32_t>()[i];
    throw std::runtime_error("npz: unsupported dtype");
  }
};

typedef std::map<std::string, Array> Archive;

namespace detail {
inline uint16_t rd16(const unsigned char *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline uint32_t rd32(const unsigned char *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint64_t rd64(const unsigned char *p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }

inline Array parse_npy(const unsigned char *p, size_t n) {
  if (n < 10 || std::memcmp(p, "\x93NUMPY", 6) != 0) throw std::runtime_error("npz: bad npy magic");
  const int major = p[6];
  size_t hlen, hoff;
  if (major == 1) { hlen = rd16(p + 8); hoff = 10; } else { hlen = rd32(p + 8); hoff = 12; }
  if (hoff + hlen > n) throw std::runtime_error("npz: truncated npy header");
  const std::string hdr(reinterpret_cast<const char *>(p + hoff), hlen);
  Array a;
  size_t d = hdr.find("'descr'");
  size_t q1 = hdr.find('\'', hdr.find(':', d)), q2 = hdr.find('\'', q1 + 1);
  const std::string descr = hdr.substr(q1 + 1, q2 - q1 - 1);  // e.g. "<f8"
  if (descr.size() < 3 || (descr[0] != '<' && descr[0] != '|' && descr[0] != '=')) throw std::runtime_error("npz: unsupported byte order " + descr);
  a.kind = descr[1];
  a.word_size = (size_t)std::stoul(descr.substr(2));
  a.fortran_order = hdr.find("True", hdr.find("'fortran_order'")) != std::string::npos &&
                    hdr.find("True", hdr.find("'fortran_order'")) < hdr.find(',', hdr.find("'fortran_order'"));
  size_t s1 = hdr.find('(', hdr.find("'shape'")), s2 = hdr.find(')', s1);
  std::string sh = hdr.substr(s1 + 1, s2 - s1 - 1);
  size_t pos = 0;
  while (pos < sh.size()) {
    while (pos < sh.size() && (sh[pos] == ' ' || sh[pos] == ',')) pos++;
    if (pos >= sh.size()) break;
    size_t end = pos;
    while (end < sh.size() && sh[end] >= '0' && sh[end] <= '9') end++;
    if (end == pos) break;
    a.shape.push_back((size_t)std::stoull(sh.substr(pos, end - pos)));
    pos = end;
  }
  const size_t nbytes = a.num_vals() * a.word_size;
  if (hoff + hlen + nbytes > n) throw std::runtime_error("npz: truncated npy payload");
  a.bytes.assign(p + hoff + hlen, p + hoff + hlen + nbytes);
  return a;
}
}  // namespace detail

inline Archive load(const std::string &path) {
  using namespace detail;
  std::ifstream f(path.c_str(), std::ios::binary);
  if (!f) throw std::runtime_error("npz: cannot open " + path);
  std::vector<unsigned char> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  if (buf.size() < 22) throw std::runtime_error("npz: file too small " + path);
  // end-of-central-directory record
  size_t eocd = std::string::npos;
  for (size_t i = buf.size() - 22 + 1; i-- > 0;) {
    if (rd32(&buf[i]) == 0x06054b50u) { eocd = i; break; }
    if (buf.size() - i > 22 + 65535) break;
  }
  if (eocd == std::string::npos) throw std::runtime_error("npz: no zip directory in " + path);
  size_t nent = rd16(&buf[eocd + 10]);
  uint64_t cd_off = rd32(&buf[eocd + 16]);
  if (cd_off == 0xffffffffu || nent == 0xffff) {  // zip64 end-of-central-directory
    for (size_t i = eocd; i-- > 0;)
      if (rd32(&buf[i]) == 0x06064b50u) { nent = (size_t)rd64(&buf[i + 32]); cd_off = rd64(&buf[i + 48]); break; }
  }
  Archive out;
  size_t p = (size_t)cd_off;
  for (size_t e = 0; e < nent; e++) {
    if (p + 46 > buf.size() || rd32(&buf[p]) != 0x02014b50u) throw std::runtime_error("npz: bad central directory");
    const uint16_t method = rd16(&buf[p + 10]);
    uint64_t csize = rd32(&buf[p + 20]), usize = rd32(&buf[p + 24]), lho = rd32(&buf[p + 42]);
    const uint16_t nlen = rd16(&buf[p + 28]), xlen = rd16(&buf[p + 30]), clen = rd16(&buf[p + 32]);
    std::string name(reinterpret_cast<const char *>(&buf[p + 46]), nlen);
    // zip64 extra field
    size_t x = p + 46 + nlen, xend = x + xlen;
    while (x + 4 <= xend) {
      const uint16_t id = rd16(&buf[x]), sz = rd16(&buf[x + 2]);
      if (id == 0x0001) {
        size_t q = x + 4;
        if (usize == 0xffffffffu) { usize = rd64(&buf[q]); q += 8; }
        if (csize == 0xffffffffu) { csize = rd64(&buf[q]); q += 8; }
        if (lho == 0xffffffffu) { lho = rd64(&buf[q]); q += 8; }
      }
      x += 4 + sz;
    }
    if (method != 0) throw std::runtime_error("npz: member " + name + " is compressed; save with numpy.savez (not savez_compressed)");
    const size_t l = (size_t)lho;
    if (l + 30 > buf.size() || rd32(&buf[l]) != 0x04034b50u) throw std::runtime_error("npz: bad local header");
    const size_t data = l + 30 + rd16(&buf[l + 26]) + rd16(&buf[l + 28]);
    if (data + usize > buf.size()) throw std::runtime_error("npz: truncated member " + name);
    if (name.size() > 4 && name.substr(name.size() - 4) == ".npy") name = name.substr(0, name.size() - 4);
    out[name] = parse_npy(&buf[data], (size_t)usize);
    p += 46 + nlen + xlen + clen;
  }
  return out;
}

// ---- writer (stored members, npy 1.0) ----
namespace detail {
inline uint32_t crc32(const unsigned char *p, size_t n) {
  static uint32_t table[256];
  static bool init = false;
  if (!init) {
    for (uint32_t i = 0; i < 256; i++) { uint32_t c = i; for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1; table[i] = c; }
    init = true;
  }
  uint32_t c = 0xffffffffu;
  for (size_t i = 0; i < n; i++) c = table[(c ^ p[i]) & 0xff] ^ (c >> 8);
  return c ^ 0xffffffffu;
}
inline void put16(std::vector<unsigned char> &v, uint16_t x) { v.push_back(x & 0xff); v.push_back(x >> 8); }
inline void put32(std::vector<unsigned char> &v, uint32_t x) { for (int i = 0; i < 4; i++) v.push_back((x >> (8 * i)) & 0xff); }
}  // namespace detail

class Writer {
 public:
  template <class T>
  void add(const std::string &name, const T *vals, const std::vector<size_t> &shape) {
    const char *descr = sizeof(T) == 8 ? "<f8" : "<f4";
    std::string sh = "(";
    size_t n = 1;
    for (size_t i = 0; i < shape.size(); i++) { sh += std::to_string(shape[i]) + (shape.size() == 1 || i + 1 < shape.size() ? "," : ""); n *= shape[i]; }
    sh += ")";
    std::string hdr = std::string("{'descr': '") + descr + "', 'fortran_order': False, 'shape': " + sh + ", }";
    while ((10 + hdr.size() + 1) % 64 != 0) hdr += ' ';
    hdr += '\n';
    std::vector<unsigned char> npy;
    const char magic[] = "\x93NUMPY\x01\x00";
    npy.insert(npy.end(), magic, magic + 8);
    detail::put16(npy, (uint16_t)hdr.size());
    npy.insert(npy.end(), hdr.begin(), hdr.end());
    const unsigned char *raw = reinterpret_cast<const unsigned char *>(vals);
    npy.insert(npy.end(), raw, raw + n * sizeof(T));
    members_.push_back(std::make_pair(name + ".npy", npy));
  }
  void save(const std::string &path) const {
    std::vector<unsigned char> out, cd;
    for (size_t m = 0; m < members_.size(); m++) {
      const std::string &name = members_[m].first;
      const std::vector<unsigned char> &d = members_[m].second;
      const uint32_t crc = detail::crc32(d.data(), d.size()), off = (uint32_t)out.size();
      detail::put32(out, 0x04034b50u); detail::put16(out, 20); detail::put16(out, 0); detail::put16(out, 0);
      detail::put16(out, 0); detail::put16(out, 0x21); detail::put32(out, crc);
      detail::put32(out, (uint32_t)d.size()); detail::put32(out, (uint32_t)d.size());
      detail::put16(out, (uint16_t)name.size()); detail::put16(out, 0);
      out.insert(out.end(), name.begin(), name.end());
      out.insert(out.end(), d.begin(), d.end());
      detail::put32(cd, 0x02014b50u); detail::put16(cd, 20); detail::put16(cd, 20); detail::put16(cd, 0); detail::put16(cd, 0);
      detail::put16(cd, 0); detail::put16(cd, 0x21); detail::put32(cd, crc);
      detail::put32(cd, (uint32_t)d.size()); detail::put32(cd, (uint32_t)d.size());
      detail::put16(cd, (uint16_t)name.size()); detail::put16(cd, 0); detail::put16(cd, 0); detail::put16(cd, 0); detail::put16(cd, 0);
      detail::put32(cd, 0); detail::put32(cd, off);
      cd.insert(cd.end(), name.begin(), name.end());
    }
    const uint32_t cd_off = (uint32_t)out.size();
    out.insert(out.end(), cd.begin(), cd.end());
    detail::put32(out, 0x06054b50u); detail::put16(out, 0); detail::put16(out, 0);
    detail::put16(out, (uint16_t)members_.size()); detail::put16(out, (uint16_t)members_.size());
    detail::put32(out, (uint32_t)cd.size()); detail::put32(out, cd_off); detail::put16(out, 0);
    std::ofstream f(path.c_str(), std::ios::binary);
    if (!f) throw std::runtime_error("npz: cannot write " + path);
    f.write(reinterpret_cast<const char *>(out.data()), (std::streamsize)out.size());
  }
 private:
  std::vector<std::pair<std::string, std::vector<unsigned char> > > members_;
};

}  // namespace npz
}  // namespace autorally_control
#endif

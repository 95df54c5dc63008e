// param_getter.h -- parameter front end (API of PI/param_getter.h:50-88, SRC/param_getter.cpp:75-151):
// fileExists() and loadParams(map*, launch_file), which reads the <param name= type= value=/> entries
// of the FIRST <node> of a roslaunch file into std::map<std::string, XmlRpc::XmlRpcValue>, expanding one
// $(env X) per value.  No ROS, Boost or XML library is needed: launch files are flat attribute lists.
// Deviation (documented): a <param> without a type attribute gets roslaunch's own inference
// (bool / int / double / str) instead of inheriting the previous entry's type.
#ifndef PARAM_GETTER_H_
#define PARAM_GETTER_H_
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <string>

#if __has_include(<xmlrpcpp/XmlRpcValue.h>)
#include <xmlrpcpp/XmlRpcValue.h>
#else
#include <XmlRpc/XmlRpcValue.h>  // include/compat
#endif

namespace autorally_control {

inline bool fileExists(const std::string &name) { return access(name.c_str(), F_OK) != -1; }

namespace param_detail {
inline bool attr(const std::string &tag, const std::string &key, std::string *out) {
  size_t p = 0;
  while ((p = tag.find(key, p)) != std::string::npos) {
    const bool word_start = p == 0 || tag[p - 1] == ' ' || tag[p - 1] == '\t' || tag[p - 1] == '\n';
    size_t q = p + key.size();
    while (q < tag.size() && (tag[q] == ' ' || tag[q] == '\t')) q++;
    if (word_start && q < tag.size() && tag[q] == '=') {
      q++;
      while (q < tag.size() && (tag[q] == ' ' || tag[q] == '\t')) q++;
      if (q < tag.size() && (tag[q] == '"' || tag[q] == '\'')) {
        const size_t e = tag.find(tag[q], q + 1);
        if (e == std::string::npos) return false;
        *out = tag.substr(q + 1, e - q - 1);
        return true;
      }
    }
    p += key.size();
  }
  return false;
}
inline XmlRpc::XmlRpcValue typed(const std::string &type, const std::string &v) {
  if (type == "int") return XmlRpc::XmlRpcValue(std::stoi(v));
  if (type == "double") return XmlRpc::XmlRpcValue(std::stod(v));
  if (type == "bool") return XmlRpc::XmlRpcValue(v == "true");
  if (type == "str") return XmlRpc::XmlRpcValue(v);
  // no type attribute: infer like roslaunch
  if (v == "true" || v == "false") return XmlRpc::XmlRpcValue(v == "true");
  char *end = nullptr;
  const long li = std::strtol(v.c_str(), &end, 10);
  if (!v.empty() && *end == '\0') return XmlRpc::XmlRpcValue((int)li);
  const double d = std::strtod(v.c_str(), &end);
  if (!v.empty() && *end == '\0') return XmlRpc::XmlRpcValue(d);
  return XmlRpc::XmlRpcValue(v);
}
}  // namespace param_detail

inline void loadParams(std::map<std::string, XmlRpc::XmlRpcValue> *params, const std::string &file_path) {
  if (!fileExists(file_path)) {
    fprintf(stderr, "Could not load roslaunch file containing mppi controller params at path: %s\n", file_path.c_str());
    return;
  }
  std::ifstream f(file_path.c_str());
  std::stringstream ss;
  ss << f.rdbuf();
  std::string xml = ss.str();
  // strip comments
  for (size_t a; (a = xml.find("<!--")) != std::string::npos;) {
    const size_t b = xml.find("-->", a);
    xml.erase(a, b == std::string::npos ? std::string::npos : b + 3 - a);
  }
  const size_t node = xml.find("<node");
  if (node == std::string::npos) return;
  const size_t node_end = xml.find("</node>", node);
  size_t p = xml.find('>', node);
  while (p != std::string::npos) {
    const size_t a = xml.find("<param", p);
    if (a == std::string::npos || (node_end != std::string::npos && a > node_end)) break;
    const size_t b = xml.find('>', a);
    if (b == std::string::npos) break;
    const std::string tag = xml.substr(a + 6, b - a - 6);
    std::string name, type, value;
    param_detail::attr(tag, "name", &name);
    param_detail::attr(tag, "type", &type);
    if (param_detail::attr(tag, "value", &value)) {
      const size_t e0 = value.find("$(env");
      if (e0 != std::string::npos) {
        const size_t e1 = value.find(')', e0);
        const std::string var = value.substr(e0 + 6, e1 - e0 - 6);
        const char *ev = std::getenv(var.c_str());
        value = value.substr(0, e0) + (ev ? ev : "") + value.substr(e1 + 1);
      }
      if (!name.empty() && params->find(name) == params->end()) (*params)[name] = param_detail::typed(type, value);
    }
    p = b;
  }
}

}  // namespace autorally_control
#endif

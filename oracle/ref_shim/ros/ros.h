// Stand-in for <ros/ros.h>, used ONLY to compile the reference's own MPPI sources (oracle/refbuild.py).
// TEST INFRASTRUCTURE.  Logging macros print to stderr; NodeHandle has no parameters.
#ifndef REF_SHIM_ROS_H_
#define REF_SHIM_ROS_H_
#include <cstdio>
#include <map>
#include <string>
#include "../../../include/compat/XmlRpc/XmlRpcValue.h"

#define ROS_FATAL(...) do { fprintf(stderr, "[FATAL] " __VA_ARGS__); fprintf(stderr, "\n"); } while (0)
#define ROS_ERROR(...) do { fprintf(stderr, "[ERROR] " __VA_ARGS__); fprintf(stderr, "\n"); } while (0)
#define ROS_WARN(...) do { fprintf(stderr, "[WARN] " __VA_ARGS__); fprintf(stderr, "\n"); } while (0)
#define ROS_INFO(...) do { } while (0)

namespace ros {
class NodeHandle {
 public:
  bool searchParam(const std::string &, std::string &) const { return false; }
  template <class T> bool getParam(const std::string &, T &) const { return false; }
  std::string getNamespace() const { return "/"; }
};
}  // namespace ros
#endif

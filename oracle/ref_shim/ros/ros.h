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
#define ROS_WARN_STREAM(x) do { } while (0)
#define ROS_INFO_STREAM(x) do { } while (0)

namespace ros {
class NodeHandle {
 public:
  bool searchParam(const std::string &, std::string &) const { return false; }
  template <class T> bool getParam(const std::string &, T &) const { return false; }
  std::string getNamespace() const { return "/"; }
};
// what PI/autorally_plant.h and PI/run_control_loop.cuh name: time stamps (seconds), and handles that do nothing
class Duration {
 public:
  Duration() : s_(0) {}
  explicit Duration(double s) : s_(s) {}
  double toSec() const { return s_; }
 private:
  double s_;
};
class Time {
 public:
  Time() : s_(0) {}
  explicit Time(double s) : s_(s) {}
  double toSec() const { return s_; }
  bool operator==(const Time &o) const { return s_ == o.s_; }
  bool operator!=(const Time &o) const { return s_ != o.s_; }
  Duration operator-(const Time &o) const { return Duration(s_ - o.s_); }
 private:
  double s_;
};
struct TimerEvent {};
class Publisher {};
class Subscriber {};
class Timer {};
inline void shutdown() {}
}  // namespace ros
#endif

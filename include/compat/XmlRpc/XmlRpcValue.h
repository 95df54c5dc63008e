// Minimal stand-in for XmlRpc::XmlRpcValue (ROS): the typed value the reference's parameter map
// holds (std::map<std::string, XmlRpc::XmlRpcValue>, PI/param_getter.h:81).  Supports the
// conversions the MPPI code performs: (bool), (int), (double), (std::string).
#ifndef MPPI_COMPAT_XMLRPCVALUE_H_
#define MPPI_COMPAT_XMLRPCVALUE_H_
#include <stdexcept>
#include <string>

namespace XmlRpc {

class XmlRpcException : public std::runtime_error {
 public:
  explicit XmlRpcException(const std::string &m) : std::runtime_error(m) {}
};

class XmlRpcValue {
 public:
  enum Type { TypeInvalid, TypeBoolean, TypeInt, TypeDouble, TypeString };
  XmlRpcValue() : type_(TypeInvalid), b_(false), i_(0), d_(0) {}
  XmlRpcValue(bool v) : type_(TypeBoolean), b_(v), i_(0), d_(0) {}
  XmlRpcValue(int v) : type_(TypeInt), b_(false), i_(v), d_(0) {}
  XmlRpcValue(double v) : type_(TypeDouble), b_(false), i_(0), d_(v) {}
  XmlRpcValue(const std::string &v) : type_(TypeString), b_(false), i_(0), d_(0), s_(v) {}
  XmlRpcValue(const char *v) : type_(TypeString), b_(false), i_(0), d_(0), s_(v) {}
  Type getType() const { return type_; }
  bool valid() const { return type_ != TypeInvalid; }
  operator bool &() { need(TypeBoolean); return b_; }
  operator int &() { need(TypeInt); return i_; }
  operator double &() { need(TypeDouble); return d_; }
  operator std::string &() { need(TypeString); return s_; }

 private:
  void need(Type t) const {
    if (type_ != t) throw XmlRpcException("type error");
  }
  Type type_;
  bool b_;
  int i_;
  double d_;
  std::string s_;
};

}  // namespace XmlRpc
#endif

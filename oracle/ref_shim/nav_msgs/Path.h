// Stand-in for the ROS message header (TEST INFRASTRUCTURE, oracle/refbuild.py).
#ifndef REF_SHIM_NAV_MSGS_Path_H_
#define REF_SHIM_NAV_MSGS_Path_H_
namespace nav_msgs { struct Path {}; }
#endif

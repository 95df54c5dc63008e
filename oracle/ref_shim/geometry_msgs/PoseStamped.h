// Stand-in for the ROS message header (TEST INFRASTRUCTURE, oracle/refbuild.py).
#ifndef REF_SHIM_GEOMETRY_MSGS_PoseStamped_H_
#define REF_SHIM_GEOMETRY_MSGS_PoseStamped_H_
namespace geometry_msgs { struct PoseStamped {}; }
#endif

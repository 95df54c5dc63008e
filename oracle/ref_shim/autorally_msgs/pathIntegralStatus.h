// Stand-in for the generated ROS message header (TEST INFRASTRUCTURE, oracle/refbuild.py): PI/autorally_plant.h only needs the type.
#ifndef REF_SHIM_AUTORALLY_MSGS_pathIntegralStatus_H_
#define REF_SHIM_AUTORALLY_MSGS_pathIntegralStatus_H_
namespace autorally_msgs { struct pathIntegralStatus {}; }
#endif

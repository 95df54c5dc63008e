// Stand-in for <dynamic_reconfigure/server.h> (TEST INFRASTRUCTURE, oracle/refbuild.py).
#ifndef REF_SHIM_DYNAMIC_RECONFIGURE_SERVER_H_
#define REF_SHIM_DYNAMIC_RECONFIGURE_SERVER_H_
namespace dynamic_reconfigure { template <class C> class Server {}; }
#endif

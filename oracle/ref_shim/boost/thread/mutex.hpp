// Stand-in for <boost/thread/mutex.hpp> (TEST INFRASTRUCTURE, oracle/refbuild.py).
#ifndef REF_SHIM_BOOST_MUTEX_HPP_
#define REF_SHIM_BOOST_MUTEX_HPP_
namespace boost { class mutex { public: void lock() {} void unlock() {} }; }
#endif

// Stand-in for <boost/thread/thread.hpp> (TEST INFRASTRUCTURE, oracle/refbuild.py).
#ifndef REF_SHIM_BOOST_THREAD_HPP_
#define REF_SHIM_BOOST_THREAD_HPP_
#include <thread>
namespace boost { typedef std::thread thread; }
#endif

// Stand-in for <boost/typeof/typeof.hpp> (TEST INFRASTRUCTURE, oracle/refbuild.py).  DDP/ddp_model_wrapper.h:13 spells the GNU
// keyword `typeof`, which g++ only accepts in gnu++ modes; nvcc compiles the host side as strict c++14.
#ifndef REF_SHIM_BOOST_TYPEOF_HPP_
#define REF_SHIM_BOOST_TYPEOF_HPP_
#define typeof __typeof__
#endif

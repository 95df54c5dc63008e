// Stand-in for <boost/type_traits.hpp> (TEST INFRASTRUCTURE, oracle/refbuild.py): the reference uses <type_traits> only.
#include <type_traits>

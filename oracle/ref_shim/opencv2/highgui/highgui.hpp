// Stand-in for OpenCV's highgui header (TEST INFRASTRUCTURE, oracle/refbuild.py): nothing is displayed.
#include "../core/core.hpp"

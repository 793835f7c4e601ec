// TEST INFRASTRUCTURE: the reference's benchmark drivers include "xcl2.hpp" (Xilinx host helpers) for
// aligned_allocator and the OpenCL types; both come from include/graphlily/global.h and opencl_compat.h.
#ifndef GLB_REF_COMPAT_XCL2_HPP_
#define GLB_REF_COMPAT_XCL2_HPP_
#include <cmath>
#include "graphlily/global.h"
#include "opencl_compat.h"
using std::floor;
#endif

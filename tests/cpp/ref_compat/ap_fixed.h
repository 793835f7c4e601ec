// TEST INFRASTRUCTURE: the reference's tests include <ap_fixed.h> (Vitis HLS); with val_t = float
// (include/graphlily/global.h) nothing of it is used.
#ifndef GLB_REF_COMPAT_AP_FIXED_H_
#define GLB_REF_COMPAT_AP_FIXED_H_
#endif

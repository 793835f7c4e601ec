// TEST INFRASTRUCTURE ONLY (oracle build shim) -- never included by the product.
//
// Stand-in for Xilinx Vitis-HLS "ap_fixed.h", which is not vendored in the
// reference tree and not installed here.  The reference's CPU path
// (compute_reference_results, /root/reference/graphlily/module/*.h) only ever
// converts semiring constants (0, 1, 255) through val_t and otherwise works on
// float containers, so a float-backed ap_ufixed is exact for the oracle.
#ifndef ORACLE_SHIM_AP_FIXED_H_
#define ORACLE_SHIM_AP_FIXED_H_

enum ap_q_mode { AP_RND, AP_RND_ZERO, AP_RND_MIN_INF, AP_RND_INF, AP_RND_CONV, AP_TRN, AP_TRN_ZERO };
enum ap_o_mode { AP_SAT, AP_SAT_ZERO, AP_SAT_SYM, AP_WRAP, AP_WRAP_SM };

template <int W, int I, ap_q_mode Q = AP_TRN, ap_o_mode O = AP_WRAP>
struct ap_ufixed {
    float v;
    ap_ufixed() : v(0.0f) {}
    template <typename T> ap_ufixed(T x) : v(static_cast<float>(x)) {}
    operator float() const { return v; }
    ap_ufixed operator++(int) { ap_ufixed old(*this); v += 1.0f; return old; }
    ap_ufixed &operator++() { v += 1.0f; return *this; }
};

template <int W, int I, ap_q_mode Q = AP_TRN, ap_o_mode O = AP_WRAP>
using ap_fixed = ap_ufixed<W, I, Q, O>;

template <int W> struct ap_uint {
    unsigned long long v;
    ap_uint() : v(0) {}
    template <typename T> ap_uint(T x) : v(static_cast<unsigned long long>(x)) {}
    operator unsigned long long() const { return v; }
};

#endif  // ORACLE_SHIM_AP_FIXED_H_

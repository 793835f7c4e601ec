// Value types and semiring functors shared by the SpMV, SpMSpV and apply kernels.
//
// The reference's processing elements apply (x) then a read-modify-write (+)
// (/root/reference/graphlily/hw/ufixed_pe_fwd.h:23-65) on `val_t`, which its global.h:60-64 offers in
// three flavours.  All three are 32-bit words, so vectors and matrix values keep one layout and the
// kernels carry them in `float` registers as BIT CONTAINERS (moves, shuffles, loads and stores preserve
// the bits); only the functions below interpret them:
//
//   GLB_VAL_F32     float -- the type of the reference's CPU path (compute_reference_results,
//                   spmv_module.h:488-510): the parity target, and the default of every entry point
//   GLB_VAL_U32     unsigned: C arithmetic modulo 2^32, infinity UINT_INF = 0xffffffff (global.h:62,78)
//   GLB_VAL_UFIXED  ap_ufixed<32, 8, AP_RND, AP_SAT>, the type of the shipped bitstream (global.h:63):
//                   Q8.24, value = word / 2^24 in [0, 256); a product is the exact 64-bit product
//                   rounded to 24 fraction bits (AP_RND: add half an ulp, truncate) and saturated at
//                   0xffffffff (AP_SAT); a sum saturates; infinity UFIXED_INF = 255 (global.h:79)
//
//   kMulAdd        y += a * b
//   kLogicalAndOr  y  = y || (a && b)      -> 0 / 1 (1.0f, 1, 1 << 24)
//   kAddMin        y  = min(y, a + b)
// Kernels accumulate from the (+)-identity and fold the caller's runtime `zero`
// (SemiringType::zero, global.h:90-100) in once per row with with_zero().  The integer types are exact
// and saturating / modular addition of non-negative words is associative, so their results do not
// depend on the reduction order: bit-exact against a sequential model (oracle/valtype_model.h).
#ifndef GLB_SEMIRING_CUH_
#define GLB_SEMIRING_CUH_

#include <cuda_runtime.h>
#include <math_constants.h>

#include "graphlily_b200.h"

template <int VT> struct Val;

template <> struct Val<GLB_VAL_F32> {
    static __host__ __device__ __forceinline__ float zero() { return 0.0f; }
    static __host__ __device__ __forceinline__ float one() { return 1.0f; }
    static __device__ __forceinline__ float inf() { return CUDART_INF_F; }
    static __device__ __forceinline__ bool is_zero(float v) { return v == 0.0f; }   // -0.0f is zero, NaN is not
    static __device__ __forceinline__ bool equal(float a, float b) { return a == b; }
    static __device__ __forceinline__ bool less(float a, float b) { return a < b; }
    // separate multiply and add roundings, like the un-contracted host loop
    static __device__ __forceinline__ float times(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float plus(float a, float b) { return __fadd_rn(a, b); }
};

template <> struct Val<GLB_VAL_U32> {
    static __host__ __device__ __forceinline__ float zero() { return 0.0f; }
    static __device__ __forceinline__ float one() { return __uint_as_float(1u); }
    static __device__ __forceinline__ float inf() { return __uint_as_float(0xffffffffu); }
    static __device__ __forceinline__ bool is_zero(float v) { return __float_as_uint(v) == 0u; }
    static __device__ __forceinline__ bool equal(float a, float b) { return __float_as_uint(a) == __float_as_uint(b); }
    static __device__ __forceinline__ bool less(float a, float b) { return __float_as_uint(a) < __float_as_uint(b); }
    static __device__ __forceinline__ float times(float a, float b) { return __uint_as_float(__float_as_uint(a) * __float_as_uint(b)); }
    static __device__ __forceinline__ float plus(float a, float b) { return __uint_as_float(__float_as_uint(a) + __float_as_uint(b)); }
};

template <> struct Val<GLB_VAL_UFIXED> {
    static __host__ __device__ __forceinline__ float zero() { return 0.0f; }
    static __device__ __forceinline__ float one() { return __uint_as_float(1u << 24); }
    static __device__ __forceinline__ float inf() { return __uint_as_float(0xffffffffu); }
    static __device__ __forceinline__ bool is_zero(float v) { return __float_as_uint(v) == 0u; }
    static __device__ __forceinline__ bool equal(float a, float b) { return __float_as_uint(a) == __float_as_uint(b); }
    static __device__ __forceinline__ bool less(float a, float b) { return __float_as_uint(a) < __float_as_uint(b); }
    static __device__ __forceinline__ float times(float a, float b) {
        const unsigned long long p = (unsigned long long)__float_as_uint(a) * __float_as_uint(b) + (1ull << 23);  // AP_RND
        const unsigned long long q = p >> 24;
        return __uint_as_float(q > 0xffffffffull ? 0xffffffffu : unsigned(q));                                  // AP_SAT
    }
    static __device__ __forceinline__ float plus(float a, float b) {
        const unsigned s = __float_as_uint(a) + __float_as_uint(b);
        return __uint_as_float(s < __float_as_uint(a) ? 0xffffffffu : s);                                       // AP_SAT
    }
};

template <int OP, int VT = GLB_VAL_F32> struct Semi;

template <int VT> struct Semi<GLB_OP_MUL_ADD, VT> {
    static __device__ __forceinline__ float ident() { return Val<VT>::zero(); }
    static __device__ __forceinline__ float mul(float a, float b) { return Val<VT>::times(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return Val<VT>::plus(a, b); }
    static __device__ __forceinline__ float with_zero(float zero, float t) { return Val<VT>::plus(zero, t); }
};

template <int VT> struct Semi<GLB_OP_LOGICAL_AND_OR, VT> {
    static __device__ __forceinline__ float ident() { return Val<VT>::zero(); }
    static __device__ __forceinline__ float mul(float a, float b) {
        return (!Val<VT>::is_zero(a) && !Val<VT>::is_zero(b)) ? Val<VT>::one() : Val<VT>::zero();
    }
    // operands are exactly 0 or 1 of the value type: OR of the bit patterns
    static __device__ __forceinline__ float add(float a, float b) {
        return __int_as_float(__float_as_int(a) | __float_as_int(b));
    }
    static __device__ __forceinline__ float with_zero(float zero, float t) {
        return (!Val<VT>::is_zero(zero) || !Val<VT>::is_zero(t)) ? Val<VT>::one() : Val<VT>::zero();
    }
};

template <int VT> struct Semi<GLB_OP_ADD_MIN, VT> {
    static __device__ __forceinline__ float ident() { return Val<VT>::inf(); }
    static __device__ __forceinline__ float mul(float a, float b) { return Val<VT>::plus(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return Val<VT>::less(b, a) ? b : a; }
    static __device__ __forceinline__ float with_zero(float zero, float t) { return Val<VT>::less(t, zero) ? t : zero; }
};

#endif  // GLB_SEMIRING_CUH_

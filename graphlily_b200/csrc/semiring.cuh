// Semiring functors shared by the SpMV and SpMSpV kernels.
//
// The reference's processing elements apply (x) then a read-modify-write (+)
// (/root/reference/graphlily/hw/ufixed_pe_fwd.h:23-65); the oracle semantics are the fp32
// ones of SpMVModule::compute_reference_results (spmv_module.h:488-510):
//   kMulAdd        y += a * b
//   kLogicalAndOr  y  = y || (a && b)      -> 0.0f / 1.0f
//   kAddMin        y  = min(y, a + b)
// Kernels accumulate from the (+)-identity and fold the caller's runtime `zero`
// (SemiringType::zero, global.h:90-100) in once per row with with_zero().
#ifndef GLB_SEMIRING_CUH_
#define GLB_SEMIRING_CUH_

#include <cuda_runtime.h>
#include <math_constants.h>

#include "graphlily_b200.h"

template <int OP> struct Semi;

template <> struct Semi<GLB_OP_MUL_ADD> {
    static __device__ __forceinline__ float ident() { return 0.0f; }
    // separate multiply and add roundings, like the un-contracted host loop
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float with_zero(float zero, float t) { return __fadd_rn(zero, t); }
};

template <> struct Semi<GLB_OP_LOGICAL_AND_OR> {
    static __device__ __forceinline__ float ident() { return 0.0f; }
    static __device__ __forceinline__ float mul(float a, float b) { return (a != 0.0f && b != 0.0f) ? 1.0f : 0.0f; }
    // operands are exactly 0.0f or 1.0f: OR of the bit patterns
    static __device__ __forceinline__ float add(float a, float b) {
        return __int_as_float(__float_as_int(a) | __float_as_int(b));
    }
    static __device__ __forceinline__ float with_zero(float zero, float t) {
        return (zero != 0.0f || t != 0.0f) ? 1.0f : 0.0f;
    }
};

template <> struct Semi<GLB_OP_ADD_MIN> {
    static __device__ __forceinline__ float ident() { return CUDART_INF_F; }
    static __device__ __forceinline__ float mul(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return (b < a) ? b : a; }
    static __device__ __forceinline__ float with_zero(float zero, float t) { return (t < zero) ? t : zero; }
};

#endif  // GLB_SEMIRING_CUH_

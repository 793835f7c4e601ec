// graphlily-b200: shared types and constants of the host API.
//
// Mirrors /root/reference/graphlily/global.h in meaning: val_t / idx_t / idx_val_t (:62-70), the
// aligned host vector types (:72-76), the infinities (:78-80), OperationType / SemiringType and the
// three semirings (:83-100), MaskType (:103-107), pack_size / num_hbm_channels (:57-59, kept because
// the apps pad matrix dimensions to their product) and convert_sparse_vec_to_dense_vec (:153-164).
// The FPGA-only parts (find_device, HBM[] bank ids, makefile strings, :27-54,110-143) have no
// counterpart: the kernels live in libgraphlily_b200.so (include/graphlily_b200.h).
//
// val_t is float: parity is defined against the reference's fp32 compute_reference_results path
// (its shipped bitstream uses ap_ufixed<32,8>; global.h:64 lists float as the alternative).
#ifndef GRAPHLILY_GLOBAL_H_
#define GRAPHLILY_GLOBAL_H_

#include <algorithm>
#include <cassert>
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <new>
#include <string>
#include <vector>

#include "graphlily_b200.h"

// Page-aligned host allocator with the interface of xcl2's aligned_allocator
// (/root/reference/xrt/includes/xcl2/xcl2.hpp:61-76).
template <typename T>
struct aligned_allocator {
    using value_type = T;
    aligned_allocator() {}
    aligned_allocator(const aligned_allocator &) {}
    template <typename U> aligned_allocator(const aligned_allocator<U> &) {}
    T *allocate(std::size_t num) {
        void *ptr = nullptr;
        const std::size_t bytes = num * sizeof(T);
        if (posix_memalign(&ptr, 4096, bytes != 0 ? bytes : 4096)) throw std::bad_alloc();
        return reinterpret_cast<T *>(ptr);
    }
    void deallocate(T *p, std::size_t) { free(p); }
};
template <typename T, typename U>
bool operator==(const aligned_allocator<T> &, const aligned_allocator<U> &) { return true; }
template <typename T, typename U>
bool operator!=(const aligned_allocator<T> &, const aligned_allocator<U> &) { return false; }

namespace graphlily {

inline std::string get_root_path() {
    const char *root_path = getenv("GRAPHLILY_ROOT_PATH");
    return root_path == nullptr ? std::string("") : std::string(root_path);
}
const std::string root_path = get_root_path();

// Dimension padding unit of the apps (rows / cols rounded to num_channels * pack_size).
const uint32_t pack_size = 8;
const uint32_t num_hbm_channels = 16;

using val_t = float;
typedef uint32_t idx_t;
typedef struct { idx_t index; val_t val; } idx_val_t;    // same layout as glb_idx_val_t
typedef struct { idx_t index; float val; } idx_float_t;
static_assert(sizeof(idx_val_t) == sizeof(glb_idx_val_t), "idx_val_t must match the C ABI");

using aligned_dense_vec_t = std::vector<val_t, aligned_allocator<val_t>>;
using aligned_sparse_vec_t = std::vector<idx_val_t, aligned_allocator<idx_val_t>>;
using aligned_dense_float_vec_t = std::vector<float, aligned_allocator<float>>;
using aligned_sparse_float_vec_t = std::vector<idx_float_t, aligned_allocator<idx_float_t>>;

const val_t UINT_INF = 0xffffffff;   // global.h:78 (the infinity of the unsigned val_t build)
const val_t UFIXED_INF = 255;        // the shipped TropicalSemiring zero (global.h:79,99)
const val_t FLOAT_INF = 999999999;   // global.h:80

// Operation type, named as k<opx><op+>
enum OperationType {
    kMulAdd = GLB_OP_MUL_ADD,
    kLogicalAndOr = GLB_OP_LOGICAL_AND_OR,
    kAddMin = GLB_OP_ADD_MIN,
};

struct SemiringType {
    OperationType op;
    val_t one;   // identity element of <x>
    val_t zero;  // identity element of <+>
};

const SemiringType ArithmeticSemiring = {kMulAdd, 1, 0};
const SemiringType LogicalSemiring = {kLogicalAndOr, 1, 0};
const SemiringType TropicalSemiring = {kAddMin, 0, UFIXED_INF};

enum MaskType {
    kNoMask = GLB_MASK_NONE,
    kMaskWriteToZero = GLB_MASK_WRITE_TO_ZERO,
    kMaskWriteToOne = GLB_MASK_WRITE_TO_ONE,
};

// Sparse vector convention: element 0 is {nnz, -}, elements 1..nnz are {index, val}.
template <typename sparse_vec_t, typename dense_vec_t, typename value_t>
dense_vec_t convert_sparse_vec_to_dense_vec(const sparse_vec_t &sparse_vector, uint32_t range, value_t zero) {
    dense_vec_t dense_vector(range);
    std::fill(dense_vector.begin(), dense_vector.end(), zero);
    const uint32_t nnz = sparse_vector[0].index;
    for (uint32_t i = 1; i <= nnz; i++) dense_vector[sparse_vector[i].index] = sparse_vector[i].val;
    return dense_vector;
}

}  // namespace graphlily

#endif  // GRAPHLILY_GLOBAL_H_

// graphlily-b200: shared types and constants of the host API.
//
// Mirrors /root/reference/graphlily/global.h in meaning: val_t / idx_t / idx_val_t (:62-70), the
// aligned host vector types (:72-76), the infinities (:78-80), OperationType / SemiringType and the
// three semirings (:83-100), MaskType (:103-107), pack_size / num_hbm_channels (:57-59, kept because
// the apps pad matrix dimensions to their product) and convert_sparse_vec_to_dense_vec (:153-164).
// The FPGA-only parts (find_device, HBM[] bank ids, makefile strings, :27-54,110-143) have no
// counterpart: the kernels live in libgraphlily_b200.so (include/graphlily_b200.h).
//
// val_t is chosen at compile time like the reference's (global.h:60-64, by editing the typedef there):
//   default                      float -- the type of the reference's CPU path (compute_reference_results): the parity target
//   -DGRAPHLILY_VAL_T_UNSIGNED   unsigned (UINT_INF, arithmetic modulo 2^32)
//   -DGRAPHLILY_VAL_T_UFIXED     ap_ufixed<32, 8, AP_RND, AP_SAT> of the shipped bitstream, as graphlily::ufixed_32_8
// The module classes take any of the three as their data type (val_traits) and call the matching kernels (glb_*_vt).
#ifndef GRAPHLILY_GLOBAL_H_
#define GRAPHLILY_GLOBAL_H_

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
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

// Host-side value class of ap_ufixed<32, 8, AP_RND, AP_SAT>: one 32-bit word, value = word / 2^24 in [0, 256).
// Construction quantises like the HLS type (round half up to 24 fraction bits, saturate at both ends); the arithmetic
// itself happens on the device (semiring.cuh) -- this class only builds inputs and reads results.
class ufixed_32_8 {
public:
    ufixed_32_8() : word_(0) {}
    ufixed_32_8(double v) : word_(quantise(v)) {}
    ufixed_32_8(float v) : word_(quantise(v)) {}
    ufixed_32_8(int v) : word_(quantise(v)) {}
    ufixed_32_8(unsigned v) : word_(quantise(v)) {}
    static ufixed_32_8 from_word(uint32_t w) { ufixed_32_8 r; r.word_ = w; return r; }
    uint32_t word() const { return word_; }
    double to_double() const { return double(word_) / 16777216.0; }
    operator float() const { return float(to_double()); }
    bool operator==(const ufixed_32_8 &o) const { return word_ == o.word_; }
    bool operator!=(const ufixed_32_8 &o) const { return word_ != o.word_; }
    bool operator<(const ufixed_32_8 &o) const { return word_ < o.word_; }
private:
    static uint32_t quantise(double v) {
        if (!(v > 0.0)) return 0u;                       // negative values and NaN saturate at 0
        const double scaled = std::floor(v * 16777216.0 + 0.5);
        return scaled >= 4294967295.0 ? 0xffffffffu : uint32_t(scaled);
    }
    uint32_t word_;
};
static_assert(sizeof(ufixed_32_8) == 4, "ufixed_32_8 must be one 32-bit word");

// What the modules need to know about a value type: its id in the C ABI, its word, the conversion the reference's
// formatters apply to the float matrix ((val_t)x, data_formatter.h).
template <typename T> struct val_traits;
template <> struct val_traits<float> {
    static constexpr int id = GLB_VAL_F32;
    static uint32_t bits(float v) { uint32_t b; std::memcpy(&b, &v, 4); return b; }
    static float from_float(float v) { return v; }
};
template <> struct val_traits<unsigned> {
    static constexpr int id = GLB_VAL_U32;
    static uint32_t bits(unsigned v) { return v; }
    static unsigned from_float(float v) { return v <= 0.0f ? 0u : v >= 4294967295.0f ? 0xffffffffu : unsigned(v); }
};
template <> struct val_traits<ufixed_32_8> {
    static constexpr int id = GLB_VAL_UFIXED;
    static uint32_t bits(ufixed_32_8 v) { return v.word(); }
    static ufixed_32_8 from_float(float v) { return ufixed_32_8(v); }
};
// a value as the 32-bit container the C ABI passes scalars in (float parameters carry the word's bits)
template <typename T> inline float val_container(T v) {
    const uint32_t b = val_traits<T>::bits(v);
    float f;
    std::memcpy(&f, &b, 4);
    return f;
}

#if defined(GRAPHLILY_VAL_T_UNSIGNED)
using val_t = unsigned;
#elif defined(GRAPHLILY_VAL_T_UFIXED)
using val_t = ufixed_32_8;
#else
using val_t = float;
#endif
typedef uint32_t idx_t;
typedef struct { idx_t index; val_t val; } idx_val_t;    // same layout as glb_idx_val_t
typedef struct { idx_t index; float val; } idx_float_t;
static_assert(sizeof(idx_val_t) == sizeof(glb_idx_val_t), "idx_val_t must match the C ABI");

using aligned_dense_vec_t = std::vector<val_t, aligned_allocator<val_t>>;
using aligned_sparse_vec_t = std::vector<idx_val_t, aligned_allocator<idx_val_t>>;
using aligned_dense_float_vec_t = std::vector<float, aligned_allocator<float>>;
using aligned_sparse_float_vec_t = std::vector<idx_float_t, aligned_allocator<idx_float_t>>;

const val_t UINT_INF = val_t(0xffffffffu);   // global.h:78 (the infinity of the unsigned val_t build)
const val_t UFIXED_INF = val_t(255);         // the shipped TropicalSemiring zero (global.h:79,99)
const val_t FLOAT_INF = val_t(999999999);    // global.h:80

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

const SemiringType ArithmeticSemiring = {kMulAdd, val_t(1), val_t(0)};
const SemiringType LogicalSemiring = {kLogicalAndOr, val_t(1), val_t(0)};
#if defined(GRAPHLILY_VAL_T_UNSIGNED)
const SemiringType TropicalSemiring = {kAddMin, val_t(0), UINT_INF};     // global.h:98
#else
const SemiringType TropicalSemiring = {kAddMin, val_t(0), UFIXED_INF};   // global.h:99 (the shipped choice)
#endif

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

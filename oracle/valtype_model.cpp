// TEST INFRASTRUCTURE ONLY -- never linked by the product.
//
// Sequential model of the operators on the two non-float `val_t` choices of the reference
// (/root/reference/graphlily/global.h:60-64): `unsigned` and ap_ufixed<32, 8, AP_RND, AP_SAT>.  The loop
// structure is that of compute_reference_results (spmv_module.h:478-532, spmspv_module.h:446-520,
// add_scalar_vector_dense_module.h:195-204, assign_vector_dense_module.h:223-246,
// assign_vector_sparse_module.h:306-335); the arithmetic is that of the processing elements
// (graphlily/hw/ufixed_pe_fwd.h:23-65: `a * b`, `a && b`, `a + b` / `a + b`, `a || b`, MIN(a, b)) evaluated
// on val_t itself -- for the fixed-point type through the software ap_ufixed of oracle/shim/ap_fixed.h.
// PARITY UNPINNED: the reference ships nothing (bitstream, emulator, golden vector) that pins its device
// numerics; this model follows the documented ap_ufixed semantics.  Words travel as uint32.
#include <cstdint>
#include <cstring>
#include <vector>

#include "shim/ap_fixed.h"

namespace {

typedef ap_ufixed<32, 8, AP_RND, AP_SAT> ufixed_t;

struct U32 {   // `using val_t = unsigned;`
    uint32_t w;
    static U32 from_word(uint32_t x) { U32 v; v.w = x; return v; }
    uint32_t word() const { return w; }
    static U32 one() { return from_word(1u); }
    static U32 inf() { return from_word(0xffffffffu); }
    friend U32 operator*(U32 a, U32 b) { return from_word(a.w * b.w); }
    friend U32 operator+(U32 a, U32 b) { return from_word(a.w + b.w); }
    friend bool operator<(U32 a, U32 b) { return a.w < b.w; }
    bool nonzero() const { return w != 0; }
};

struct UFX {   // `using val_t = ap_ufixed<32, 8, AP_RND, AP_SAT>;`
    ufixed_t v;
    static UFX from_word(uint32_t x) { UFX r; r.v = ufixed_t::from_raw(x); return r; }
    uint32_t word() const { return uint32_t(v.raw); }
    static UFX one() { UFX r; r.v = ufixed_t(1); return r; }
    static UFX inf() { return from_word(0xffffffffu); }
    friend UFX operator*(UFX a, UFX b) { UFX r; r.v = a.v * b.v; return r; }
    friend UFX operator+(UFX a, UFX b) { UFX r; r.v = a.v + b.v; return r; }
    friend bool operator<(UFX a, UFX b) { return a.v < b.v; }
    bool nonzero() const { return v.raw != 0; }
};

enum { MUL_ADD = 0, AND_OR = 1, ADD_MIN = 2 };

template <typename V> V semi_mul(int op, V a, V b) {   // pe_ufixed_mul_alu
    switch (op) {
        case MUL_ADD: return a * b;
        case AND_OR: return (a.nonzero() && b.nonzero()) ? V::one() : V::from_word(0);
        default: return a + b;
    }
}
template <typename V> V semi_add(int op, V a, V b) {   // pe_ufixed_add_alu
    switch (op) {
        case MUL_ADD: return a + b;
        case AND_OR: return (a.nonzero() || b.nonzero()) ? V::one() : V::from_word(0);
        default: return b < a ? b : a;
    }
}

template <typename V>
void spmv(uint32_t nrows, const uint32_t *indptr, const uint32_t *indices, const uint32_t *data, int op, uint32_t zero,
          int mask_type, const uint32_t *x, const uint32_t *mask, uint32_t *y) {
    for (uint32_t r = 0; r < nrows; r++) {
        V acc = V::from_word(zero);                       // spmv_module.h:491-492: results start at semiring.zero
        for (uint32_t i = indptr[r]; i < indptr[r + 1]; i++)
            acc = semi_add(op, acc, semi_mul(op, V::from_word(data[i]), V::from_word(x[indices[i]])));
        uint32_t w = acc.word();
        if (mask_type == 1 && mask[r] != 0) w = 0;        // kMaskWriteToZero: spmv_module.h:518-523 (literal 0)
        if (mask_type == 2 && mask[r] == 0) w = 0;        // kMaskWriteToOne
        y[r] = w;
    }
}

template <typename V>
void spmspv(uint32_t nrows, const uint32_t *indptr, const uint32_t *indices, const uint32_t *data, int op, uint32_t zero,
            int mask_type, const uint32_t *x_idx, const uint32_t *x_val, uint32_t x_nnz, const uint32_t *mask, uint32_t *y) {
    for (uint32_t r = 0; r < nrows; r++) y[r] = zero;     // spmspv_module.h:452-453
    for (uint32_t k = 0; k < x_nnz; k++) {
        const uint32_t c = x_idx[k];
        for (uint32_t i = indptr[c]; i < indptr[c + 1]; i++) {
            const uint32_t r = indices[i];
            y[r] = semi_add(op, V::from_word(y[r]), semi_mul(op, V::from_word(data[i]), V::from_word(x_val[k]))).word();
        }
    }
    for (uint32_t r = 0; r < nrows && mask_type != 0; r++) {   // spmspv_module.h:499-516: compared with `zero`
        const bool off = (mask_type == 2) ? (mask[r] == zero) : (mask[r] != zero);
        if (off) y[r] = zero;
    }
}

}  // namespace

extern "C" {

int vt_spmv(int val_type, uint32_t nrows, uint32_t ncols, const uint32_t *indptr, const uint32_t *indices, const uint32_t *data,
            int op, uint32_t zero, int mask_type, const uint32_t *x, const uint32_t *mask, uint32_t *y) {
    (void)ncols;
    if (val_type == 1) spmv<U32>(nrows, indptr, indices, data, op, zero, mask_type, x, mask, y);
    else if (val_type == 2) spmv<UFX>(nrows, indptr, indices, data, op, zero, mask_type, x, mask, y);
    else return 1;
    return 0;
}

int vt_spmspv(int val_type, uint32_t nrows, uint32_t ncols, const uint32_t *indptr, const uint32_t *indices, const uint32_t *data,
              int op, uint32_t zero, int mask_type, const uint32_t *x_idx, const uint32_t *x_val, uint32_t x_nnz,
              const uint32_t *mask, uint32_t *y) {
    (void)ncols;
    if (val_type == 1) spmspv<U32>(nrows, indptr, indices, data, op, zero, mask_type, x_idx, x_val, x_nnz, mask, y);
    else if (val_type == 2) spmspv<UFX>(nrows, indptr, indices, data, op, zero, mask_type, x_idx, x_val, x_nnz, mask, y);
    else return 1;
    return 0;
}

int vt_ewise_add(int val_type, const uint32_t *in, uint32_t *out, uint32_t len, uint32_t val) {
    for (uint32_t i = 0; i < len; i++)
        out[i] = val_type == 1 ? (U32::from_word(in[i]) + U32::from_word(val)).word()
                               : (UFX::from_word(in[i]) + UFX::from_word(val)).word();
    return 0;
}

// new_frontier (idx, val) pairs in list order; returns the count (assign_vector_sparse_module.h:318-335)
int vt_assign_sparse_relax(int val_type, const uint32_t *m_idx, const uint32_t *m_val, uint32_t nnz, uint32_t *inout,
                           uint32_t *nf_idx, uint32_t *nf_val) {
    (void)val_type;   // both types compare as unsigned words
    int cnt = 0;
    for (uint32_t i = 0; i < nnz; i++)
        if (inout[m_idx[i]] > m_val[i]) {
            inout[m_idx[i]] = m_val[i];
            nf_idx[cnt] = m_idx[i];
            nf_val[cnt++] = m_val[i];
        }
    return cnt;
}

// conversions of the software ap_ufixed<32, 8, AP_RND, AP_SAT> (for the known-answer tests)
uint32_t vt_ufixed_from_double(double x) { return uint32_t(ufixed_t(x).raw); }
double vt_ufixed_to_double(uint32_t w) { return ufixed_t::from_raw(w).to_double(); }
uint32_t vt_ufixed_mul(uint32_t a, uint32_t b) { return uint32_t((ufixed_t::from_raw(a) * ufixed_t::from_raw(b)).raw); }
uint32_t vt_ufixed_add(uint32_t a, uint32_t b) { return uint32_t((ufixed_t::from_raw(a) + ufixed_t::from_raw(b)).raw); }

}  // extern "C"

// TEST INFRASTRUCTURE ONLY (oracle build shim) -- never included by the product.
//
// Stand-in for Xilinx Vitis-HLS "ap_fixed.h", which is not vendored in the reference tree and not
// installed here.  Two models of ap_ufixed<W, I, Q, O>, selected at compile time:
//
//   default                      a software model of the documented semantics, NOT float-backed: the value
//                                is the integer word raw / 2^(W - I); conversions and assignments quantise
//                                with the Q mode (AP_RND: round half up, AP_TRN: truncate) and handle
//                                overflow with the O mode (AP_SAT: clamp to [0, 2^W - 1], AP_WRAP: modulo
//                                2^W); products and sums are formed exactly (128-bit) and then quantised to
//                                the operand type, which is what `ValT out = a * b;` / `a + b` do in the
//                                reference's processing elements (graphlily/hw/ufixed_pe_fwd.h:23-65).
//                                oracle/valtype_model.cpp evaluates the GLB_VAL_UFIXED variant with it.
//   -DGLB_AP_FIXED_FLOAT_BACKED  the float-backed stand-in of round 1.  The reference's CPU path
//                                (compute_reference_results, graphlily/module/*.h) only ever converts
//                                semiring constants (0, 1, 255) through val_t and otherwise works on float
//                                containers, so both models give the same oracle: oracle/Makefile builds
//                                the reference both ways and tests/test_oracle_vs_ref.py compares them.
// No artefact of the reference (bitstream, emulator, golden vector) pins these device numerics.
#ifndef ORACLE_SHIM_AP_FIXED_H_
#define ORACLE_SHIM_AP_FIXED_H_

#include <cmath>
#include <cstdint>

enum ap_q_mode { AP_RND, AP_RND_ZERO, AP_RND_MIN_INF, AP_RND_INF, AP_RND_CONV, AP_TRN, AP_TRN_ZERO };
enum ap_o_mode { AP_SAT, AP_SAT_ZERO, AP_SAT_SYM, AP_WRAP, AP_WRAP_SM };

#ifdef GLB_AP_FIXED_FLOAT_BACKED

template <int W, int I, ap_q_mode Q = AP_TRN, ap_o_mode O = AP_WRAP>
struct ap_ufixed {
    float v;
    ap_ufixed() : v(0.0f) {}
    template <typename T> ap_ufixed(T x) : v(static_cast<float>(x)) {}
    operator float() const { return v; }
    ap_ufixed operator++(int) { ap_ufixed old(*this); v += 1.0f; return old; }
    ap_ufixed &operator++() { v += 1.0f; return *this; }
};

#else

template <int W, int I, ap_q_mode Q = AP_TRN, ap_o_mode O = AP_WRAP>
struct ap_ufixed {
    static_assert(W >= 1 && W <= 63 && I <= W, "model covers words of up to 63 bits");
    static const int F = W - I;          // fraction bits
    unsigned long long raw;              // the word: value = raw / 2^F

    static unsigned long long max_raw() { return (1ull << W) - 1ull; }
    // overflow handling of an exact non-negative integer word (may exceed W bits)
    static unsigned long long overflow(unsigned __int128 w) {
        if (w <= (unsigned __int128)max_raw()) return (unsigned long long)w;
        return (O == AP_SAT || O == AP_SAT_ZERO || O == AP_SAT_SYM) ? (O == AP_SAT_ZERO ? 0ull : max_raw())
                                                                   : (unsigned long long)(w & (unsigned __int128)max_raw());
    }
    // drop `drop` low bits of an exact word with the quantisation mode
    static unsigned __int128 quantize(unsigned __int128 w, int drop) {
        if (drop <= 0) return w << (-drop);
        if (Q == AP_RND || Q == AP_RND_INF) w += (unsigned __int128)1 << (drop - 1);   // round half up (unsigned: towards +inf)
        else if (Q == AP_RND_CONV) {                                                      // round half to even
            const unsigned __int128 half = (unsigned __int128)1 << (drop - 1), mask = ((unsigned __int128)1 << drop) - 1;
            const unsigned __int128 rem = w & mask;
            if (rem > half || (rem == half && ((w >> drop) & 1))) w += (unsigned __int128)1 << drop;
            return w >> drop;
        }
        return w >> drop;                                                                  // AP_TRN: truncate
    }
    static ap_ufixed from_raw(unsigned long long r) { ap_ufixed x; x.raw = r; return x; }

    ap_ufixed() : raw(0) {}
    ap_ufixed(double x) { set_real(x); }
    ap_ufixed(float x) { set_real(double(x)); }
    ap_ufixed(int x) { set_real(double(x)); }
    ap_ufixed(unsigned x) { set_real(double(x)); }
    ap_ufixed(long x) { set_real(double(x)); }
    ap_ufixed(unsigned long x) { set_real(double(x)); }
    ap_ufixed(long long x) { set_real(double(x)); }
    ap_ufixed(unsigned long long x) { set_real(double(x)); }
    void set_real(double x) {
        if (!(x > 0.0)) {   // negative values and NaN: an unsigned type saturates (or wraps) at 0
            raw = 0;
            return;
        }
        const double scaled = std::ldexp(x, F);
        if (scaled >= std::ldexp(1.0, 100)) { raw = overflow(~(unsigned __int128)0 >> 8); return; }
        // keep one guard bit for the rounding decision (doubles carry 53 bits: exact for the float inputs used here)
        const unsigned __int128 w2 = (unsigned __int128)std::floor(std::ldexp(scaled, 1));
        raw = overflow(quantize(w2, 1));
    }
    operator float() const { return float(std::ldexp(double(raw), -F)); }
    double to_double() const { return std::ldexp(double(raw), -F); }

    friend ap_ufixed operator*(ap_ufixed a, ap_ufixed b) {   // exact product has 2F fraction bits
        return from_raw(overflow(quantize((unsigned __int128)a.raw * b.raw, F)));
    }
    friend ap_ufixed operator+(ap_ufixed a, ap_ufixed b) { return from_raw(overflow((unsigned __int128)a.raw + b.raw)); }
    friend bool operator<(ap_ufixed a, ap_ufixed b) { return a.raw < b.raw; }
    friend bool operator>(ap_ufixed a, ap_ufixed b) { return a.raw > b.raw; }
    friend bool operator==(ap_ufixed a, ap_ufixed b) { return a.raw == b.raw; }
    friend bool operator!=(ap_ufixed a, ap_ufixed b) { return a.raw != b.raw; }
    // mixed comparisons with built-in arithmetic types are made in double, as Vitis-HLS does
#define GLB_AP_CMP(OPR)                                                                                    \
    friend bool operator OPR(ap_ufixed a, double b) { return a.to_double() OPR b; }                        \
    friend bool operator OPR(double a, ap_ufixed b) { return a OPR b.to_double(); }                        \
    friend bool operator OPR(ap_ufixed a, float b) { return a.to_double() OPR double(b); }                 \
    friend bool operator OPR(float a, ap_ufixed b) { return double(a) OPR b.to_double(); }                 \
    friend bool operator OPR(ap_ufixed a, int b) { return a.to_double() OPR double(b); }                   \
    friend bool operator OPR(int a, ap_ufixed b) { return double(a) OPR b.to_double(); }
    GLB_AP_CMP(<) GLB_AP_CMP(>) GLB_AP_CMP(<=) GLB_AP_CMP(>=) GLB_AP_CMP(==) GLB_AP_CMP(!=)
#undef GLB_AP_CMP
    explicit operator bool() const { return raw != 0; }
    friend bool operator&&(ap_ufixed a, ap_ufixed b) { return a.raw != 0 && b.raw != 0; }
    friend bool operator||(ap_ufixed a, ap_ufixed b) { return a.raw != 0 || b.raw != 0; }
    ap_ufixed operator++(int) { ap_ufixed old(*this); *this = *this + ap_ufixed(1); return old; }
    ap_ufixed &operator++() { *this = *this + ap_ufixed(1); return *this; }
};

#endif  // GLB_AP_FIXED_FLOAT_BACKED

template <int W, int I, ap_q_mode Q = AP_TRN, ap_o_mode O = AP_WRAP>
using ap_fixed = ap_ufixed<W, I, Q, O>;

template <int W> struct ap_uint {
    unsigned long long v;
    ap_uint() : v(0) {}
    template <typename T> ap_uint(T x) : v(static_cast<unsigned long long>(x)) {}
    operator unsigned long long() const { return v; }
};

#endif  // ORACLE_SHIM_AP_FIXED_H_

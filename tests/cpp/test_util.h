// Shared helpers of the C++ tests: seeded synthetic matrices, oracle wrappers, comparisons.
// The oracle (oracle/liboracle.so, prototypes in oracle/oracle.h) is TEST infrastructure: it gives
// the results the reference's compute_reference_results would, and only tests link it.
#ifndef TEST_UTIL_H_
#define TEST_UTIL_H_
#include <algorithm>
#include <cstdint>
#include <random>
#include <set>
#include <vector>

#include "../../oracle/oracle.h"
#include "graphlily/global.h"
#include "graphlily/io/data_formatter.h"
#include "graphlily/io/data_loader.h"
#include "mini_test.h"

using graphlily::io::CSCMatrix;
using graphlily::io::CSRMatrix;
using dense_t = graphlily::aligned_dense_vec_t;
using sparse_t = graphlily::aligned_sparse_vec_t;

// n x n, exactly k distinct random columns per row, sorted: the shape of the reference's
// uniform_10K_10 dataset (tests/test_module_spmv_spmspv.cpp:174), which is not shipped.
inline CSRMatrix<float> uniform_csr(uint32_t n, uint32_t k, uint32_t seed, float value) {
    std::mt19937 rng(seed);
    CSRMatrix<float> m;
    m.num_rows = m.num_cols = n;
    m.adj_indptr.push_back(0);
    for (uint32_t r = 0; r < n; r++) {
        std::set<uint32_t> cols;
        while (cols.size() < k) cols.insert(rng() % n);
        for (uint32_t c : cols) { m.adj_indices.push_back(c); m.adj_data.push_back(value); }
        m.adj_indptr.push_back(uint32_t(m.adj_indices.size()));
    }
    return m;
}

// dense n x n (dense_32 / dense_1K of the reference tests)
inline CSRMatrix<float> dense_csr(uint32_t n, float value) {
    CSRMatrix<float> m;
    m.num_rows = m.num_cols = n;
    m.adj_indptr.push_back(0);
    for (uint32_t r = 0; r < n; r++) {
        for (uint32_t c = 0; c < n; c++) { m.adj_indices.push_back(c); m.adj_data.push_back(value); }
        m.adj_indptr.push_back(uint32_t(m.adj_indices.size()));
    }
    return m;
}

// skewed rows (a few giants, empty rows) with random values
inline CSRMatrix<float> skewed_csr(uint32_t n, uint32_t seed) {
    std::mt19937 rng(seed);
    CSRMatrix<float> m;
    m.num_rows = m.num_cols = n;
    m.adj_indptr.push_back(0);
    for (uint32_t r = 0; r < n; r++) {
        uint32_t deg = (r % 97 == 0) ? 3000 + rng() % 2000 : (r % 5 == 0 ? 0 : 1 + rng() % 12);
        deg = std::min(deg, n);
        std::set<uint32_t> cols;
        while (cols.size() < deg) cols.insert(rng() % n);
        for (uint32_t c : cols) { m.adj_indices.push_back(c); m.adj_data.push_back(float(rng() % 1000) / 1000.0f); }
        m.adj_indptr.push_back(uint32_t(m.adj_indices.size()));
    }
    return m;
}

#if !defined(GRAPHLILY_VAL_T_UNSIGNED) && !defined(GRAPHLILY_VAL_T_UFIXED)   // fp32 helpers over the fp32 oracle
inline dense_t random_01(uint32_t n, uint32_t seed) {
    std::mt19937 rng(seed);
    dense_t v(n);
    for (auto &x : v) x = float(rng() % 2);
    return v;
}

inline dense_t ref_spmv(const CSRMatrix<float> &m, graphlily::SemiringType s, graphlily::MaskType mt, const dense_t &x,
                        const dense_t &mask) {
    dense_t y(m.num_rows);
    oracle_spmv(m.num_rows, m.num_cols, m.adj_indptr.data(), m.adj_indices.data(), m.adj_data.data(), s.op, s.zero, mt,
                x.data(), mask.empty() ? nullptr : mask.data(), y.data());
    return y;
}

// |kernel - ref| <= 1e-5 * |ref| (the fp32 tolerance of this project; the reference's own tests use
// 1e-4 absolute, tests/test_module_spmv_spmspv.cpp:32-40), or bit-exact when `exact`.
inline void verify(const dense_t &ref, const dense_t &got, bool exact) {
    ASSERT_EQ(ref.size(), got.size());
    for (size_t i = 0; i < ref.size(); i++) {
        const bool ok = exact ? (std::memcmp(&ref[i], &got[i], 4) == 0 || ref[i] == got[i])
                              : (std::fabs(got[i] - ref[i]) <= 1e-5f * std::fabs(ref[i]) || std::fabs(got[i] - ref[i]) < 1e-12f);
        if (!ok) MT_FAIL_(true, "mismatch at %zu: reference %.9g kernel %.9g", i, ref[i], got[i]);
    }
}
#endif  // fp32
#endif

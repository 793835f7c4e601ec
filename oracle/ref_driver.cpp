// TEST INFRASTRUCTURE ONLY -- never linked, imported or executed by the product path.
//
// oracle/_ref/libgraphlily_ref.so: a C ABI around the REFERENCE'S OWN CPU code,
// compiled from the headers where they lie under /root/reference (nothing is
// copied).  Build recipe: oracle/Makefile (g++ -std=c++14 -O3, no -march, no
// -ffast-math; three stub headers under oracle/shim stand in for ap_fixed.h,
// xcl2.hpp and cnpy.h).
//
// What is exposed is exactly the reference's compute_reference_results() family
// plus the host-side IO helpers the apps run before it:
//   SpMVModule::compute_reference_results            spmv_module.h:478-532
//   SpMSpVModule::compute_reference_results          spmspv_module.h:446-520
//   eWiseAddModule::compute_reference_results        add_scalar_vector_dense_module.h:195-204
//   AssignVectorDenseModule::compute_reference_results   assign_vector_dense_module.h:223-246
//   AssignVectorSparseModule::compute_reference_results  assign_vector_sparse_module.h:306-335
//   BFS / PageRank / SSSP ::compute_reference_results    bfs.h:350-360, pagerank.h:150-159, sssp.h:246-253
//   csr2csc, util_round_csr_matrix_dim, util_normalize_csr_matrix_by_outdegree,
//   load_csr_matrix_from_float_npz, _preprocess (sssp.h:16-62)
//
// The reference keeps the matrix in private members that only
// load_and_format_matrix() sets -- and that function also builds the FPGA CPSR
// layout (seconds and GBs for large graphs).  `fast=1` entry points therefore
// assign those members directly (the headers are compiled with private made
// public); `fast=0` goes through the reference's load_and_format_matrix().  Both
// run the identical compute_reference_results() code; tests check they agree.

#include <algorithm>
#include <cassert>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <numeric>
#include <sstream>
#include <string>
#include <type_traits>
#include <vector>
#include <zlib.h>

#include "../include/graphlily/io/npz.h"  // in-repo npz reader backing the cnpy shim

#define private public
#define protected public
#include "graphlily/app/bfs.h"
#include "graphlily/app/pagerank.h"
#include "graphlily/app/sssp.h"
#undef private
#undef protected

using graphlily::aligned_dense_float_vec_t;
using graphlily::aligned_sparse_float_vec_t;
using graphlily::idx_float_t;
using graphlily::val_t;
using graphlily::io::CSCMatrix;
using graphlily::io::CSRMatrix;

namespace {

typedef graphlily::module::SpMVModule<val_t, val_t> SpMV;
typedef graphlily::module::SpMSpVModule<val_t, val_t, graphlily::idx_val_t> SpMSpV;

CSRMatrix<float> make_csr(uint32_t nrows, uint32_t ncols, const uint32_t *indptr, const uint32_t *indices,
                          const float *data) {
    CSRMatrix<float> m;
    m.num_rows = nrows;
    m.num_cols = ncols;
    uint32_t nnz = indptr[nrows];
    m.adj_indptr.assign(indptr, indptr + nrows + 1);
    m.adj_indices.assign(indices, indices + nnz);
    m.adj_data.assign(data, data + nnz);
    return m;
}

graphlily::SemiringType make_semiring(int op, float zero) {
    graphlily::SemiringType s;
    s.op = static_cast<graphlily::OperationType>(op);
    s.one = (op == graphlily::kAddMin) ? 0 : 1;
    s.zero = zero;
    return s;
}

// Install a float CSR as the module's matrix without building the FPGA layout.
void install_csr(SpMV &m, const CSRMatrix<float> &csr, bool fast) {
    if (fast) {
        m.csr_matrix_float_ = csr;
        m.csr_matrix_.num_rows = csr.num_rows;
        m.csr_matrix_.num_cols = csr.num_cols;
        m.csr_matrix_.adj_indptr = csr.adj_indptr;  // get_nnz() reads it
    } else {
        m.load_and_format_matrix(csr, false);
    }
}

void install_csc(SpMSpV &m, const CSCMatrix<float> &csc, bool fast) {
    if (fast) {
        m.csc_matrix_float_ = csc;
        m.csc_matrix_ = graphlily::io::csc_matrix_convert_from_float<val_t>(csc);
    } else {
        m.load_and_format_matrix(csc);
    }
}

}  // namespace

extern "C" {

// ---------------------------------------------------------------- module level
int ref_spmv(uint32_t nrows, uint32_t ncols, const uint32_t *indptr, const uint32_t *indices, const float *data,
             int op, float zero, int mask_type, const float *x, const float *mask, float *y, int fast) {
    CSRMatrix<float> csr = make_csr(nrows, ncols, indptr, indices, data);
    // Tiny buffer sizes when the real formatter runs keep its partition tables small.
    SpMV m(16, 1024 * 1024, 32 * 1024);
    m.set_semiring(make_semiring(op, zero));
    m.set_mask_type(static_cast<graphlily::MaskType>(mask_type));
    install_csr(m, csr, fast != 0);
    aligned_dense_float_vec_t xv(x, x + ncols);
    aligned_dense_float_vec_t out;
    if (mask_type == graphlily::kNoMask) {
        out = m.compute_reference_results(xv);
    } else {
        aligned_dense_float_vec_t mv(mask, mask + nrows);
        out = m.compute_reference_results(xv, mv);
    }
    std::memcpy(y, out.data(), sizeof(float) * nrows);
    return 0;
}

// Timed variant for the CPU baseline: matrix installed once, `reps` reference SpMVs,
// returns the best wall-clock seconds of a single compute_reference_results() call.
double ref_spmv_timed(uint32_t nrows, uint32_t ncols, const uint32_t *indptr, const uint32_t *indices,
                      const float *data, int op, float zero, const float *x, float *y, int reps) {
    CSRMatrix<float> csr = make_csr(nrows, ncols, indptr, indices, data);
    SpMV m(16, 1024 * 1024, 32 * 1024);
    m.set_semiring(make_semiring(op, zero));
    m.set_mask_type(graphlily::kNoMask);
    install_csr(m, csr, true);
    aligned_dense_float_vec_t xv(x, x + ncols);
    double best = 1e30;
    for (int r = 0; r < reps; r++) {
        auto t0 = std::chrono::steady_clock::now();
        aligned_dense_float_vec_t out = m.compute_reference_results(xv);
        auto t1 = std::chrono::steady_clock::now();
        best = std::min(best, std::chrono::duration<double>(t1 - t0).count());
        if (r == reps - 1) std::memcpy(y, out.data(), sizeof(float) * nrows);
    }
    return best;
}

// CSC arrays: indptr has ncols+1 entries, indices are row ids.
// x_idx / x_val: the x_nnz active entries (no head element; the driver builds the {nnz,-} head).
int ref_spmspv(uint32_t nrows, uint32_t ncols, const uint32_t *indptr, const uint32_t *indices, const float *data,
               int op, float zero, int mask_type, const uint32_t *x_idx, const float *x_val, uint32_t x_nnz,
               const float *mask, float *y, int fast) {
    CSCMatrix<float> csc;
    csc.num_rows = nrows;
    csc.num_cols = ncols;
    uint32_t nnz = indptr[ncols];
    csc.adj_indptr.assign(indptr, indptr + ncols + 1);
    csc.adj_indices.assign(indices, indices + nnz);
    csc.adj_data.assign(data, data + nnz);
    SpMSpV m(256 * 1024);
    m.set_semiring(make_semiring(op, zero));
    m.set_mask_type(static_cast<graphlily::MaskType>(mask_type));
    install_csc(m, csc, fast != 0);
    aligned_sparse_float_vec_t xv(x_nnz + 1);
    xv[0].index = x_nnz;
    xv[0].val = 0;
    for (uint32_t i = 0; i < x_nnz; i++) {
        xv[i + 1].index = x_idx[i];
        xv[i + 1].val = x_val[i];
    }
    aligned_dense_float_vec_t mv(nrows, 0.0f);
    if (mask) mv.assign(mask, mask + nrows);
    aligned_dense_float_vec_t out = m.compute_reference_results(xv, mv);
    std::memcpy(y, out.data(), sizeof(float) * nrows);
    return 0;
}

int ref_ewise_add(const float *in, float *out, uint32_t len, float val) {
    graphlily::module::eWiseAddModule<val_t> m;
    aligned_dense_float_vec_t iv(in, in + len);
    aligned_dense_float_vec_t ov = m.compute_reference_results(iv, len, val);
    std::memcpy(out, ov.data(), sizeof(float) * len);
    return 0;
}

int ref_assign_dense(const float *mask, float *inout, uint32_t len, float val, int mask_type) {
    graphlily::module::AssignVectorDenseModule<val_t> m;
    m.mask_type_ = static_cast<graphlily::MaskType>(mask_type);  // set_mask_type() exits on kNoMask
    if (mask_type == graphlily::kNoMask) return 1;               // reference: print + exit(EXIT_FAILURE)
    aligned_dense_float_vec_t mv(mask, mask + len), iv(inout, inout + len);
    m.compute_reference_results(mv, iv, len, val);
    std::memcpy(inout, iv.data(), sizeof(float) * len);
    return 0;
}

int ref_assign_sparse(const uint32_t *m_idx, const float *m_val, uint32_t m_nnz, float *inout, uint32_t len,
                      float val) {
    graphlily::module::AssignVectorSparseModule<val_t, graphlily::idx_val_t> m(false);
    aligned_sparse_float_vec_t mv(m_nnz + 1);
    mv[0].index = m_nnz;
    mv[0].val = 0;
    for (uint32_t i = 0; i < m_nnz; i++) { mv[i + 1].index = m_idx[i]; mv[i + 1].val = m_val ? m_val[i] : 0; }
    aligned_dense_float_vec_t iv(inout, inout + len);
    m.compute_reference_results(mv, iv, val);
    std::memcpy(inout, iv.data(), sizeof(float) * len);
    return 0;
}

// Returns the new-frontier count; nf_idx / nf_val must hold m_nnz entries.
int ref_assign_sparse_relax(const uint32_t *m_idx, const float *m_val, uint32_t m_nnz, float *inout, uint32_t len,
                            uint32_t *nf_idx, float *nf_val) {
    graphlily::module::AssignVectorSparseModule<val_t, graphlily::idx_val_t> m(true);
    aligned_sparse_float_vec_t mv(m_nnz + 1), nf;
    mv[0].index = m_nnz;
    mv[0].val = 0;
    for (uint32_t i = 0; i < m_nnz; i++) { mv[i + 1].index = m_idx[i]; mv[i + 1].val = m_val[i]; }
    aligned_dense_float_vec_t iv(inout, inout + len);
    m.compute_reference_results(mv, iv, nf);
    std::memcpy(inout, iv.data(), sizeof(float) * len);
    uint32_t n = nf[0].index;
    for (uint32_t i = 0; i < n; i++) { nf_idx[i] = nf[i + 1].index; nf_val[i] = nf[i + 1].val; }
    return int(n);
}

// ------------------------------------------------------------------- io level
int ref_csr2csc(uint32_t nrows, uint32_t ncols, const uint32_t *indptr, const uint32_t *indices, const float *data,
                uint32_t *o_indptr, uint32_t *o_indices, float *o_data) {
    CSCMatrix<float> csc = graphlily::io::csr2csc(make_csr(nrows, ncols, indptr, indices, data));
    std::copy(csc.adj_indptr.begin(), csc.adj_indptr.end(), o_indptr);
    std::copy(csc.adj_indices.begin(), csc.adj_indices.end(), o_indices);
    std::copy(csc.adj_data.begin(), csc.adj_data.end(), o_data);
    return 0;
}

// o_indptr must hold round_up(nrows,row_div)+1 entries. out_dims = {rows, cols}.
int ref_round_dim(uint32_t nrows, uint32_t ncols, const uint32_t *indptr, uint32_t row_div, uint32_t col_div,
                  uint32_t *o_indptr, uint32_t *out_dims) {
    CSRMatrix<float> m;
    m.num_rows = nrows;
    m.num_cols = ncols;
    m.adj_indptr.assign(indptr, indptr + nrows + 1);
    graphlily::io::util_round_csr_matrix_dim(m, row_div, col_div);
    std::copy(m.adj_indptr.begin(), m.adj_indptr.end(), o_indptr);
    out_dims[0] = m.num_rows;
    out_dims[1] = m.num_cols;
    return 0;
}

int ref_normalize_outdegree(uint32_t nrows, uint32_t ncols, const uint32_t *indptr, const uint32_t *indices,
                            float *data) {
    CSRMatrix<float> m = make_csr(nrows, ncols, indptr, indices, data);
    graphlily::io::util_normalize_csr_matrix_by_outdegree(m);
    std::copy(m.adj_data.begin(), m.adj_data.end(), data);
    return 0;
}

// SSSP _preprocess (sssp.h:16-62). Outputs sized nnz + nrows; returns the new nnz.
int64_t ref_sssp_preprocess(uint32_t nrows, uint32_t ncols, const uint32_t *indptr, const uint32_t *indices,
                            const float *data, uint32_t *o_indptr, uint32_t *o_indices, float *o_data) {
    CSRMatrix<float> m = make_csr(nrows, ncols, indptr, indices, data);
    _preprocess(m);
    std::copy(m.adj_indptr.begin(), m.adj_indptr.end(), o_indptr);
    std::copy(m.adj_indices.begin(), m.adj_indices.end(), o_indices);
    std::copy(m.adj_data.begin(), m.adj_data.end(), o_data);
    return int64_t(m.adj_indices.size());
}

// load_csr_matrix_from_float_npz (data_loader.h:51-70). Two-phase: pass NULL arrays to query dims.
// dims = {rows, cols, nnz}.
int ref_load_npz(const char *path, uint32_t *dims, uint32_t *o_indptr, uint32_t *o_indices, float *o_data) {
    CSRMatrix<float> m = graphlily::io::load_csr_matrix_from_float_npz(path);
    dims[0] = m.num_rows;
    dims[1] = m.num_cols;
    dims[2] = uint32_t(m.adj_data.size());
    if (o_indptr) std::copy(m.adj_indptr.begin(), m.adj_indptr.end(), o_indptr);
    if (o_indices) std::copy(m.adj_indices.begin(), m.adj_indices.end(), o_indices);
    if (o_data) std::copy(m.adj_data.begin(), m.adj_data.end(), o_data);
    return 0;
}

// ------------------------------------------------------------------ app level
// Full reference pipeline from an .npz on disk: load_and_format_matrix(path) then
// compute_reference_results().  `out` must hold round_up(rows,128) floats; returns that length.
int64_t ref_app_bfs_npz(const char *path, uint32_t source, uint32_t iters, float *out) {
    graphlily::app::BFS app(16, 1024 * 1024, 256 * 1024, 32 * 1024);
    app.load_and_format_matrix(path, true);
    aligned_dense_float_vec_t r = app.compute_reference_results(source, iters);
    std::memcpy(out, r.data(), sizeof(float) * r.size());
    return int64_t(r.size());
}

int64_t ref_app_pagerank_npz(const char *path, float damping, uint32_t iters, float *out) {
    graphlily::app::PageRank app(16, 1024 * 1024, 32 * 1024);
    app.load_and_format_matrix(path, damping, true);
    aligned_dense_float_vec_t r = app.compute_reference_results(damping, iters);
    std::memcpy(out, r.data(), sizeof(float) * r.size());
    return int64_t(r.size());
}

int64_t ref_app_sssp_npz(const char *path, uint32_t source, uint32_t iters, float *out) {
    graphlily::app::SSSP app(16, 1024 * 1024, 256 * 1024, 32 * 1024);
    app.load_and_format_matrix(path, true);
    aligned_dense_float_vec_t r = app.compute_reference_results(source, iters);
    std::memcpy(out, r.data(), sizeof(float) * r.size());
    return int64_t(r.size());
}

// Same three apps on an in-memory, ALREADY PREPROCESSED CSR (dims rounded, values set):
// only SpMV_/DenseAssign_/eWiseAdd_ reference loops run -- used for large graphs where the
// FPGA formatter would dominate.  The caller applies the app's own preprocessing through
// ref_round_dim / ref_normalize_outdegree / ref_sssp_preprocess above.
int ref_app_bfs_csr(uint32_t n, const uint32_t *indptr, const uint32_t *indices, const float *data, uint32_t source,
                    uint32_t iters, float *out) {
    graphlily::app::BFS app(16, 1024 * 1024, 256 * 1024, 32 * 1024);
    install_csr(*app.SpMV_, make_csr(n, n, indptr, indices, data), true);
    app.matrix_num_rows_ = app.matrix_num_cols_ = n;
    aligned_dense_float_vec_t r = app.compute_reference_results(source, iters);
    std::memcpy(out, r.data(), sizeof(float) * n);
    return 0;
}

int ref_app_pagerank_csr(uint32_t n, const uint32_t *indptr, const uint32_t *indices, const float *data,
                         float damping, uint32_t iters, float *out) {
    graphlily::app::PageRank app(16, 1024 * 1024, 32 * 1024);
    install_csr(*app.SpMV_, make_csr(n, n, indptr, indices, data), true);
    app.matrix_num_rows_ = app.matrix_num_cols_ = n;
    aligned_dense_float_vec_t r = app.compute_reference_results(damping, iters);
    std::memcpy(out, r.data(), sizeof(float) * n);
    return 0;
}

int ref_app_sssp_csr(uint32_t n, const uint32_t *indptr, const uint32_t *indices, const float *data,
                     uint32_t source, uint32_t iters, float *out) {
    graphlily::app::SSSP app(16, 1024 * 1024, 256 * 1024, 32 * 1024);
    install_csr(*app.SpMV_, make_csr(n, n, indptr, indices, data), true);
    app.matrix_num_rows_ = app.matrix_num_cols_ = n;
    aligned_dense_float_vec_t r = app.compute_reference_results(source, iters);
    std::memcpy(out, r.data(), sizeof(float) * n);
    return 0;
}

// Semiring constants as the reference header defines them (global.h:83-107).
void ref_constants(float *out) {
    out[0] = float(graphlily::ArithmeticSemiring.zero);
    out[1] = float(graphlily::LogicalSemiring.zero);
    out[2] = float(graphlily::TropicalSemiring.zero);
    out[3] = float(graphlily::ArithmeticSemiring.one);
    out[4] = float(graphlily::LogicalSemiring.one);
    out[5] = float(graphlily::TropicalSemiring.one);
    out[6] = float(graphlily::FLOAT_INF);
    out[7] = float(graphlily::pack_size * graphlily::num_hbm_channels);
}

}  // extern "C"

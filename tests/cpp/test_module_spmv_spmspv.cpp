// GPU tests of SpMVModule / SpMSpVModule against the oracle, in the shape of the reference's
// tests/test_module_spmv_spmspv.cpp: semirings x mask types on dense_32 and uniform_10K_10
// (:137-178), SpMSpV at vector sparsities 0 / 0.5 / 0.99 with the strided index pattern and
// (rand % 10) / 10 values (:197-214), results compared after densifying (:236-240).
#include "graphlily/module/spmspv_module.h"
#include "graphlily/module/spmv_module.h"
#include "test_util.h"

using namespace graphlily;

static const SemiringType kSemirings[] = {ArithmeticSemiring, LogicalSemiring, TropicalSemiring};
static const MaskType kMasks[] = {kNoMask, kMaskWriteToZero, kMaskWriteToOne};

static void check_spmv(const CSRMatrix<float> &csr, SemiringType semiring, MaskType mask_type, bool skip_empty_rows) {
    module::SpMVModule<val_t, val_t> spmv(16, 1024, 256);
    spmv.set_semiring(semiring);
    spmv.set_mask_type(mask_type);
    spmv.set_target("hw");
    spmv.set_up_runtime("ignored.xclbin");
    spmv.load_and_format_matrix(csr, skip_empty_rows);
    spmv.send_matrix_host_to_device();
    dense_t vector = random_01(csr.num_cols, 11), mask = random_01(csr.num_rows, 12);
    spmv.send_vector_host_to_device(vector);
    if (mask_type != kNoMask) spmv.send_mask_host_to_device(mask);
    spmv.run();
    dense_t kernel = spmv.send_results_device_to_host();
    dense_t ref = ref_spmv(csr, semiring, mask_type, vector, mask);
    verify(ref, kernel, semiring.op != kMulAdd);
}

TEST(SpMV, Dense32AllSemiringsAllMasks) {
    auto csr = dense_csr(32, 1.0f / 32);
    for (auto s : kSemirings)
        for (auto mt : kMasks)
            for (bool skip : {false, true}) check_spmv(csr, s, mt, skip);
}

TEST(SpMV, Uniform10K10) {
    auto csr = uniform_csr(10000, 10, 3, 1.0f / 10000);
    graphlily::io::util_round_csr_matrix_dim(csr, 128, 8);
    for (auto s : kSemirings)
        for (auto mt : kMasks) check_spmv(csr, s, mt, true);
}

TEST(SpMV, SkewedRowsWithEmptyRowsAndGiants) {
    auto csr = skewed_csr(6000, 5);
    for (auto s : kSemirings) check_spmv(csr, s, kMaskWriteToOne, true);
}

static void check_spmspv(const CSRMatrix<float> &csr, SemiringType semiring, MaskType mask_type, float sparsity) {
    CSCMatrix<float> csc = graphlily::io::csr2csc(csr);
    module::SpMSpVModule<val_t, val_t, idx_val_t> spmspv(512);
    spmspv.set_semiring(semiring);
    spmspv.set_mask_type(mask_type);
    spmspv.set_up_runtime("ignored.xclbin");
    spmspv.load_and_format_matrix(csc);
    spmspv.send_matrix_host_to_device();
    // strided active columns, values (rand % 10) / 10
    const uint32_t n = csr.num_cols;
    uint32_t nnz = uint32_t(std::floor(n * (1 - sparsity)));
    if (nnz == 0) nnz = 1;
    const uint32_t stride = n / nnz;
    std::mt19937 rng(7);
    sparse_t vector(nnz + 1);
    vector[0] = {nnz, 0};
    std::vector<uint32_t> x_idx(nnz);
    std::vector<float> x_val(nnz);
    for (uint32_t i = 0; i < nnz; i++) {
        x_idx[i] = i * stride;
        x_val[i] = float(rng() % 10) / 10;
        vector[i + 1] = {x_idx[i], x_val[i]};
    }
    dense_t mask = random_01(csr.num_rows, 13);
    spmspv.send_vector_host_to_device(vector);
    if (mask_type != kNoMask) spmspv.send_mask_host_to_device(mask);
    spmspv.run();
    sparse_t kernel_sparse = spmspv.send_results_device_to_host();
    ASSERT_EQ(kernel_sparse[0].index, spmspv.get_results_nnz());
    dense_t kernel = convert_sparse_vec_to_dense_vec<sparse_t, dense_t, val_t>(kernel_sparse, csr.num_rows, semiring.zero);
    dense_t ref(csr.num_rows);
    oracle_spmspv(csc.num_rows, csc.num_cols, csc.adj_indptr.data(), csc.adj_indices.data(), csc.adj_data.data(),
                  semiring.op, semiring.zero, mask_type, x_idx.data(), x_val.data(), nnz,
                  mask_type == kNoMask ? nullptr : mask.data(), ref.data());
    verify(ref, kernel, semiring.op != kMulAdd);
    // every listed entry differs from zero and no index repeats
    std::vector<char> seen(csr.num_rows, 0);
    for (uint32_t i = 1; i <= kernel_sparse[0].index; i++) {
        ASSERT_TRUE(kernel_sparse[i].index < csr.num_rows && !seen[kernel_sparse[i].index]);
        seen[kernel_sparse[i].index] = 1;
        ASSERT_TRUE(kernel_sparse[i].val != semiring.zero);
    }
}

TEST(SpMSpV, Dense1KAllSemiringsAllMasks) {
    auto csr = dense_csr(1024, 1.0f / 1024);
    for (auto s : kSemirings)
        for (auto mt : kMasks) check_spmspv(csr, s, mt, 0.5f);
}

TEST(SpMSpV, Uniform10K10Sparsities) {
    auto csr = uniform_csr(10000, 10, 3, 1.0f / 10000);
    graphlily::io::util_round_csr_matrix_dim(csr, 128, 128);
    for (float sp : {0.0f, 0.5f, 0.99f})
        for (auto s : kSemirings) check_spmspv(csr, s, kMaskWriteToZero, sp);
}

TEST(SpMSpV, HeavyColumns) {
    auto csr = skewed_csr(6000, 9);
    auto t = graphlily::io::csr2csc(csr);  // transpose so the giants become columns
    CSRMatrix<float> as_csr{t.num_cols, t.num_rows, t.adj_data, t.adj_indices, t.adj_indptr};
    for (auto s : kSemirings) check_spmspv(as_csr, s, kNoMask, 0.9f);
}

MINI_TEST_MAIN

// GPU tests of the apply modules against the oracle, in the shape of the reference's
// tests/test_module_apply.cpp: eWiseAdd (:54-75), dense assign (:78-103), sparse assign in BFS mode
// (:106-143) and SSSP mode incl. the new frontier (:146-206), device copy + buffer aliasing (:209-261).
#include "graphlily/module/add_scalar_vector_dense_module.h"
#include "graphlily/module/assign_vector_dense_module.h"
#include "graphlily/module/assign_vector_sparse_module.h"
#include "test_util.h"

using namespace graphlily;

TEST(AddScalarVectorDense, Basic) {
    module::eWiseAddModule<val_t> m;
    m.set_up_runtime("ignored.xclbin");
    for (uint32_t len : {128u, 1u, 1000003u}) {
        dense_t in(len);
        for (uint32_t i = 0; i < len; i++) in[i] = float(i % 1000) * 0.25f;
        m.send_in_host_to_device(in);
        m.allocate_out_buf(len);
        m.run(len, 1.5f);
        dense_t ref(len);
        oracle_ewise_add(in.data(), ref.data(), len, 1.5f);
        verify(ref, m.send_out_device_to_host(), true);
    }
}

TEST(AssignVectorDense, BothMaskTypes) {
    for (MaskType mt : {kMaskWriteToOne, kMaskWriteToZero}) {
        module::AssignVectorDenseModule<val_t> m;
        m.set_mask_type(mt);
        m.set_up_runtime("ignored.xclbin");
        const uint32_t len = 4099;
        dense_t mask = random_01(len, 3), inout(len, 7.0f), ref = inout;
        m.send_mask_host_to_device(mask);
        m.send_inout_host_to_device(inout);
        m.run(len, 23);
        oracle_assign_dense(mask.data(), ref.data(), len, 23, mt);
        verify(ref, m.send_inout_device_to_host(), true);
    }
}

TEST(AssignVectorSparse, NoNewFrontier) {
    module::AssignVectorSparseModule<val_t, idx_val_t> m(false);
    m.set_up_runtime("ignored.xclbin");
    const uint32_t len = 8192, nnz = 777;
    sparse_t mask(nnz + 1);
    std::vector<uint32_t> idx(nnz);
    mask[0] = {nnz, 0};
    for (uint32_t i = 0; i < nnz; i++) { idx[i] = (i * 37) % len; mask[i + 1] = {idx[i], float(i)}; }
    dense_t inout(len, 0.0f), ref = inout;
    m.send_mask_host_to_device(mask);
    m.send_inout_host_to_device(inout);
    m.run(9.0f);
    oracle_assign_sparse(idx.data(), nnz, ref.data(), 9.0f);
    verify(ref, m.send_inout_device_to_host(), true);
}

TEST(AssignVectorSparse, NewFrontier) {
    module::AssignVectorSparseModule<val_t, idx_val_t> m(true);
    m.set_up_runtime("ignored.xclbin");
    const uint32_t len = 8192, nnz = 3000;
    std::mt19937 rng(5);
    sparse_t mask(nnz + 1);
    std::vector<uint32_t> idx(nnz);
    std::vector<float> val(nnz);
    mask[0] = {nnz, 0};
    for (uint32_t i = 0; i < nnz; i++) { idx[i] = (i * 2) % len; val[i] = float(rng() % 10); mask[i + 1] = {idx[i], val[i]}; }
    dense_t inout(len), ref;
    for (auto &x : inout) x = float(rng() % 10);
    ref = inout;
    m.send_mask_host_to_device(mask);
    m.send_inout_host_to_device(inout);
    m.run();
    std::vector<uint32_t> nf_idx(nnz);
    std::vector<float> nf_val(nnz);
    const int n_new = oracle_assign_sparse_relax(idx.data(), val.data(), nnz, ref.data(), nf_idx.data(), nf_val.data());
    verify(ref, m.send_inout_device_to_host(), true);
    sparse_t nf = m.send_new_frontier_device_to_host();
    ASSERT_EQ(nf[0].index, uint32_t(n_new));
    sparse_t ref_nf(n_new + 1);
    ref_nf[0] = {uint32_t(n_new), 0};
    for (int i = 0; i < n_new; i++) ref_nf[i + 1] = {nf_idx[i], nf_val[i]};
    // order is unspecified: compare after densifying, as the reference's test does
    verify(convert_sparse_vec_to_dense_vec<sparse_t, dense_t, val_t>(ref_nf, len, -1.0f),
           convert_sparse_vec_to_dense_vec<sparse_t, dense_t, val_t>(nf, len, -1.0f), true);
}

TEST(DataTransfer, CopyAndBindBuffers) {
    module::eWiseAddModule<val_t> a, b;
    a.set_up_runtime("ignored.xclbin");
    b.set_runtime(a.get_runtime());
    const uint32_t len = 1024;
    dense_t in(len);
    for (uint32_t i = 0; i < len; i++) in[i] = float(i);
    a.send_in_host_to_device(in);
    a.allocate_out_buf(len);
    a.run(len, 1.0f);                 // a.out = in + 1
    b.bind_in_buf(a.out_buf);         // alias
    b.allocate_out_buf(len);
    b.run(len, 2.0f);                 // b.out = in + 3
    dense_t got = b.send_out_device_to_host();
    for (uint32_t i = 0; i < len; i++) ASSERT_EQ(got[i], float(i) + 3.0f);
    a.copy_buffer_device_to_device(b.out_buf, a.out_buf, sizeof(val_t) * len);
    got = a.send_out_device_to_host();
    for (uint32_t i = 0; i < len; i++) ASSERT_EQ(got[i], float(i) + 3.0f);
}

MINI_TEST_MAIN

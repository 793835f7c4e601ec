// graphlily-b200: SpMSpV operator module, sparse y = A[:, idx(x)] (+).(x) x with a dense mask.
//
// Same public surface as /root/reference/graphlily/module/spmspv_module.h:26-254.  formatCSC +
// upload (:264-370) becomes glb_csc_create; run() (:437-441) becomes glb_spmspv; get_results_nnz
// (:239-242) reads slot 0 of the result list.  Vectors follow the sparse convention of global.h:
// element 0 = {nnz, -}.
#ifndef GRAPHLILY_SPMSPV_MODULE_H_
#define GRAPHLILY_SPMSPV_MODULE_H_

#include <type_traits>

#include "graphlily/global.h"
#include "graphlily/io/data_loader.h"
#include "graphlily/module/base_module.h"

using graphlily::io::CSCMatrix;  // at file scope, as in the reference (spmspv_module.h:17)

namespace graphlily {
namespace module {

template <typename matrix_data_t, typename vector_data_t, typename idx_val_t>
class SpMSpVModule : public BaseModule {
    static_assert(std::is_same<matrix_data_t, vector_data_t>::value && sizeof(vector_data_t) == 4,
                  "matrix and vector share one 32-bit value type: float, unsigned or graphlily::ufixed_32_8");
    static_assert(sizeof(idx_val_t) == sizeof(glb_idx_val_t), "idx_val_t must match the C ABI");
    using VT = graphlily::val_traits<vector_data_t>;
private:
    graphlily::MaskType mask_type_ = graphlily::kNoMask;
    graphlily::SemiringType semiring_ = graphlily::ArithmeticSemiring;
    uint32_t out_buf_len_;  // FPGA tuning argument: accepted, unused
    using aligned_dense_vec_t = std::vector<vector_data_t, aligned_allocator<vector_data_t>>;
    using aligned_sparse_vec_t = std::vector<idx_val_t, aligned_allocator<idx_val_t>>;
    aligned_sparse_vec_t vector_, results_;
    aligned_dense_vec_t mask_;
    CSCMatrix<float> csc_matrix_float_;
    glb_csc_t matrix_ = nullptr;

    uint32_t count_of(const DeviceBuffer &list) {
        uint32_t n = 0;
        GLB_CHECK(glb_sparse_count(ctx(), list.sparse(), &n));
        return n;
    }

public:
    // Device buffers
    DeviceBuffer vector_buf;
    DeviceBuffer mask_buf;
    DeviceBuffer results_buf;

    explicit SpMSpVModule(uint32_t out_buf_len) : BaseModule("overlay"), out_buf_len_(out_buf_len) {}
    ~SpMSpVModule() override { glb_csc_destroy(matrix_); }

    uint32_t get_num_rows() { return csc_matrix_float_.num_rows; }
    uint32_t get_num_cols() { return csc_matrix_float_.num_cols; }
    uint32_t get_nnz() { return csc_matrix_float_.adj_indptr[csc_matrix_float_.num_cols]; }
    void set_semiring(graphlily::SemiringType semiring) { semiring_ = semiring; }
    void set_mask_type(graphlily::MaskType mask_type) { mask_type_ = mask_type; }

    void load_and_format_matrix(CSCMatrix<float> const &csc_matrix_float) { csc_matrix_float_ = csc_matrix_float; }

    // Matrix upload + the result (rows + 1) and vector (cols + 1) lists, spmspv_module.h:290-370.
    void send_matrix_host_to_device() { send_matrix_host_to_device(0, csc_matrix_float_.num_rows); }
    // Row shard of the push direction (one process per GPU): of every column only the entries whose row lies in
    // [row_begin, row_end) go to this device; run() then lists only rows of the shard (ids stay global).
    void send_matrix_host_to_device(uint32_t row_begin, uint32_t row_end) {
        const CSCMatrix<float> &m = csc_matrix_float_;
        glb_csc_destroy(matrix_);
        matrix_ = nullptr;
        const float *data = m.adj_data.data();
        std::vector<uint32_t> words;   // the stored words of the value type ((val_t)x, spmspv_module.h:300-310)
        if (VT::id != GLB_VAL_F32) {
            words.resize(m.adj_data.size());
            for (size_t i = 0; i < words.size(); i++) words[i] = VT::bits(VT::from_float(m.adj_data[i]));
            data = reinterpret_cast<const float *>(words.data());
        }
        if (row_begin == 0 && row_end == m.num_rows)
            GLB_CHECK(glb_csc_create(ctx(), m.num_rows, m.num_cols, m.adj_indptr.data(), m.adj_indices.data(), data, &matrix_));
        else
            GLB_CHECK(glb_csc_create_rows(ctx(), m.num_rows, m.num_cols, m.adj_indptr.data(), m.adj_indices.data(), data,
                                          row_begin, row_end, &matrix_));
        aligned_sparse_vec_t empty_rows(size_t(m.num_rows) + 1, idx_val_t{0, vector_data_t(0)});
        aligned_sparse_vec_t empty_cols(size_t(m.num_cols) + 1, idx_val_t{0, vector_data_t(0)});
        results_buf = upload(empty_rows);
        vector_buf = upload(empty_cols);
    }

    void send_vector_host_to_device(aligned_sparse_vec_t &vector) {
        vector_ = vector;
        if (!vector_buf.valid() || vector_buf.bytes() < vector_.size() * sizeof(idx_val_t))
            vector_buf = DeviceBuffer(runtime_, vector_.size() * sizeof(idx_val_t));
        GLB_CHECK(glb_buffer_h2d(ctx(), vector_buf.ptr(), vector_.data(), vector_.size() * sizeof(idx_val_t)));
    }
    void send_mask_host_to_device(aligned_dense_vec_t &mask) {
        mask_ = mask;
        mask_buf = upload(mask_);
    }
    void bind_mask_buf(DeviceBuffer src_buf) { mask_buf = src_buf; }
    void set_mask_constant(vector_data_t value, uint32_t index, vector_data_t index_value) {
        mask_buf = constant_on_device(mask_buf, get_num_rows(), graphlily::val_container(value), true, index,
                                      graphlily::val_container(index_value));
    }

    void run() {
        const float *mask_ptr = mask_type_ == graphlily::kNoMask ? nullptr : mask_buf.f32();
        if (VT::id == GLB_VAL_F32)
            GLB_CHECK(glb_spmspv(ctx(), matrix_, semiring_.op, float(semiring_.zero), mask_type_, vector_buf.sparse(), mask_ptr,
                                 results_buf.sparse()));
        else
            GLB_CHECK(glb_spmspv_vt(ctx(), matrix_, VT::id, semiring_.op, VT::bits(vector_data_t(semiring_.zero)), mask_type_,
                                    vector_buf.sparse(), mask_ptr, results_buf.sparse()));
        end_run();
    }
    static constexpr bool fused_levels_available() { return VT::id == GLB_VAL_F32; }   // run_fused computes in fp32

    // A whole push level in one launch over explicit list buffers: SpMSpV + the sparse assign / relax of the
    // apps + (next != nullptr) the push-or-pull decision of pull_push taken on the device (glb_spmspv_fused).
    void run_fused(const DeviceBuffer &vector, const DeviceBuffer &results, const glb_spmspv_epilogue_t *epilogue,
                   const glb_spmspv_next_t *next) {
        assert(VT::id == GLB_VAL_F32 && "the fused push levels compute in fp32");
        GLB_CHECK(glb_spmspv_fused(ctx(), matrix_, semiring_.op, float(semiring_.zero), mask_type_, vector.sparse(),
                                   mask_type_ == graphlily::kNoMask ? nullptr : mask_buf.f32(), results.sparse(), epilogue, next));
    }
    // the one-entry start frontier, written on the device (no blocking upload)
    void set_vector_single(uint32_t index, vector_data_t val) {
        GLB_CHECK(glb_sparse_fill_one(ctx(), vector_buf.sparse(), index, graphlily::val_container(val)));
    }
    void home_lists() {  // canonical roles of the two list buffers at the start of a run: recorded sequences are found again
        if (vector_buf.valid() && results_buf.valid() && vector_buf.bytes() == results_buf.bytes() &&
            vector_buf.ptr() > results_buf.ptr())
            std::swap(vector_buf, results_buf);
    }
    uint32_t push_levels() {  // levels of the last pull_push run that carried the decision (blocking, on demand)
        uint32_t keep = 0, levels = 0;
        GLB_CHECK(glb_spmspv_push_state(ctx(), matrix_, &keep, &levels));
        return levels;
    }
    void reset_push_levels() { GLB_CHECK(glb_spmspv_reset_levels(ctx(), matrix_)); }
    glb_csc_t device_matrix() { return matrix_; }

    aligned_sparse_vec_t send_vector_device_to_host() {
        download(vector_, vector_buf, size_t(count_of(vector_buf)) + 1);
        return vector_;
    }
    aligned_dense_vec_t send_mask_device_to_host() {
        download(mask_, mask_buf, get_num_rows());
        return mask_;
    }
    aligned_sparse_vec_t send_results_device_to_host() {
        download(results_, results_buf, size_t(count_of(results_buf)) + 1);
        return results_;
    }
    uint32_t get_results_nnz() { return count_of(results_buf); }
    uint32_t get_vector_nnz() { return count_of(vector_buf); }

    // compute_reference_results (reference: spmspv_module.h:446-520) is DECLARED here so that the reference's own callers
    // compile, but the product does not contain a CPU implementation: the definition is test
    // infrastructure (tests/cpp/ref_compat/reference_results.h, which links oracle/); without it the call
    // fails at link time.
    graphlily::aligned_dense_float_vec_t compute_reference_results(graphlily::aligned_sparse_float_vec_t &vector,
                                                                   graphlily::aligned_dense_float_vec_t &mask);

    CSCMatrix<float> const &host_matrix() { return csc_matrix_float_; }
    graphlily::SemiringType semiring() const { return semiring_; }
    graphlily::MaskType mask_type() const { return mask_type_; }
};

}  // namespace module
}  // namespace graphlily

#endif  // GRAPHLILY_SPMSPV_MODULE_H_

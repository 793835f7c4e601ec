// graphlily-b200: SpMV operator module, y = A (+).(x) x with an optional dense mask.
//
// Same public surface as /root/reference/graphlily/module/spmv_module.h:26-272 (constructor
// arguments, set_semiring / set_mask_type, load_and_format_matrix, send_*_host_to_device,
// bind_mask_buf, run, send_*_device_to_host, public vector_buf / mask_buf / results_buf).  The CPSR
// formatting + 16-channel upload (:282-420) becomes glb_csr_create (lane-segment layout); run()
// (:471-475: setArg + enqueueTask + finish) becomes glb_spmv on the runtime's stream followed by a
// stream synchronisation, as in the reference (set_async_run(true) drops it); run_fused, which the apps'
// loops use, only enqueues.
// compute_reference_results (:478-532) is declared but not defined by the product: the CPU restatement
// lives in oracle/ and is linked by the tests only (tests/cpp/ref_compat/reference_results.h).
#ifndef GRAPHLILY_SPMV_MODULE_H_
#define GRAPHLILY_SPMV_MODULE_H_

#include <type_traits>

#include "graphlily/global.h"
#include "graphlily/io/data_loader.h"
#include "graphlily/module/base_module.h"

using graphlily::io::CSRMatrix;  // at file scope, as in the reference (spmv_module.h:17): its callers name it unqualified

namespace graphlily {
namespace module {

template <typename matrix_data_t, typename vector_data_t>
class SpMVModule : public BaseModule {
    static_assert(std::is_same<matrix_data_t, vector_data_t>::value && sizeof(vector_data_t) == 4,
                  "matrix and vector share one 32-bit value type: float, unsigned or graphlily::ufixed_32_8");
    using VT = graphlily::val_traits<vector_data_t>;   // the kernels' value type (GLB_VAL_*): semiring.cuh
private:
    graphlily::MaskType mask_type_ = graphlily::kNoMask;
    graphlily::SemiringType semiring_ = graphlily::ArithmeticSemiring;
    uint32_t num_channels_, out_buf_len_, vec_buf_len_;  // FPGA tuning arguments: accepted, unused
    using aligned_dense_vec_t = std::vector<vector_data_t, aligned_allocator<vector_data_t>>;
    aligned_dense_vec_t vector_, mask_, results_;
    CSRMatrix<float> csr_matrix_float_;
    glb_csr_t matrix_ = nullptr;

public:
    // Device buffers
    DeviceBuffer vector_buf;
    DeviceBuffer mask_buf;
    DeviceBuffer results_buf;

    SpMVModule(uint32_t num_channels, uint32_t out_buf_len, uint32_t vec_buf_len)
        : BaseModule("overlay"), num_channels_(num_channels), out_buf_len_(out_buf_len), vec_buf_len_(vec_buf_len) {}
    ~SpMVModule() override { glb_csr_destroy(matrix_); }

    void set_semiring(graphlily::SemiringType semiring) { semiring_ = semiring; }
    void set_mask_type(graphlily::MaskType mask_type) { mask_type_ = mask_type; }
    uint32_t get_num_rows() { return csr_matrix_float_.num_rows; }
    uint32_t get_num_cols() { return csr_matrix_float_.num_cols; }
    uint32_t get_nnz() { return csr_matrix_float_.adj_indptr[csr_matrix_float_.num_rows]; }

    // skip_empty_rows is a CPSR option (spmv_module.h:306-312); the lane-segment layout always
    // keeps empty rows out of the nnz stream.
    void load_and_format_matrix(CSRMatrix<float> const &csr_matrix_float, bool /*skip_empty_rows*/) {
        csr_matrix_float_ = csr_matrix_float;
    }

    // Builds the device layout for rows [row_begin, row_end) (default: all) and uploads it.
    void send_matrix_host_to_device(uint32_t row_begin = 0, uint32_t row_end = 0xffffffffu) {
        const CSRMatrix<float> &m = csr_matrix_float_;
        if (row_end == 0xffffffffu) row_end = m.num_rows;
        glb_csr_destroy(matrix_);
        matrix_ = nullptr;
        const float *data = m.adj_data.data();
        std::vector<uint32_t> words;   // (val_t)x of the reference's formatter: the stored words of the value type
        if (VT::id != GLB_VAL_F32) {
            words.resize(m.adj_data.size());
            for (size_t i = 0; i < words.size(); i++) words[i] = VT::bits(VT::from_float(m.adj_data[i]));
            data = reinterpret_cast<const float *>(words.data());
        }
        GLB_CHECK(glb_csr_create(ctx(), m.num_rows, m.num_cols, m.adj_indptr.data(), m.adj_indices.data(),
                                 data, row_begin, row_end, &matrix_));
        results_buf = DeviceBuffer(runtime_, sizeof(vector_data_t) * m.num_rows);
        GLB_CHECK(glb_buffer_fill_f32(ctx(), results_buf.f32(), 0.0f, m.num_rows));
    }

    void send_vector_host_to_device(aligned_dense_vec_t &vector) {
        vector_ = vector;
        vector_buf = upload(vector_);
    }
    void send_mask_host_to_device(aligned_dense_vec_t &mask) {
        mask_ = mask;
        mask_buf = upload(mask_);
    }
    void bind_mask_buf(DeviceBuffer src_buf) { mask_buf = src_buf; }

    // Start vectors of the apps (constant but for the source entry) without a host upload.
    // The app loops ping-pong vector / results and end with the roles swapped after an odd number of
    // iterations; a new run starts from the same assignment every time (lower address = vector), so
    // the recorded launch sequence of the previous run with these arguments is found again.
    void ensure_vector_buf() {  // allocate first, then settle the roles: the very first run is canonical too
        if (!vector_buf.valid() || vector_buf.bytes() != sizeof(vector_data_t) * size_t(get_num_cols()))
            vector_buf = DeviceBuffer(runtime_, sizeof(vector_data_t) * size_t(get_num_cols()));
        home_buffers();
    }
    void home_buffers() {
        if (vector_buf.valid() && results_buf.valid() && vector_buf.bytes() == results_buf.bytes() &&
            vector_buf.ptr() > results_buf.ptr())
            std::swap(vector_buf, results_buf);
    }
    void set_vector_constant(vector_data_t value) {
        ensure_vector_buf();
        vector_buf = constant_on_device(vector_buf, get_num_cols(), graphlily::val_container(value));
    }
    void set_vector_constant(vector_data_t value, uint32_t index, vector_data_t index_value) {
        ensure_vector_buf();
        vector_buf = constant_on_device(vector_buf, get_num_cols(), graphlily::val_container(value), true, index,
                                        graphlily::val_container(index_value));
    }
    void set_mask_constant(vector_data_t value, uint32_t index, vector_data_t index_value) {
        mask_buf = constant_on_device(mask_buf, get_num_rows(), graphlily::val_container(value), true, index,
                                      graphlily::val_container(index_value));
    }

    void run() {
        run_fused(nullptr);
        end_run();
    }

    // One launch for SpMV + the eWiseAdd / dense assign the apps run right after it (glb_spmv_fused).
    // (for unsigned / ufixed_32_8 the epilogue's add_val / assign_val carry the word: graphlily::val_container)
    void run_fused(const glb_spmv_epilogue_t *epilogue) { run_fused(vector_buf, mask_buf, results_buf, epilogue); }

    // The same launch over explicit buffers: the app loops ping-pong vector / results instead of copying.
    void run_fused(const DeviceBuffer &vector, const DeviceBuffer &mask, const DeviceBuffer &results,
                   const glb_spmv_epilogue_t *epilogue) {
        const float *mask_ptr = mask_type_ == graphlily::kNoMask ? nullptr : mask.f32();
        if (VT::id == GLB_VAL_F32)
            GLB_CHECK(glb_spmv_fused(ctx(), matrix_, semiring_.op, float(semiring_.zero), mask_type_, vector.f32(), mask_ptr,
                                     results.f32(), epilogue));
        else
            GLB_CHECK(glb_spmv_vt(ctx(), matrix_, VT::id, semiring_.op, VT::bits(vector_data_t(semiring_.zero)), mask_type_,
                                  vector.ptr(), mask_ptr, results.ptr(), epilogue));
    }

    // Row-sharded pull loop in one call (glb_spmv_exchange_iterate): vector -> results -> vector ... over two
    // vectors of the exchange, eps = nullptr or one fused epilogue per step.
    void iterate_exchange(Exchange &xc, const DeviceBuffer &vector, const DeviceBuffer &mask, const DeviceBuffer &results,
                          const glb_spmv_epilogue_t *eps, int n_steps) {
        assert(vector.exchange_vector() >= 0 && results.exchange_vector() >= 0);
        assert(VT::id == GLB_VAL_F32 && "row-sharded runs compute in fp32");
        GLB_CHECK(glb_spmv_exchange_iterate(ctx(), matrix_, semiring_.op, float(semiring_.zero), mask_type_, xc.handle(),
                                            vector.exchange_vector(), results.exchange_vector(),
                                            mask_type_ == graphlily::kNoMask ? nullptr : mask.f32(), eps, n_steps, nullptr));
    }

    aligned_dense_vec_t send_vector_device_to_host() {
        download(vector_, vector_buf, get_num_cols());
        return vector_;
    }
    aligned_dense_vec_t send_mask_device_to_host() {
        download(mask_, mask_buf, get_num_rows());
        return mask_;
    }
    aligned_dense_vec_t send_results_device_to_host() {
        download(results_, results_buf, get_num_rows());
        return results_;
    }

    // compute_reference_results (reference: spmv_module.h:478-532) is DECLARED here so that the reference's own callers
    // compile, but the product does not contain a CPU implementation: the definition is test
    // infrastructure (tests/cpp/ref_compat/reference_results.h, which links oracle/); without it the call
    // fails at link time.
    graphlily::aligned_dense_float_vec_t compute_reference_results(graphlily::aligned_dense_float_vec_t &vector);
    graphlily::aligned_dense_float_vec_t compute_reference_results(graphlily::aligned_dense_float_vec_t &vector,
                                                                   graphlily::aligned_dense_float_vec_t &mask);

    glb_csr_t device_matrix() { return matrix_; }
    CSRMatrix<float> const &host_matrix() { return csr_matrix_float_; }
    graphlily::SemiringType semiring() const { return semiring_; }
    graphlily::MaskType mask_type() const { return mask_type_; }
};

}  // namespace module
}  // namespace graphlily

#endif  // GRAPHLILY_SPMV_MODULE_H_

// SpMV throughput: the flow of /root/reference/benchmark/bench_spmv.cpp:37-113 on the B200 engine.
// values = 1 / num_rows, x = rand() % 2, 100 timed runs, GTEPS = nnz / seconds / 1e9 (:96-112).
#include <algorithm>

#include "bench_common.h"
#include "graphlily/io/data_formatter.h"
#include "graphlily/module/spmv_module.h"

int main(int argc, char *argv[]) {
    BenchArgs args = parse_args(argc, argv, "bench_spmv [num_channels out_buf_len vec_buf_len bitstream] <dataset.npz>");
    graphlily::module::SpMVModule<graphlily::val_t, graphlily::val_t> spmv(16, 0, 0);
    spmv.set_target("hw");
    spmv.set_mask_type(graphlily::kNoMask);
    spmv.set_semiring(graphlily::ArithmeticSemiring);
    spmv.set_up_runtime("");

    graphlily::io::CSRMatrix<float> csr_matrix = graphlily::io::load_csr_matrix_from_float_npz(args.dataset);
    for (auto &x : csr_matrix.adj_data) x = 1.0 / csr_matrix.num_rows;
    graphlily::io::util_round_csr_matrix_dim(csr_matrix, graphlily::num_hbm_channels * graphlily::pack_size,
                                             graphlily::pack_size);
    graphlily::aligned_dense_vec_t vector(csr_matrix.num_cols);
    std::generate(vector.begin(), vector.end(), [&] { return float(rand() % 2); });

    auto t0 = std::chrono::high_resolution_clock::now();
    spmv.load_and_format_matrix(csr_matrix, true);
    spmv.send_matrix_host_to_device();
    std::cout << "finished load_and_format_matrix + upload in " << seconds_since(t0) << " s" << std::endl;
    spmv.send_vector_host_to_device(vector);

    spmv.run();
    auto kernel_results = spmv.send_results_device_to_host();

    // bench_spmv.cpp:96-112: 100 runs; run() returns when the kernel has (as the reference's does)
    const uint32_t num_runs = 100;
    auto t1 = std::chrono::high_resolution_clock::now();
    for (size_t i = 0; i < num_runs; i++) spmv.run();
    spmv.get_runtime()->finish();
    const double average_time_in_sec = seconds_since(t1) / num_runs;
    std::cout << "average_time: " << average_time_in_sec * 1000 << " ms" << std::endl;
    std::cout << "Compute THROUGHPUT = " << double(spmv.get_nnz()) / 1e9 / average_time_in_sec << " GTEPS" << std::endl;
    // the same 100 launches enqueued back to back, one synchronisation at the end
    spmv.set_async_run(true);
    auto t2 = std::chrono::high_resolution_clock::now();
    for (size_t i = 0; i < num_runs; i++) spmv.run();
    spmv.get_runtime()->finish();
    const double async_time_in_sec = seconds_since(t2) / num_runs;
    std::cout << "back-to-back average_time: " << async_time_in_sec * 1000 << " ms, "
              << double(spmv.get_nnz()) / 1e9 / async_time_in_sec << " GTEPS" << std::endl;
    return 0;
}

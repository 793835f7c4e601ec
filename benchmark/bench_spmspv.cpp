// SpMSpV sweep: the flow of /root/reference/benchmark/bench_spmspv.cpp:256-330 -- for each dataset
// and vector sparsity 0.90 ... 0.9999 (strided active columns, as its set_up_bench_case), 20 timed
// runs; throughput = bytes of the active matrix columns (8 B per non-zero, measure_data_usage
// :61-76) / time, GTEPS = that / 8.  Usage: bench_spmspv [hw xclbin] <dataset.npz>... [logfile.txt]
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <sstream>

#include "bench_common.h"
#include "graphlily/io/data_formatter.h"
#include "graphlily/module/spmspv_module.h"

using namespace graphlily;

static double measure_data_usage(const io::CSCMatrix<float> &matrix, const aligned_sparse_vec_t &vector) {
    double bytes = 0;
    for (uint32_t k = 0; k < vector[0].index; k++) {
        const idx_t c = vector[k + 1].index;
        bytes += double(sizeof(val_t) + sizeof(idx_t)) * (matrix.adj_indptr[c + 1] - matrix.adj_indptr[c]);
    }
    return bytes;
}

int main(int argc, char *argv[]) {
    std::vector<std::string> datasets;
    std::string logfile;
    for (int i = 1; i < argc; i++) {
        std::string s = argv[i];
        if (ends_with(s, ".npz")) datasets.push_back(s);
        else if (ends_with(s, ".txt") || ends_with(s, ".log")) logfile = s;
    }
    if (datasets.empty()) {
        std::cerr << "usage: bench_spmspv [hw xclbin] <dataset.npz>... [logfile.txt]" << std::endl;
        return EXIT_FAILURE;
    }
    const float sparsities[] = {0.50f, 0.90f, 0.95f, 0.99f, 0.995f, 0.999f, 0.9995f, 0.9999f, 0.99999f};
    struct Case { const char *name; SemiringType semiring; } semirings[] = {
        {"Arithmetic", ArithmeticSemiring}, {"Logical", LogicalSemiring}, {"Tropical", TropicalSemiring}};
    // roofline denominator: GLB_HBM_PEAK_GBS (the measured copy bandwidth, MEASURED_PEAKS.json) or the fallback
    const char *peak_env = std::getenv("GLB_HBM_PEAK_GBS");
    const double peak_gbs = peak_env ? std::atof(peak_env) : 6650.0;
    std::ostringstream table;
    table << std::setw(28) << "test case" << std::setw(12) << "semiring" << std::setw(10) << "sparsity" << std::setw(10) << "nnz(x)"
          << std::setw(12) << "time(us)" << std::setw(12) << "GTEPS" << std::setw(12) << "GB/s" << std::setw(12) << "of HBM peak" << "\n";
    for (const std::string &path : datasets) {
        io::CSRMatrix<float> csr = io::load_csr_matrix_from_float_npz(path);
        io::util_round_csr_matrix_dim(csr, num_hbm_channels * pack_size, num_hbm_channels * pack_size);
        for (auto &x : csr.adj_data) x = 1.0 / csr.num_rows;
        io::CSCMatrix<float> csc = io::csr2csc(csr);
        for (const Case &sr : semirings) {
            module::SpMSpVModule<val_t, val_t, idx_val_t> spmspv(0);
            spmspv.set_semiring(sr.semiring);
            spmspv.set_mask_type(kNoMask);
            spmspv.set_target("hw");
            spmspv.set_up_runtime("");
            spmspv.load_and_format_matrix(csc);
            spmspv.send_matrix_host_to_device();
            spmspv.set_async_run(true);   // the timed loop enqueues its 20 runs back to back and synchronises once
            for (float sparsity : sparsities) {
                uint32_t nnz = uint32_t(std::floor(csc.num_cols * (1 - sparsity)));
                if (nnz == 0) nnz = 1;
                const uint32_t stride = csc.num_cols / nnz;
                aligned_sparse_vec_t vector(nnz + 1);
                vector[0] = {nnz, 0};
                for (uint32_t i = 0; i < nnz; i++) vector[i + 1] = {i * stride, float(rand() % 10) / 10};
                spmspv.send_vector_host_to_device(vector);
                spmspv.run();
                spmspv.get_runtime()->finish();
                // matrix term as measure_data_usage (8 B per non-zero of the active columns) + the lists
                const double out_nnz = spmspv.get_results_nnz();
                const double bytes = measure_data_usage(csc, vector) + 8.0 * nnz + 8.0 * out_nnz;
                const int num_runs = 20;
                auto t1 = std::chrono::high_resolution_clock::now();
                for (int i = 0; i < num_runs; i++) spmspv.run();
                spmspv.get_runtime()->finish();
                const double ms = seconds_since(t1) * 1e3 / num_runs;
                const double gbps = bytes / 1e6 / ms;
                table << std::setw(28) << path.substr(path.find_last_of('/') + 1) << std::setw(12) << sr.name << std::setw(10)
                      << sparsity << std::setw(10) << nnz << std::setw(12) << ms * 1e3 << std::setw(12)
                      << measure_data_usage(csc, vector) / 8 / 1e6 / ms << std::setw(12) << gbps << std::setw(12) << gbps / peak_gbs
                      << "\n";
            }
        }
    }
    std::cout << table.str();
    if (!logfile.empty()) std::ofstream(logfile) << "Kernel SpMSpV Benchmark, Target = B200\n" << table.str();
    return 0;
}

// SSSP: the flow of /root/reference/benchmark/bench_sssp.cpp (pull, then pull-push; GTEPS = nnz *
// iterations / seconds).
#include "bench_common.h"
#include "graphlily/app/sssp.h"

int main(int argc, char *argv[]) {
    BenchArgs args = parse_args(argc, argv, "bench_sssp [tuning ints... bitstream] <dataset.npz> <num_iterations>");
    const uint32_t num_iterations = args.ints_after.empty() ? 10 : uint32_t(args.ints_after[0]);
    graphlily::app::SSSP sssp(16, 0, 0, 0);
    sssp.set_target("hw");
    sssp.set_up_runtime("");
    sssp.load_and_format_matrix(args.dataset, true);
    std::cout << "finished load_and_format_matrix" << std::endl;
    sssp.send_matrix_host_to_device();
    const uint32_t source = 0;
    const double op_count = double(sssp.get_nnz()) * num_iterations;

    auto kernel_results = sssp.pull(source, num_iterations);
    auto t1 = std::chrono::high_resolution_clock::now();
    kernel_results = sssp.pull(source, num_iterations);
    double sec = seconds_since(t1);
    std::cout << "Pull average_time: " << sec * 1000 << " ms" << std::endl;
    std::cout << "Pull Compute THROUGHPUT = " << op_count / 1e9 / sec << " GTEPS" << std::endl;

    const float threshold = 0.01;
    kernel_results = sssp.pull_push(source, num_iterations, threshold);
    t1 = std::chrono::high_resolution_clock::now();
    kernel_results = sssp.pull_push(source, num_iterations, threshold);
    sec = seconds_since(t1);
    std::cout << "SpMSpV runs for " << sssp.get_push_iterations() << " iterations" << std::endl;
    std::cout << "Pull-Push average_time: " << sec * 1000 << " ms" << std::endl;
    std::cout << "Pull-Push Compute THROUGHPUT = " << op_count / 1e9 / sec << " GTEPS" << std::endl;
    size_t reached = 0;
    for (auto d : kernel_results) reached += d != graphlily::TropicalSemiring.zero;
    std::cout << "reached " << reached << " of " << kernel_results.size() << " vertices" << std::endl;
    return 0;
}

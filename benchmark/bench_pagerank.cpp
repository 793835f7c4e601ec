// PageRank: the flow of /root/reference/benchmark/bench_pagerank.cpp:33-66 (damping 0.9, 10
// iterations, GTEPS = nnz / seconds per iteration).
#include "bench_common.h"
#include "graphlily/app/pagerank.h"

int main(int argc, char *argv[]) {
    BenchArgs args = parse_args(argc, argv, "bench_pagerank [tuning ints... bitstream] <dataset.npz>");
    graphlily::app::PageRank pagerank(graphlily::num_hbm_channels, 0, 0);
    pagerank.set_target("hw");
    pagerank.set_up_runtime("");
    const float damping = 0.9;
    pagerank.load_and_format_matrix(args.dataset, damping, true);
    std::cout << "finished load_and_format_matrix" << std::endl;
    pagerank.send_matrix_host_to_device();
    const uint32_t num_iterations = 10;
    auto kernel_results = pagerank.pull(damping, num_iterations);
    auto t1 = std::chrono::high_resolution_clock::now();
    kernel_results = pagerank.pull(damping, num_iterations);
    const double sec = seconds_since(t1) / num_iterations;
    std::cout << "PageRank time for one iteration: " << sec * 1000 << " ms" << std::endl;
    std::cout << "PageRank Compute THROUGHPUT = " << double(pagerank.get_nnz()) / 1e9 / sec << " GTEPS" << std::endl;
    double sum = 0;
    for (auto r : kernel_results) sum += r;
    std::cout << "sum of ranks " << sum << std::endl;
    return 0;
}

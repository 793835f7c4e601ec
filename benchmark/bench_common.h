// Shared argv handling of the benchmark drivers.  The reference's drivers take positional
// arguments `<tuning ints...> <bitstream.xclbin> <dataset.npz> [iterations]`
// (/root/reference/benchmark/bench_*.cpp main()); here the tuning integers and the bitstream are
// optional and ignored, so both the reference's command lines and `bench_x <dataset.npz> [iters]` work.
#ifndef BENCH_COMMON_H_
#define BENCH_COMMON_H_
#include <chrono>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

struct BenchArgs {
    std::string dataset;
    std::vector<long> ints_before;  // tuning integers (ignored)
    std::vector<long> ints_after;   // e.g. the iteration count
};

inline bool ends_with(const std::string &s, const std::string &suffix) {
    return s.size() >= suffix.size() && s.compare(s.size() - suffix.size(), suffix.size(), suffix) == 0;
}

inline BenchArgs parse_args(int argc, char **argv, const char *usage) {
    BenchArgs a;
    for (int i = 1; i < argc; i++) {
        std::string s = argv[i];
        if (ends_with(s, ".npz")) a.dataset = s;
        else if (ends_with(s, ".xclbin")) continue;
        else (a.dataset.empty() ? a.ints_before : a.ints_after).push_back(strtol(argv[i], NULL, 10));
    }
    if (a.dataset.empty()) {
        std::cerr << "usage: " << usage << std::endl;
        exit(EXIT_FAILURE);
    }
    return a;
}

inline double seconds_since(std::chrono::high_resolution_clock::time_point t1) {
    return double(std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::high_resolution_clock::now() - t1).count()) / 1e6;
}
#endif

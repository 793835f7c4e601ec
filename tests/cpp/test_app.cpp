// GPU end-to-end tests of BFS / PageRank / SSSP against the oracle, in the shape of the reference's
// tests/test_app.cpp (:51-135): uniform_10K_10-shaped graph, source 0, 10 iterations, every mode,
// plus the 8-vertex chain (tests/test_data/line_8) whose answer is closed-form.  BFS and SSSP must
// be bit-exact, PageRank within 1e-5 relative.
#include "graphlily/app/bfs.h"
#include "graphlily/app/pagerank.h"
#include "graphlily/app/sssp.h"
#include "test_util.h"

using namespace graphlily;

static CSRMatrix<float> test_graph() { return uniform_csr(10000, 10, 21, 1.0f); }

static dense_t ref_bfs(CSRMatrix<float> m, uint32_t source, uint32_t iters) {
    graphlily::io::util_round_csr_matrix_dim(m, 128, 128);
    for (auto &x : m.adj_data) x = 1;
    dense_t d(m.num_rows);
    oracle_bfs(m.num_rows, m.adj_indptr.data(), m.adj_indices.data(), m.adj_data.data(), source, iters, d.data());
    return d;
}

TEST(BFS, AllModesFusedAndUnfused) {
    auto g = test_graph();
    dense_t ref = ref_bfs(g, 0, 10);
    for (bool fused : {true, false}) {
        app::BFS bfs(16, 1024, 512, 256);
        bfs.set_fused(fused);
        bfs.set_target("hw");
        bfs.set_up_runtime("ignored.xclbin");
        bfs.load_and_format_matrix(g, true);
        bfs.send_matrix_host_to_device();
        verify(ref, bfs.pull(0, 10), true);
        verify(ref, bfs.push(0, 10), true);
        verify(ref, bfs.pull_push(0, 10, 0.1f), true);
        EXPECT_TRUE(bfs.get_push_iterations() >= 1 && bfs.get_push_iterations() < 10);
        verify(ref, bfs.pull(0, 10), true);  // modules are reusable after a pull_push
        // the direction decided on the device (fused) / by the host (unfused): every threshold, other sources
        // and iteration counts, the recorded sequence replayed
        for (float thr : {0.001f, 0.1f, 1.1f}) verify(ref, bfs.pull_push(0, 10, thr), true);
        EXPECT_EQ(bfs.get_push_iterations(), 9u);   // threshold 1.1: push until the last level, which always pulls
        for (uint32_t src : {5u, 123u, 5u}) {
            dense_t r2 = ref_bfs(g, src, 4);
            verify(r2, bfs.pull_push(src, 4, 0.01f), true);
            verify(r2, bfs.push(src, 4), true);
        }
    }
}

TEST(BFS, Line8FromNpz) {
    const char *dir = getenv("GLB_TEST_DATA");
    std::string base = dir ? dir : "tests/golden";
    app::BFS bfs(16, 1024, 512, 256);
    bfs.set_up_runtime("ignored.xclbin");
    bfs.load_and_format_matrix(base + "/line_8_csr_float32.npz", true);
    bfs.send_matrix_host_to_device();
    dense_t d = bfs.pull_push(0, 10, 0.5f);
    ASSERT_EQ(d.size(), 128u);
    for (uint32_t k = 0; k < 8; k++) ASSERT_EQ(d[k], float(k + 1));  // vertex k is at level k
    for (uint32_t k = 8; k < 128; k++) ASSERT_EQ(d[k], 0.0f);
}

TEST(PageRank, Pull) {
    auto g = test_graph();
    CSRMatrix<float> m = g;
    graphlily::io::util_round_csr_matrix_dim(m, 128, 128);
    graphlily::io::util_normalize_csr_matrix_by_outdegree(m);
    const float damping = 0.9f;
    for (auto &x : m.adj_data) x = x * damping;
    dense_t ref(m.num_rows);
    oracle_pagerank(m.num_rows, m.adj_indptr.data(), m.adj_indices.data(), m.adj_data.data(), damping, 10, ref.data());
    for (bool fused : {true, false}) {
        app::PageRank pr(16, 1024, 256);
        pr.set_fused(fused);
        pr.set_up_runtime("ignored.xclbin");
        pr.load_and_format_matrix(g, damping, true);
        pr.send_matrix_host_to_device();
        verify(ref, pr.pull(damping, 10), false);
    }
}

TEST(SSSP, AllModes) {
    auto g = test_graph();
    CSRMatrix<float> m = g;
    app::detail::sssp_preprocess(m);
    graphlily::io::util_round_csr_matrix_dim(m, 128, 128);
    dense_t ref(m.num_rows);
    oracle_sssp(m.num_rows, m.adj_indptr.data(), m.adj_indices.data(), m.adj_data.data(), 0, 10, TropicalSemiring.zero,
                ref.data());
    for (bool fused : {true, false}) {
        app::SSSP sssp(16, 1024, 512, 256);
        sssp.set_fused(fused);
        sssp.set_up_runtime("ignored.xclbin");
        sssp.load_and_format_matrix(g, true);
        sssp.send_matrix_host_to_device();
        verify(ref, sssp.pull(0, 10), true);
        verify(ref, sssp.push(0, 10), true);
        verify(ref, sssp.pull_push(0, 10, 0.1f), true);
        for (float thr : {0.001f, 0.1f, 1.1f}) verify(ref, sssp.pull_push(0, 10, thr), true);
        verify(ref, sssp.push(0, 10), true);
        verify(ref, sssp.pull(0, 10), true);
    }
}

MINI_TEST_MAIN

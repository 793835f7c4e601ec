// The module classes on the other two value types of the reference (global.h:60-64), compiled twice:
//   -DGRAPHLILY_VAL_T_UNSIGNED   val_t = unsigned (UINT_INF)
//   -DGRAPHLILY_VAL_T_UFIXED     val_t = graphlily::ufixed_32_8 = ap_ufixed<32, 8, AP_RND, AP_SAT> of the shipped bitstream
// SpMVModule / SpMSpVModule / the apply modules instantiated on val_t, three semirings x three mask types, against
// the sequential model of those types (oracle/valtype_model.cpp: the loops of compute_reference_results with the ALUs
// of ufixed_pe_fwd.h) -- bit for bit, the arithmetic is exact.  Then BFS and SSSP (all three directions, unfused
// launch sequence) against the same graphs run through the model operator by operator.  "Parity unpinned": the
// reference ships no artefact that pins its device numerics; the model follows the documented ap_ufixed semantics.
#include "graphlily/app/bfs.h"
#include "graphlily/app/pagerank.h"
#include "graphlily/app/sssp.h"
#include "graphlily/module/add_scalar_vector_dense_module.h"
#include "graphlily/module/assign_vector_dense_module.h"
#include "graphlily/module/assign_vector_sparse_module.h"
#include "graphlily/module/spmspv_module.h"
#include "graphlily/module/spmv_module.h"
#include "test_util.h"

using namespace graphlily;
using VT = val_traits<val_t>;
static_assert(VT::id != GLB_VAL_F32, "compile with -DGRAPHLILY_VAL_T_UNSIGNED or -DGRAPHLILY_VAL_T_UFIXED");

extern "C" {
int vt_spmv(int val_type, uint32_t nrows, uint32_t ncols, const uint32_t *indptr, const uint32_t *indices, const uint32_t *data,
            int op, uint32_t zero, int mask_type, const uint32_t *x, const uint32_t *mask, uint32_t *y);
int vt_spmspv(int val_type, uint32_t nrows, uint32_t ncols, const uint32_t *indptr, const uint32_t *indices, const uint32_t *data,
              int op, uint32_t zero, int mask_type, const uint32_t *x_idx, const uint32_t *x_val, uint32_t x_nnz,
              const uint32_t *mask, uint32_t *y);
int vt_ewise_add(int val_type, const uint32_t *in, uint32_t *out, uint32_t len, uint32_t val);
int vt_assign_sparse_relax(int val_type, const uint32_t *m_idx, const uint32_t *m_val, uint32_t nnz, uint32_t *inout,
                           uint32_t *nf_idx, uint32_t *nf_val);
uint32_t vt_ufixed_from_double(double x);
}

static std::vector<uint32_t> words_of(const std::vector<float> &v) {
    std::vector<uint32_t> w(v.size());
    for (size_t i = 0; i < v.size(); i++) w[i] = VT::bits(VT::from_float(v[i]));
    return w;
}
static std::vector<uint32_t> words_of(const dense_t &v) {
    std::vector<uint32_t> w(v.size());
    for (size_t i = 0; i < v.size(); i++) w[i] = VT::bits(v[i]);
    return w;
}
static void expect_words(const std::vector<uint32_t> &want, const dense_t &got, const char *what) {
    EXPECT_EQ(want.size(), got.size());
    size_t bad = 0;
    for (size_t i = 0; i < want.size() && i < got.size(); i++) bad += VT::bits(got[i]) != want[i];
    if (bad) MT_FAIL_(false, "%s: %zu of %zu words differ from the model", what, bad, want.size());
}

// values around the interesting points of the type: small integers, fractions, numbers whose products saturate
static val_t sample_value(std::mt19937 &rng) {
    switch (rng() % 6) {
        case 0: return val_t(0);
        case 1: return val_t(1);
        case 2: return val_t(float(rng() % 7));
        case 3: return VT::from_float(float(rng() % 1000) / 64.0f);   // fractions (ufixed) / truncated (unsigned)
        case 4: return VT::from_float(200.0f + float(rng() % 50));    // products saturate Q8.24
        default: return VT::from_float(float(rng() % 16) / 4.0f);
    }
}

static CSRMatrix<float> value_matrix(uint32_t n, uint32_t seed) {
    CSRMatrix<float> m = skewed_csr(n, seed);
    std::mt19937 rng(seed + 1);
    for (auto &x : m.adj_data) x = (rng() % 4 == 0) ? 1.0f : float(rng() % 640) / 64.0f;   // exactly representable in Q8.24
    return m;
}

static const SemiringType kSemirings[3] = {ArithmeticSemiring, LogicalSemiring, TropicalSemiring};
static const MaskType kMasks[3] = {kNoMask, kMaskWriteToZero, kMaskWriteToOne};

TEST(ValTypes, HostValueClass) {
#if defined(GRAPHLILY_VAL_T_UFIXED)
    // the host class quantises like the software ap_ufixed of oracle/shim (AP_RND: half up; AP_SAT)
    for (double x : {0.0, 1.0, 0.5, 1.0 / 3.0, 255.0, 255.99999997, 256.0, 1e9, -3.0, 5.96046e-08 / 2, 5.96046e-08 * 1.5})
        EXPECT_EQ(ufixed_32_8(x).word(), vt_ufixed_from_double(x));
    EXPECT_EQ(TropicalSemiring.zero.word(), 255u << 24);
#else
    EXPECT_EQ(TropicalSemiring.zero, 0xffffffffu);
#endif
    EXPECT_EQ(sizeof(idx_val_t), sizeof(glb_idx_val_t));
}

TEST(ValTypes, SpMVModule) {
    auto rt = std::make_shared<Runtime>(0);
    CSRMatrix<float> m = value_matrix(3000, 5);
    const std::vector<uint32_t> data = words_of(m.adj_data);
    std::mt19937 rng(9);
    for (const SemiringType &sr : kSemirings)
        for (MaskType mt : kMasks) {
            module::SpMVModule<val_t, val_t> spmv(16, 1024, 256);
            spmv.set_runtime(rt);
            spmv.set_semiring(sr);
            spmv.set_mask_type(mt);
            spmv.load_and_format_matrix(m, false);
            spmv.send_matrix_host_to_device();
            dense_t x(m.num_cols), mask(m.num_rows);
            for (auto &v : x) v = (sr.op == kAddMin && rng() % 3 == 0) ? sr.zero : sample_value(rng);
            for (auto &v : mask) v = (rng() % 2) ? val_t(0) : sample_value(rng);
            spmv.send_vector_host_to_device(x);
            if (mt != kNoMask) spmv.send_mask_host_to_device(mask);
            spmv.run();
            std::vector<uint32_t> want(m.num_rows), xw = words_of(x), mw = words_of(mask);
            EXPECT_EQ(0, vt_spmv(VT::id, m.num_rows, m.num_cols, m.adj_indptr.data(), m.adj_indices.data(), data.data(), sr.op,
                                 VT::bits(sr.zero), mt, xw.data(), mw.data(), want.data()));
            expect_words(want, spmv.send_results_device_to_host(), "SpMVModule");
        }
}

TEST(ValTypes, SpMSpVModule) {
    auto rt = std::make_shared<Runtime>(0);
    CSRMatrix<float> m = value_matrix(3000, 6);
    CSCMatrix<float> csc = io::csr2csc(m);
    const std::vector<uint32_t> data = words_of(csc.adj_data);
    std::mt19937 rng(10);
    for (const SemiringType &sr : kSemirings)
        for (MaskType mt : kMasks)
            for (uint32_t stride : {3u, 40u}) {
                module::SpMSpVModule<val_t, val_t, idx_val_t> spmspv(256);
                spmspv.set_runtime(rt);
                spmspv.set_semiring(sr);
                spmspv.set_mask_type(mt);
                spmspv.load_and_format_matrix(csc);
                spmspv.send_matrix_host_to_device();
                sparse_t x;
                x.push_back(idx_val_t{0, val_t(0)});
                std::vector<uint32_t> xi, xv;
                for (uint32_t c = rng() % stride; c < m.num_cols; c += stride) {
                    val_t v = sample_value(rng);
                    if (v == sr.zero) v = val_t(1);
                    x.push_back(idx_val_t{c, v});
                    xi.push_back(c);
                    xv.push_back(VT::bits(v));
                }
                x[0].index = uint32_t(x.size() - 1);
                dense_t mask(m.num_rows);
                for (auto &v : mask) v = (rng() % 2) ? sr.zero : sample_value(rng);
                spmspv.send_vector_host_to_device(x);
                if (mt != kNoMask) spmspv.send_mask_host_to_device(mask);
                spmspv.run();
                std::vector<uint32_t> want(m.num_rows), mw = words_of(mask);
                EXPECT_EQ(0, vt_spmspv(VT::id, m.num_rows, m.num_cols, csc.adj_indptr.data(), csc.adj_indices.data(), data.data(), sr.op,
                                       VT::bits(sr.zero), mt, xi.data(), xv.data(), uint32_t(xi.size()), mw.data(), want.data()));
                sparse_t got = spmspv.send_results_device_to_host();
                dense_t dense = convert_sparse_vec_to_dense_vec<sparse_t, dense_t, val_t>(got, m.num_rows, sr.zero);
                expect_words(want, dense, "SpMSpVModule");
            }
}

TEST(ValTypes, ApplyModules) {
    auto rt = std::make_shared<Runtime>(0);
    std::mt19937 rng(11);
    const uint32_t n = 5000;
    dense_t in(n), mask(n), inout(n);
    for (uint32_t i = 0; i < n; i++) { in[i] = sample_value(rng); mask[i] = (rng() % 2) ? val_t(0) : sample_value(rng); inout[i] = sample_value(rng); }
    {   // eWiseAdd: saturating for Q8.24, modulo 2^32 for unsigned
        module::eWiseAddModule<val_t> add;
        add.set_runtime(rt);
        add.send_in_host_to_device(in);
        add.allocate_out_buf(n);
        const val_t val = VT::from_float(100.25f);
        add.run(n, val);
        std::vector<uint32_t> want(n), iw = words_of(in);
        vt_ewise_add(VT::id, iw.data(), want.data(), n, VT::bits(val));
        expect_words(want, add.send_out_device_to_host(), "eWiseAddModule");
    }
    for (MaskType mt : {kMaskWriteToZero, kMaskWriteToOne}) {   // assign_vector_dense_module.h:223-246 on words
        module::AssignVectorDenseModule<val_t> assign;
        assign.set_runtime(rt);
        assign.set_mask_type(mt);
        assign.send_mask_host_to_device(mask);
        assign.send_inout_host_to_device(inout);
        const val_t val = val_t(7);
        assign.run(n, val);
        std::vector<uint32_t> want = words_of(inout);
        for (uint32_t i = 0; i < n; i++) {
            const bool mz = VT::bits(mask[i]) == 0;
            if ((mt == kMaskWriteToZero && mz) || (mt == kMaskWriteToOne && !mz)) want[i] = VT::bits(val);
        }
        expect_words(want, assign.send_inout_device_to_host(), "AssignVectorDenseModule");
    }
    {   // sparse assign, both modes
        sparse_t list;
        list.push_back(idx_val_t{0, val_t(0)});
        for (uint32_t i = rng() % 7; i < n; i += 1 + rng() % 9) list.push_back(idx_val_t{i, sample_value(rng)});
        list[0].index = uint32_t(list.size() - 1);
        module::AssignVectorSparseModule<val_t, idx_val_t> bfs_mode(false);
        bfs_mode.set_runtime(rt);
        bfs_mode.send_mask_host_to_device(list);
        bfs_mode.send_inout_host_to_device(inout);
        bfs_mode.run(val_t(9));
        std::vector<uint32_t> want = words_of(inout);
        for (size_t k = 1; k < list.size(); k++) want[list[k].index] = VT::bits(val_t(9));
        expect_words(want, bfs_mode.send_inout_device_to_host(), "AssignVectorSparseModule (assign)");

        module::AssignVectorSparseModule<val_t, idx_val_t> sssp_mode(true);
        sssp_mode.set_runtime(rt);
        sssp_mode.send_mask_host_to_device(list);
        sssp_mode.send_inout_host_to_device(inout);
        sssp_mode.run();
        want = words_of(inout);
        std::set<uint32_t> improved;
        for (size_t k = 1; k < list.size(); k++)
            if (VT::bits(list[k].val) < want[list[k].index]) { want[list[k].index] = VT::bits(list[k].val); improved.insert(list[k].index); }
        expect_words(want, sssp_mode.send_inout_device_to_host(), "AssignVectorSparseModule (relax)");
        sparse_t nf = sssp_mode.send_new_frontier_device_to_host();
        std::set<uint32_t> got;
        for (size_t k = 1; k < nf.size(); k++) got.insert(nf[k].index);
        EXPECT_TRUE(got == improved);
    }
}

// the apps on val_t, unfused launch sequence (module run() calls only), against the same loops on the model
TEST(ValTypes, BfsAndSssp) {
    auto rt = std::make_shared<Runtime>(0);
    CSRMatrix<float> g = skewed_csr(4000, 21);
    for (auto &x : g.adj_data) x = 1.0f;
    {
        CSRMatrix<float> m = g;
        io::util_round_csr_matrix_dim(m, 128, 128);
        const uint32_t n = m.num_rows;
        const std::vector<uint32_t> data(m.adj_data.size(), VT::bits(val_t(1)));
        // model of bfs.h:106-126: input = SpMV(or-and, mask write-to-zero on the distance), distance[new] = level + 1
        std::vector<uint32_t> x(n, 0), dist(n, 0), y(n);
        const uint32_t src = 3, iters = 6;
        x[src] = VT::bits(val_t(1));
        dist[src] = VT::bits(val_t(1));
        for (uint32_t it = 1; it <= iters; it++) {
            vt_spmv(VT::id, n, n, m.adj_indptr.data(), m.adj_indices.data(), data.data(), kLogicalAndOr, 0, kMaskWriteToZero, x.data(),
                    dist.data(), y.data());
            for (uint32_t r = 0; r < n; r++)
                if (y[r] != 0) dist[r] = VT::bits(val_t(float(it + 1)));
            x = y;
        }
        app::BFS bfs(16, 1024, 512, 256);
        bfs.set_runtime(rt);
        bfs.set_fused(false);
        bfs.load_and_format_matrix(g, true);
        bfs.send_matrix_host_to_device();
        expect_words(dist, bfs.pull(src, iters), "BFS pull");
        expect_words(dist, bfs.push(src, iters), "BFS push");
        expect_words(dist, bfs.pull_push(src, iters, 0.05f), "BFS pull_push");
        bfs.set_fused(true);   // the pull levels run fused (one launch, the assign in the SpMV write-back); push stays unfused
        expect_words(dist, bfs.pull(src, iters), "BFS pull, fused");
        expect_words(dist, bfs.push(src, iters), "BFS push, fused requested");
        expect_words(dist, bfs.pull_push(src, iters, 0.05f), "BFS pull_push, fused requested");
    }
    {
        CSRMatrix<float> m = g;
        app::detail::sssp_preprocess(m);
        io::util_round_csr_matrix_dim(m, 128, 128);
        const uint32_t n = m.num_rows;
        const std::vector<uint32_t> data = words_of(m.adj_data);
        const uint32_t inf = VT::bits(TropicalSemiring.zero), src = 5, iters = 5;
        CSCMatrix<float> csc = io::csr2csc(m);
        const std::vector<uint32_t> cdata = words_of(csc.adj_data);
        // sssp.h:152-166: x = A (min.+) x, iterated
        auto pull_from = [&](std::vector<uint32_t> x, uint32_t count) {
            std::vector<uint32_t> y(n);
            for (uint32_t it = 0; it < count; it++) {
                vt_spmv(VT::id, n, n, m.adj_indptr.data(), m.adj_indices.data(), data.data(), kAddMin, inf, kNoMask, x.data(), nullptr, y.data());
                x = y;
            }
            return x;
        };
        // sssp.h:169-194: frontier -> SpMSpV -> relax the distance, the improved entries are the next frontier.
        // (The two directions are different algorithms on these types: UINT_INF + w wraps around in the pull direction,
        //  and rows the reference's preprocessing leaves without a zero diagonal forget their own value there.)
        auto push_levels = [&](uint32_t count) {
            std::vector<uint32_t> dist(n, inf), fi(1, src), fv(1, 0u), y(n), li, lv, nfi(n), nfv(n);
            dist[src] = 0;
            for (uint32_t it = 0; it < count; it++) {
                vt_spmspv(VT::id, n, n, csc.adj_indptr.data(), csc.adj_indices.data(), cdata.data(), kAddMin, inf, kNoMask, fi.data(),
                          fv.data(), uint32_t(fi.size()), nullptr, y.data());
                li.clear();
                lv.clear();
                for (uint32_t r = 0; r < n; r++)
                    if (y[r] != inf) { li.push_back(r); lv.push_back(y[r]); }
                const int cnt = vt_assign_sparse_relax(VT::id, li.data(), lv.data(), uint32_t(li.size()), dist.data(), nfi.data(), nfv.data());
                fi.assign(nfi.begin(), nfi.begin() + cnt);
                fv.assign(nfv.begin(), nfv.begin() + cnt);
            }
            return dist;
        };
        std::vector<uint32_t> x0(n, inf);
        x0[src] = 0;
        app::SSSP sssp(16, 1024, 512, 256);
        sssp.set_runtime(rt);
        sssp.set_fused(false);
        sssp.load_and_format_matrix(g, true);
        sssp.send_matrix_host_to_device();
        expect_words(pull_from(x0, iters), sssp.pull(src, iters), "SSSP pull");
        expect_words(push_levels(iters), sssp.push(src, iters), "SSSP push");
        // threshold 1.1: every level but the last pushes (sssp.h:214), the last one pulls from the distance vector
        expect_words(pull_from(push_levels(iters - 1), 1), sssp.pull_push(src, iters, 1.1f), "SSSP pull_push");
        EXPECT_EQ(sssp.get_push_iterations(), iters - 1);
        sssp.set_fused(true);
        expect_words(pull_from(x0, iters), sssp.pull(src, iters), "SSSP pull, fused");
        expect_words(push_levels(iters), sssp.push(src, iters), "SSSP push, fused requested");
    }
}

// PageRank on val_t (the shipped bitstream's configuration is Q8.24): rank = A' rank + teleport, saturating, both the
// reference's two-launch sequence and the fused one
TEST(ValTypes, PageRank) {
    auto rt = std::make_shared<Runtime>(0);
    CSRMatrix<float> g = uniform_csr(3000, 9, 31, 1.0f);
    const float damping = 0.9f;
    const uint32_t iters = 6;
    CSRMatrix<float> m = g;
    io::util_round_csr_matrix_dim(m, 128, 128);
    io::util_normalize_csr_matrix_by_outdegree(m);
    for (auto &x : m.adj_data) x = x * damping;
    const uint32_t n = m.num_rows;
    const std::vector<uint32_t> data = words_of(m.adj_data);
    const uint32_t teleport = VT::bits(val_t((1 - damping) / n));
    std::vector<uint32_t> x(n, VT::bits(val_t(float(1.0 / n)))), y(n);
    for (uint32_t it = 0; it < iters; it++) {
        vt_spmv(VT::id, n, n, m.adj_indptr.data(), m.adj_indices.data(), data.data(), kMulAdd, 0, kNoMask, x.data(), nullptr, y.data());
        vt_ewise_add(VT::id, y.data(), x.data(), n, teleport);
    }
    for (bool fused : {false, true}) {
        app::PageRank pr(16, 1024, 256);
        pr.set_runtime(rt);
        pr.set_fused(fused);
        pr.load_and_format_matrix(g, damping, true);
        pr.send_matrix_host_to_device();
        expect_words(x, pr.pull(damping, iters), fused ? "PageRank, fused" : "PageRank");
    }
#if defined(GRAPHLILY_VAL_T_UFIXED)
    double sum = 0;
    for (uint32_t w : x) sum += double(w) / 16777216.0;
    EXPECT_TRUE(sum > 0.5 && sum < 1.5);   // ranks still sum to about 1 in Q8.24 (quantisation of 1 / N and of the weights)
#endif
}

MINI_TEST_MAIN

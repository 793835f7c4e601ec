// CPU-only tests of the host IO layer (include/graphlily/io), restating the golden vectors of the
// reference's tests/test_io.cpp (create_csr_matrix :68-80, npz load :83-93, convert :96-107, csr2csc
// :110-118, round dim :121-130, normalise :133-140) and pinning sssp_preprocess to the oracle.
#include <unistd.h>

#include <fstream>
#include <iterator>

#include "graphlily/app/sssp.h"
#include <random>

#include "test_util.h"

static CSRMatrix<float> csr_matrix_1() {  // [[1,2,3,4],[5,0,6,0],[0,7,0,0],[0,0,0,8]]
    return graphlily::io::create_csr_matrix<float>(4, 4, {1, 2, 3, 4, 5, 6, 7, 8}, {0, 1, 2, 3, 0, 2, 1, 3}, {0, 4, 6, 7, 8});
}

TEST(DataLoader, CreateCSRMatrix) {
    auto m = csr_matrix_1();
    EXPECT_EQ(m.num_rows, 4u);
    EXPECT_EQ(m.adj_indptr.size(), 5u);
    EXPECT_EQ(m.adj_data[7], 8.0f);
}

TEST(DataLoader, LoadCSRMatrixFromFloatNpz) {
    const char *dir = getenv("GLB_TEST_DATA");
    std::string base = dir ? dir : "tests/golden";
    auto eye = graphlily::io::load_csr_matrix_from_float_npz(base + "/eye_10_csr_float32.npz");
    ASSERT_EQ(eye.num_rows, 10u);
    ASSERT_EQ(eye.num_cols, 10u);
    ASSERT_EQ(eye.adj_data.size(), 10u);
    for (uint32_t i = 0; i < 10; i++) {
        EXPECT_EQ(eye.adj_data[i], 1.0f);
        EXPECT_EQ(eye.adj_indices[i], i);
        EXPECT_EQ(eye.adj_indptr[i], i);
    }
    auto line = graphlily::io::load_csr_matrix_from_float_npz(base + "/line_8_csr_float32.npz");
    ASSERT_EQ(line.num_rows, 8u);
    ASSERT_EQ(line.adj_indices.size(), 7u);
    for (uint32_t i = 0; i < 7; i++) EXPECT_EQ(line.adj_indices[i], i);
    EXPECT_EQ(line.adj_indptr[1], 0u);
    EXPECT_EQ(line.adj_indptr[8], 7u);
}

// Files scipy.sparse.save_npz wrote (compressed or not, int32 or int64 index arrays): the Python test
// tests/test_cpp_host.py::test_cpp_npz_reader_on_scipy_files writes them with their expected sums into
// $GLB_NPZ_CASES/<name>.npz + <name>.txt ("rows cols nnz sum(indices) sum(indptr) sum(data)").
TEST(DataLoader, ScipyWrittenNpzVariants) {
    const char *dir = getenv("GLB_NPZ_CASES");
    if (!dir) return;  // only driven from the Python test
    for (const char *name : {"compressed_i32", "stored_i32", "compressed_i64", "stored_i64", "empty_rows", "no_nnz"}) {
        const std::string base = std::string(dir) + "/" + name;
        FILE *f = fopen((base + ".txt").c_str(), "r");
        ASSERT_TRUE(f != nullptr);
        unsigned long rows, cols, nnz;
        double s_indices, s_indptr, s_data;
        ASSERT_EQ(fscanf(f, "%lu %lu %lu %lf %lf %lf", &rows, &cols, &nnz, &s_indices, &s_indptr, &s_data), 6);
        fclose(f);
        auto m = graphlily::io::load_csr_matrix_from_float_npz(base + ".npz");
        EXPECT_EQ(m.num_rows, uint32_t(rows));
        EXPECT_EQ(m.num_cols, uint32_t(cols));
        EXPECT_EQ(m.adj_data.size(), size_t(nnz));
        EXPECT_EQ(m.adj_indices.size(), size_t(nnz));
        EXPECT_EQ(m.adj_indptr.size(), size_t(rows) + 1);
        double a = 0, b = 0, c = 0;
        for (auto v : m.adj_indices) a += v;
        for (auto v : m.adj_indptr) b += v;
        for (auto v : m.adj_data) c += v;
        EXPECT_EQ(a, s_indices);
        EXPECT_EQ(b, s_indptr);
        EXPECT_TRUE(std::fabs(c - s_data) <= 1e-9 * (1 + std::fabs(s_data)));
    }
}

// A damaged archive must end in a clean "npz: ..." exception (or load, if the damage hit a
// don't-care byte) -- never an out-of-bounds read: every truncation length and 1200 random byte
// corruptions of a stored and of a deflated archive.  Run under -fsanitize=address when changed.
TEST(DataLoader, MalformedNpzFailsCleanly) {
    if (std::getenv("GLB_NPZ_CASES")) return;   // the scipy-variants invocation of this binary: covered by the plain one
    const char *tmpdir = std::getenv("TMPDIR");
    const std::string base = std::string(tmpdir ? tmpdir : "/tmp") + "/glb_npz_fuzz_" + std::to_string(long(getpid()));
    std::vector<std::string> sources;
    {
        const std::string good = base + "_good.npz";
        graphlily::io::npz::Writer w(good);
        std::vector<int32_t> ix = {0, 2, 1, 3, 0}, ip = {0, 2, 4, 5};
        std::vector<float> d = {1, 2, 3, 4, 5};
        std::vector<int64_t> shape = {3, 4};
        w.add("indices", "<i4", {ix.size()}, ix.data(), ix.size() * 4);
        w.add("indptr", "<i4", {ip.size()}, ip.data(), ip.size() * 4);
        w.add("data", "<f4", {d.size()}, d.data(), d.size() * 4);
        w.add("shape", "<i8", {2}, shape.data(), 16);
        w.close();
        sources.push_back(good);
    }
    const char *data_dir = std::getenv("GLB_TEST_DATA");
    if (data_dir) sources.push_back(std::string(data_dir) + "/eye_10_csr_float32.npz");   // deflated members
    unsigned long long rng = 88172645463325252ull;
    auto next = [&rng] { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; };
    size_t clean_failures = 0, loads = 0;
    for (const std::string &src : sources) {
        std::ifstream f(src, std::ios::binary);
        std::vector<char> bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
        ASSERT_TRUE(bytes.size() > 100);
        const std::string bad = base + "_bad.npz";
        auto attempt = [&](const std::vector<char> &b) {
            { std::ofstream o(bad, std::ios::binary | std::ios::trunc); o.write(b.data(), std::streamsize(b.size())); }
            try { graphlily::io::npz::load(bad); loads++; }
            catch (const std::runtime_error &e) { EXPECT_TRUE(std::string(e.what()).compare(0, 4, "npz:") == 0); clean_failures++; }
        };
        for (size_t len = 0; len < bytes.size(); len += (bytes.size() > 2000 ? 7 : 1))
            attempt(std::vector<char>(bytes.begin(), bytes.begin() + long(len)));
        for (int k = 0; k < 1200; k++) {
            std::vector<char> b = bytes;
            const int flips = 1 + int(next() % 3);
            for (int j = 0; j < flips; j++) b[next() % b.size()] = char(next());
            attempt(b);
        }
        std::remove(bad.c_str());
    }
    std::remove((base + "_good.npz").c_str());
    EXPECT_TRUE(clean_failures > 500);
    std::printf("  %zu damaged archives rejected cleanly, %zu still loaded\n", clean_failures, loads);
}

TEST(DataLoader, Csr2Csc) {
    auto csc = graphlily::io::csr2csc(csr_matrix_1());
    std::vector<float> data = {1, 5, 2, 7, 3, 6, 4, 8};
    std::vector<uint32_t> indices = {0, 1, 0, 2, 0, 1, 0, 3}, indptr = {0, 2, 4, 6, 8};
    EXPECT_TRUE(csc.adj_data == data);
    EXPECT_TRUE(csc.adj_indices == indices);
    EXPECT_TRUE(csc.adj_indptr == indptr);
}

TEST(DataFormatter, RoundCSRMatrixDim) {
    auto m = csr_matrix_1();
    graphlily::io::util_round_csr_matrix_dim(m, 3, 5);
    EXPECT_EQ(m.num_rows, 6u);
    EXPECT_EQ(m.num_cols, 5u);
    std::vector<uint32_t> indptr = {0, 4, 6, 7, 8, 8, 8};
    EXPECT_TRUE(m.adj_indptr == indptr);
}

TEST(DataFormatter, NormalizeCSRMatrixByOutdegree) {
    auto m = csr_matrix_1();
    graphlily::io::util_normalize_csr_matrix_by_outdegree(m);
    std::vector<float> expect = {0.5f, 0.5f, 0.5f, 0.5f, 0.5f, 0.5f, 0.5f, 0.5f};
    EXPECT_TRUE(m.adj_data == expect);
}

TEST(DataFormatter, MatchesOracleOnRandomMatrices) {
    for (uint32_t seed = 1; seed <= 4; seed++) {
        auto m = skewed_csr(700 + 13 * seed, seed);
        // csr2csc
        auto csc = graphlily::io::csr2csc(m);
        std::vector<uint32_t> ip(m.num_cols + 1), ix(m.adj_indices.size());
        std::vector<float> d(m.adj_data.size());
        oracle_csr2csc(m.num_rows, m.num_cols, m.adj_indptr.data(), m.adj_indices.data(), m.adj_data.data(), ip.data(),
                       ix.data(), d.data());
        EXPECT_TRUE(csc.adj_indptr == ip && csc.adj_indices == ix && csc.adj_data == d);
        // normalise
        auto n1 = m;
        graphlily::io::util_normalize_csr_matrix_by_outdegree(n1);
        auto d2 = m.adj_data;
        oracle_normalize_outdegree(m.num_rows, m.num_cols, m.adj_indptr.data(), m.adj_indices.data(), d2.data());
        EXPECT_TRUE(n1.adj_data == d2);
        // SSSP preprocess (with and without existing diagonals)
        auto p = m;
        graphlily::app::detail::sssp_preprocess(p);
        std::vector<uint32_t> oip(m.num_rows + 1), oix(m.adj_indices.size() + m.num_rows);
        std::vector<float> od(oix.size());
        int64_t w = oracle_sssp_preprocess(m.num_rows, m.num_cols, m.adj_indptr.data(), m.adj_indices.data(),
                                           m.adj_data.data(), oip.data(), oix.data(), od.data());
        oix.resize(size_t(w));
        od.resize(size_t(w));
        EXPECT_TRUE(p.adj_indptr == oip && p.adj_indices == oix && p.adj_data == od);
    }
}

TEST(Global, ConvertSparseVecToDenseVec) {
    sparse_t s = {{2, 0}, {3, 7.0f}, {0, 5.0f}};
    auto d = graphlily::convert_sparse_vec_to_dense_vec<sparse_t, dense_t, float>(s, 5, 255.0f);
    dense_t expect = {5.0f, 255.0f, 255.0f, 7.0f, 255.0f};
    EXPECT_TRUE(d == expect);
}

// The host-side value class of the Q8.24 val_t (global.h) against the software ap_ufixed<32, 8, AP_RND, AP_SAT> of the
// oracle (oracle/shim/ap_fixed.h through libvaltype_model): same word for every double, and the traits of the three types.
extern "C" uint32_t vt_ufixed_from_double(double x);
extern "C" double vt_ufixed_to_double(uint32_t w);

TEST(Global, UfixedHostClassAndValTraits) {
    using graphlily::ufixed_32_8;
    std::mt19937_64 rng(7);
    std::uniform_real_distribution<double> small(0.0, 1.0), wide(-10.0, 300.0);
    for (int i = 0; i < 200000; i++) {
        const double x = (i % 3 == 0) ? small(rng) : (i % 3 == 1) ? wide(rng) : double(rng() % (1ull << 32)) / 16777216.0 + ((i % 2) ? 2.98023223876953125e-08 : 0.0);
        const uint32_t w = ufixed_32_8(x).word();
        if (w != vt_ufixed_from_double(x)) MT_FAIL_(true, "ufixed_32_8(%.17g) = %08x, ap_ufixed gives %08x", x, w, vt_ufixed_from_double(x));
    }
    for (uint32_t w : {0u, 1u, 1u << 24, 0x7fffffffu, 0xffffffffu})
        EXPECT_TRUE(ufixed_32_8::from_word(w).to_double() == vt_ufixed_to_double(w));
    EXPECT_EQ(ufixed_32_8(1).word(), 1u << 24);
    EXPECT_EQ(ufixed_32_8(-1.0).word(), 0u);
    EXPECT_EQ(ufixed_32_8(1e9).word(), 0xffffffffu);
    EXPECT_EQ(graphlily::val_traits<float>::id, GLB_VAL_F32);
    EXPECT_EQ(graphlily::val_traits<unsigned>::id, GLB_VAL_U32);
    EXPECT_EQ(graphlily::val_traits<ufixed_32_8>::id, GLB_VAL_UFIXED);
    EXPECT_EQ(graphlily::val_traits<unsigned>::from_float(3.9f), 3u);       // (val_t)x of the reference's formatter truncates
    EXPECT_EQ(graphlily::val_traits<unsigned>::from_float(-2.0f), 0u);
    float f = graphlily::val_container(ufixed_32_8(0.5));
    uint32_t bits;
    std::memcpy(&bits, &f, 4);
    EXPECT_EQ(bits, 1u << 23);
}

MINI_TEST_MAIN

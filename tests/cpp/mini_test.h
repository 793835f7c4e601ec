// Minimal test harness for the C++ host-mirror tests (gtest is not in this image).
// TEST(suite, name) { ... }  with EXPECT_* / ASSERT_* macros; main() runs everything, exit code =
// number of failed tests.  The reference's tests are gtest files of the same shape
// (/root/reference/tests/*.cpp).
#ifndef MINI_TEST_H_
#define MINI_TEST_H_
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

namespace mini_test {
struct Case { std::string name; std::function<void()> fn; };
inline std::vector<Case> &registry() { static std::vector<Case> r; return r; }
inline int &failures() { static int f = 0; return f; }
struct Registrar { Registrar(const char *n, std::function<void()> f) { registry().push_back({n, f}); } };
struct Abort {};
inline int run_all(int argc, char **argv) {
    int failed = 0;
    for (auto &c : registry()) {
        if (argc > 1 && c.name.find(argv[1]) == std::string::npos) continue;
        const int before = failures();
        try { c.fn(); } catch (Abort &) {} catch (std::exception &e) { std::printf("  exception: %s\n", e.what()); failures()++; }
        const bool ok = failures() == before;
        std::printf("[%s] %s\n", ok ? "  OK  " : "FAILED", c.name.c_str());
        if (!ok) failed++;
    }
    std::printf("%d test(s) failed\n", failed);
    return failed;
}
}  // namespace mini_test

#define TEST(suite, name)                                                                   \
    static void suite##_##name##_body();                                                    \
    static mini_test::Registrar suite##_##name##_reg(#suite "." #name, suite##_##name##_body); \
    static void suite##_##name##_body()
#define MT_FAIL_(fatal, ...)                                                     \
    do {                                                                         \
        std::printf("  %s:%d: ", __FILE__, __LINE__);                            \
        std::printf(__VA_ARGS__);                                                \
        std::printf("\n");                                                       \
        mini_test::failures()++;                                                 \
        if (fatal) throw mini_test::Abort();                                     \
    } while (0)
#define EXPECT_TRUE(c) do { if (!(c)) MT_FAIL_(false, "expected true: %s", #c); } while (0)
#define ASSERT_TRUE(c) do { if (!(c)) MT_FAIL_(true, "expected true: %s", #c); } while (0)
#define ASSERT_EQ(a, b) do { if (!((a) == (b))) MT_FAIL_(true, "expected %s == %s", #a, #b); } while (0)
#define EXPECT_EQ(a, b) do { if (!((a) == (b))) MT_FAIL_(false, "expected %s == %s", #a, #b); } while (0)
#define MINI_TEST_MAIN int main(int argc, char **argv) { return mini_test::run_all(argc, argv); }
#endif

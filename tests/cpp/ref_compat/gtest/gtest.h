// TEST INFRASTRUCTURE: the subset of googletest the reference's own test files use
// (/root/reference/tests/test_*.cpp: TEST, ASSERT_EQ / ASSERT_TRUE / ASSERT_LT ..., InitGoogleTest,
// RUN_ALL_TESTS) over tests/cpp/mini_test.h -- gtest is not in this image -- so that those files compile
// UNMODIFIED against include/graphlily.
#ifndef GLB_REF_COMPAT_GTEST_H_
#define GLB_REF_COMPAT_GTEST_H_
#include "mini_test.h"

#define ASSERT_FALSE(c) ASSERT_TRUE(!(c))
#define EXPECT_FALSE(c) EXPECT_TRUE(!(c))
#define ASSERT_NE(a, b) ASSERT_TRUE(!((a) == (b)))
#define EXPECT_NE(a, b) EXPECT_TRUE(!((a) == (b)))
#define ASSERT_LT(a, b) ASSERT_TRUE((a) < (b))
#define ASSERT_LE(a, b) ASSERT_TRUE((a) <= (b))
#define ASSERT_GT(a, b) ASSERT_TRUE((a) > (b))
#define ASSERT_GE(a, b) ASSERT_TRUE((a) >= (b))
#define EXPECT_LT(a, b) EXPECT_TRUE((a) < (b))
#define EXPECT_LE(a, b) EXPECT_TRUE((a) <= (b))
#define EXPECT_GT(a, b) EXPECT_TRUE((a) > (b))
#define EXPECT_GE(a, b) EXPECT_TRUE((a) >= (b))
#define ASSERT_FLOAT_EQ(a, b) ASSERT_TRUE(std::fabs(float(a) - float(b)) <= 4 * 1.1920929e-7f * std::fabs(float(b)))
#define ASSERT_NEAR(a, b, tol) ASSERT_TRUE(std::fabs(double(a) - double(b)) <= double(tol))

namespace testing {
inline void InitGoogleTest(int *, char **) {}
}  // namespace testing
inline int RUN_ALL_TESTS() { return mini_test::run_all(0, nullptr); }
#endif

// tiny assertion helpers in the style of the reference's tests (tests/compile_time_test.cpp:13-20): a plain
// main(), a failing check prints and aborts.
#pragma once
#include <cmath>
#include <cstdlib>
#include <iostream>
#define ASSERT_NEAR(a, b, tol)                                                                              \
    do {                                                                                                    \
        if (!(std::abs((a) - (b)) <= (tol))) {                                                              \
            std::cerr << __FILE__ << ":" << __LINE__ << " ASSERTION FAILED: " << #a << " (" << (a) << ") != " << #b \
                      << " (" << (b) << ") with tolerance " << (tol) << std::endl;                          \
            std::abort();                                                                                   \
        }                                                                                                   \
    } while (0)
#define ASSERT_TRUE(c)                                                                                      \
    do {                                                                                                    \
        if (!(c)) { std::cerr << __FILE__ << ":" << __LINE__ << " ASSERTION FAILED: " << #c << std::endl; std::abort(); } \
    } while (0)
#define ASSERT_THROWS(expr, Exc)                                                                            \
    do {                                                                                                    \
        bool thrown_ = false;                                                                               \
        try { (void)(expr); } catch (const Exc&) { thrown_ = true; }                                        \
        if (!thrown_) { std::cerr << __FILE__ << ":" << __LINE__ << " expected " << #Exc << " from " << #expr << std::endl; std::abort(); } \
    } while (0)

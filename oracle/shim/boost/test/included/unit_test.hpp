// Stand-in for Boost.Test's single-header runner, just enough for the reference's test/unit/icp.cpp to compile and
// run unmodified (BOOST_TEST_MODULE + BOOST_AUTO_TEST_CASE; its test cases contain no assertions).  Test infrastructure.
#pragma once
#include <cstdio>
#include <exception>
#include <vector>
namespace b200_boost_test {
struct Case { const char *name; void (*fn)(); };
inline std::vector<Case> &cases() { static std::vector<Case> c; return c; }
struct Reg { Reg(const char *n, void (*f)()) { Case c = {n, f}; cases().push_back(c); } };
} // namespace b200_boost_test
#define BOOST_AUTO_TEST_CASE(name)                                  \
    static void name();                                             \
    static ::b200_boost_test::Reg b200_reg_##name(#name, &name);    \
    static void name()
#ifdef BOOST_TEST_MODULE
int main()
{
    int failed = 0;
    for (size_t i = 0; i < b200_boost_test::cases().size(); ++i) {
        try {
            b200_boost_test::cases()[i].fn();
            fprintf(stderr, "[boost.test shim] %s: ok\n", b200_boost_test::cases()[i].name);
        } catch (const std::exception &e) {
            fprintf(stderr, "[boost.test shim] %s: exception: %s\n", b200_boost_test::cases()[i].name, e.what());
            ++failed;
        }
    }
    return failed ? 1 : 0;
}
#endif

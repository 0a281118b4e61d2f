// Rock base-types shim: base::TimeMark (a named stopwatch that the reference creates around icp.align, icpLocalization.cpp:203)
#pragma once
#include <base/Time.hpp>
#include <string>
namespace base {
struct TimeMark {
    std::string label;
    Time mark;
    explicit TimeMark(const std::string &l) : label(l), mark(Time::now()) {}
    Time passed() const { return Time::now() - mark; }
};
} // namespace base

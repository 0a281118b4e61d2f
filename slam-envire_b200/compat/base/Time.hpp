// Rock base-types shim: base::Time (microsecond timestamp) as used by icp/icpLocalization.hpp
#pragma once
#include <chrono>
#include <stdint.h>
namespace base {
struct Time {
    int64_t microseconds;
    Time() : microseconds(0) {}
    static Time fromMicroseconds(int64_t us) { Time t; t.microseconds = us; return t; }
    static Time fromSeconds(double s) { Time t; t.microseconds = (int64_t)(s * 1e6); return t; }
    static Time now()
    {
        return fromMicroseconds(std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::system_clock::now().time_since_epoch()).count());
    }
    double toSeconds() const { return microseconds * 1e-6; }
    bool operator<(const Time &o) const { return microseconds < o.microseconds; }
    bool operator==(const Time &o) const { return microseconds == o.microseconds; }
    Time operator-(const Time &o) const { return fromMicroseconds(microseconds - o.microseconds); }
};
} // namespace base

// Rock base-types shim: base::samples::LaserScan -- a planar range scan -- with the one method the ICP localisation
// wrapper needs (convertScanToPointCloud: valid ranges -> points in the x/y plane of the laser frame, mapped by a transform).
#pragma once
#include <Eigen/Geometry>
#include <base/Time.hpp>
#include <cmath>
#include <stdint.h>
#include <vector>
namespace base {
namespace samples {
struct LaserScan {
    base::Time time;
    double start_angle;        // rad, angle of ranges[0]
    double angular_resolution; // rad between consecutive ranges
    double speed;
    std::vector<uint32_t> ranges; // millimetres; values <= END_LASER_RANGE_ERRORS are error codes
    uint32_t minRange, maxRange;  // millimetres
    std::vector<float> remission;
    enum { END_LASER_RANGE_ERRORS = 5 };
    LaserScan() : start_angle(0), angular_resolution(0), speed(0), minRange(0), maxRange(0) {}
    bool isRangeValid(uint32_t r) const { return r > END_LASER_RANGE_ERRORS && r >= minRange && r <= maxRange; }
    std::vector<Eigen::Vector3d> convertScanToPointCloud(const Eigen::Affine3d &transform) const
    {
        std::vector<Eigen::Vector3d> pts;
        pts.reserve(ranges.size());
        for (size_t i = 0; i < ranges.size(); ++i) {
            if (!isRangeValid(ranges[i])) continue;
            const double d = ranges[i] * 1e-3, a = start_angle + (double)i * angular_resolution;
            pts.push_back(transform * Eigen::Vector3d(std::cos(a) * d, std::sin(a) * d, 0.0));
        }
        return pts;
    }
};
} // namespace samples
} // namespace base

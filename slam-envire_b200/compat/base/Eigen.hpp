// Rock base-types shim: base::Matrix3d etc. as used by icp/icpConfigurationTypes.hpp
#pragma once
#include <Eigen/Core>
namespace base {
typedef Eigen::Matrix3d Matrix3d;
typedef Eigen::Vector3d Vector3d;
typedef Eigen::Affine3d Affine3d;
} // namespace base

// ICPConfiguration and friends: field-for-field the configuration structs of the reference
// (icp/icpConfigurationTypes.hpp:13-72), so code that fills an ICPConfiguration keeps compiling.
// GPU-only knobs are NOT added here; they live behind icpb_set_option() (include/icp_b200.h).
#ifndef ICP_B200_ICP_CONFIGURATION_TYPES_HPP
#define ICP_B200_ICP_CONFIGURATION_TYPES_HPP

#include <string>
#include <base/Eigen.hpp>

namespace envire {
namespace icp {

enum COV_MODE { HARD_CODED, MSE_BASED, NO_COV };

/** how the covariance attached to an ICP result is produced */
struct ICPResultCovarianceConf {
    COV_MODE mode;
    base::Matrix3d cov_position;
    base::Matrix3d cov_orientation;
    ICPResultCovarianceConf() : mode(HARD_CODED) {}
};

/** how scan lines are accumulated into a point cloud before an ICP is triggered */
struct ICPPointCloudConfiguration {
    int lines_per_point_cloud;
    int min_line_advance;
    double min_distance_travelled_for_new_line;
    double min_rotation_for_new_line;
    double min_rotation_head_for_new_line;
};

/** parameters of Trimmed::align as used by ICPLocalization::doScanMatch (icp/icpLocalization.cpp:204) */
struct ICPConfiguration {
    double model_density;
    int max_iterations;
    double overlap;
    double min_mse;
    double min_mse_diff;
    double measurement_density;
    ICPResultCovarianceConf cov_conf;
    std::string environment_debug_path;
};

} // namespace icp
} // namespace envire
#endif

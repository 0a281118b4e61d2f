// icpLocalization.hpp -- the production caller of the ICP path, on top of TrimmedGPU.
//
// Mirrors envire::icp::ICPLocalization (reference icp/icpLocalization.hpp:27-125, icpLocalization.cpp): laser lines are
// accumulated into a point cloud (addScanLineToPointCloud / generatePointcloud), doScanMatch attaches the cloud under a
// new FrameNode, runs icp.align(...) with the ICPConfiguration, and fills an ICPResult incl. the covariance heuristics.
// Same class / method / field names.  Differences, all forced by scope: loadEnvironment takes an in-memory Environment
// (envire's on-disk serialization is out of scope) and Rock's base::samples::LaserScan / base::Time come from
// slam-envire_b200/compat when base-types is not installed.
#ifndef ICP_B200_ICP_LOCALIZATION_HPP
#define ICP_B200_ICP_LOCALIZATION_HPP

#include <base/Time.hpp>
#include <base/samples/LaserScan.hpp>

#include <algorithm>
#include <cmath>
#include <deque>
#include <iterator>
#include <memory>

#include "icpConfigurationTypes.hpp"
#include "icp_gpu.hpp"

namespace envire {
namespace icp {

class ICPInputData {
public:
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW
    base::Time pointCloudTime;
    envire::Pointcloud *pc;
    envire::FrameNode *fn;
    Eigen::Affine3d pc2World;
    ICPInputData() : pc(0), fn(0) {}
};

struct ICPResult {
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW
    base::Time time;
    int points;
    int pairs;
    double mse;
    Eigen::Affine3d from;
    Eigen::Affine3d to;
    std::vector<double> pairs_distance;
    Eigen::Matrix3d cov_position;
    Eigen::Matrix3d cov_orientation;
    ICPResult() : points(0), pairs(0), mse(0) {}
};

class LaserAndTransform {
public:
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW
    Eigen::Affine3d body2Odo;
    Eigen::Affine3d body2World;
    Eigen::Affine3d laser2Body;
    base::samples::LaserScan scan;
};

class ICPLocalization {
public:
    explicit ICPLocalization(int device = 0)
        : lastScanIndex(0), env(new envire::Environment()), icp(device), scanCount(0), pc(0), fn(0)
    {
    }

    void loadIcpConfiguration(ICPConfiguration conf) { this->conf = conf; }

    void initializePointCloud(ICPPointCloudConfiguration conf_point_cloud)
    {
        this->conf_point_cloud = conf_point_cloud;
        scansWithTransforms.clear();
        lastScanIndex = 0;
        scanCount = 0;
    }

    /** takes over `environment` (may be null: keep the current one) and adds every Pointcloud in it to the ICP model
     * (icpLocalization.cpp:114-130) */
    void loadEnvironment(envire::Environment *environment, double model_density)
    {
        if (environment) env.reset(environment);
        std::vector<envire::Pointcloud *> items = env->getItems<envire::Pointcloud>();
        for (size_t i = 0; i < items.size(); ++i) icp.addToModel(PointcloudAdapter(items[i], model_density));
    }

    /** a new scan line is kept only if the platform or the laser moved enough with respect to every stored line
     * (icpLocalization.cpp:49-100) */
    void addScanLineToPointCloud(Eigen::Affine3d body2Odo, Eigen::Affine3d body2World, Eigen::Affine3d laser2Body,
                                 const ::base::samples::LaserScan &scan_reading)
    {
        bool keep = true;
        for (size_t i = 0; keep && i < scansWithTransforms.size(); ++i) {
            const LaserAndTransform &old = scansWithTransforms[i];
            const Eigen::Affine3d motion(body2Odo.inverse() * old.body2Odo);
            Eigen::Vector3d y_new = laser2Body * Eigen::Vector3d::UnitY() - laser2Body.translation();
            y_new.normalize();
            // as in the reference: the newest stored line's axis, but the i-th line's origin
            Eigen::Vector3d y_old = scansWithTransforms.back().laser2Body * Eigen::Vector3d::UnitY() - old.laser2Body.translation();
            y_old.normalize();
            const double head = std::acos(y_new.dot(y_old));
            const double moved = motion.translation().norm();
            const double turned = std::fabs(Eigen::AngleAxisd(motion.rotation()).angle());
            keep = turned > conf_point_cloud.min_rotation_for_new_line || moved > conf_point_cloud.min_distance_travelled_for_new_line ||
                   head > conf_point_cloud.min_rotation_head_for_new_line;
        }
        if (keep) addLaserScan(body2Odo, body2World, laser2Body, scan_reading);
    }

    bool hasNewPointCloud()
    {
        return scanCount > conf_point_cloud.lines_per_point_cloud && (scanCount - lastScanIndex) > conf_point_cloud.min_line_advance;
    }

    /** all stored lines expressed in the CURRENT body frame, as one cloud (icpLocalization.cpp:132-168) */
    ICPInputData generatePointcloud()
    {
        envire::Pointcloud *cloud = new envire::Pointcloud();
        const LaserAndTransform &last = scansWithTransforms.back();
        curBody2World = last.body2World;
        const Eigen::Affine3d odo2World = curBody2World * last.body2Odo.inverse();
        for (std::deque<LaserAndTransform>::const_iterator it = scansWithTransforms.begin(); it != scansWithTransforms.end(); ++it) {
            const Eigen::Affine3d laser2World((odo2World * it->body2Odo) * it->laser2Body);
            const Eigen::Affine3d laser2CurBody(curBody2World.inverse() * laser2World);
            const std::vector<Eigen::Vector3d> line = it->scan.convertScanToPointCloud(laser2CurBody);
            std::copy(line.begin(), line.end(), std::back_inserter(cloud->vertices));
        }
        ICPInputData data;
        data.pc2World = curBody2World;
        data.pc = cloud;
        data.pointCloudTime = last.scan.time;
        lastScanIndex = scanCount;
        return data;
    }

    ICPInputData generatePointcloudSample(ICPInputData originalData, Eigen::Affine3d offset)
    {
        ICPInputData data;
        data.pc2World = offset * originalData.pc2World;
        data.pc2World.translation() = offset.translation() + originalData.pc2World.translation();
        data.pc = new envire::Pointcloud();
        data.pc->vertices = originalData.pc->vertices;
        data.pointCloudTime = originalData.pointCloudTime;
        return data;
    }

    /** icpLocalization.cpp:188-252 */
    ICPResult doScanMatch(struct ICPInputData &inputData, bool save)
    {
        pc = inputData.pc;
        env->attachItem(pc);
        fn = new envire::FrameNode(inputData.pc2World);
        env->attachItem(fn);
        env->addChild(env->getRootNode(), fn);
        pc->setFrameNode(fn);
        inputData.fn = fn;

        icp.align(PointcloudAdapter(pc, conf.measurement_density), conf.max_iterations, conf.min_mse, conf.min_mse_diff, conf.overlap);

        ICPResult result;
        result.time = inputData.pointCloudTime;
        result.points = (int)pc->vertices.size();
        result.from = inputData.pc2World;
        result.pairs = (int)icp.getPairs();
        if (result.pairs > 0) {
            result.mse = icp.getMeanSquareError();
            result.to = fn->getTransform();
            result.pairs_distance = icp.getPairsDistance();
            switch (conf.cov_conf.mode) {
            case HARD_CODED:
                result.cov_position = conf.cov_conf.cov_position;
                result.cov_orientation = conf.cov_conf.cov_orientation;
                break;
            case MSE_BASED: {
                // single precision on purpose: the reference computes these two factors as floats
                const float avgDist = std::max(1.0, conf.measurement_density) / 4.0;
                const float mseFactor = avgDist / std::sqrt(icp.getMeanSquareError());
                result.cov_position = Eigen::Matrix3d::Identity() * (1e-3 * 2.0 / std::pow(mseFactor * .5, 4));
                result.cov_orientation = Eigen::Matrix3d::Identity() * ((1.0 * M_PI / 180) / std::pow(mseFactor * .5, 4));
                break;
            }
            default:
                result.cov_position = Eigen::Matrix3d::Ones() * INFINITY;
                result.cov_orientation = Eigen::Matrix3d::Ones() * INFINITY;
                break;
            }
        }
        if (!save || conf.environment_debug_path.empty()) removeLastSavedPointCloud();
        return result;
    }

    void removeLastSavedPointCloud()
    {
        env->detachItem(pc);
        env->detachItem(fn);
    }

    envire::Environment *getEnvironment() { return env.get(); }
    TrimmedGPU &getICP() { return icp; }

private:
    void addLaserScan(Eigen::Affine3d body2Odo, Eigen::Affine3d body2World, Eigen::Affine3d laser2Body,
                      const ::base::samples::LaserScan &scan_reading)
    {
        LaserAndTransform lat;
        lat.scan = scan_reading;
        lat.body2Odo = body2Odo;
        lat.body2World = body2World;
        lat.laser2Body = laser2Body;
        scansWithTransforms.push_back(lat);
        while (scansWithTransforms.size() > static_cast<unsigned long>(conf_point_cloud.lines_per_point_cloud)) scansWithTransforms.pop_front();
        scanCount++;
    }

    int lastScanIndex;
    Eigen::Affine3d curBody2World;
    std::unique_ptr<envire::Environment> env;
    TrimmedGPU icp;
    ICPConfiguration conf;
    int scanCount;
    std::deque<LaserAndTransform> scansWithTransforms;
    ICPPointCloudConfiguration conf_point_cloud;
    envire::Pointcloud *pc;
    envire::FrameNode *fn;
};

} // namespace icp
} // namespace envire
#endif

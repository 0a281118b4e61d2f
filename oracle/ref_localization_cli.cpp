// Drives the REFERENCE's own envire::icp::ICPLocalization (icp/icpLocalization.hpp + icpLocalization.cpp, compiled
// unmodified where they lie under /root/reference by oracle/Makefile) on top of the drop-in: the build pre-includes
// slam-envire_b200/icp/icp_gpu.hpp with -DICP_B200_REPLACE_TRIMMEDKD (envire::icp::TrimmedKD = TrimmedGPU) and defines the
// include guard of the reference's icp.hpp, which is the one-line switch INTEGRATION.md asks a maintainer to make.
// Scenario, like a Rock component: the model is written as an envire scene directory (scene.yml + PLY),
// loadEnvironment(path) reads it back, scan lines -> generatePointcloud -> doScanMatch.  Prints one JSON object.
// TEST infrastructure (binary in oracle/_ref/, never shipped): the product holds no copy of the reference's caller.
//   ref_localization_cli <model.bin> <cov_mode 0|1|2> <measurement_density> <cloud_out.bin> <scene_dir>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <icpLocalization.hpp> // the reference's (-I$(REF)/icp)

static std::vector<double> read_bin(const char *path)
{
    FILE *f = fopen(path, "rb");
    if (!f) exit(2);
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<double> v(sz / sizeof(double));
    if (fread(v.data(), sizeof(double), v.size(), f) != v.size()) exit(2);
    fclose(f);
    return v;
}

int main(int argc, char **argv)
{
    if (argc < 4) return 2;
    using namespace envire::icp;
    const std::vector<double> mv = read_bin(argv[1]);
    if (argc < 6) return 2;
    {
        // the model as an envire scene on disk: root -> frame node (a non-trivial pose) -> point cloud in that frame
        envire::Environment env;
        const Eigen::Affine3d model2World(Eigen::Translation3d(0.5, -0.25, 0.0) * Eigen::AngleAxisd(0.1, Eigen::Vector3d::UnitZ()));
        envire::Pointcloud *model = new envire::Pointcloud();
        env.attachItem(model);
        envire::FrameNode *fm = new envire::FrameNode(model2World);
        env.addChild(env.getRootNode(), fm);
        model->setFrameNode(fm);
        const Eigen::Affine3d world2Model = model2World.inverse();
        for (size_t i = 0; i + 2 < mv.size(); i += 3) model->vertices.push_back(world2Model * Eigen::Vector3d(mv[i], mv[i + 1], mv[i + 2]));
        env.serialize(argv[5]);
    }

    ICPConfiguration conf;
    conf.model_density = 1.0;
    conf.max_iterations = 25;
    conf.overlap = 0.9;
    conf.min_mse = 1e-9;
    conf.min_mse_diff = 1e-11;
    conf.measurement_density = atof(argv[3]);
    conf.cov_conf.mode = (COV_MODE)atoi(argv[2]);
    conf.cov_conf.cov_position = Eigen::Matrix3d::Identity() * 0.01;
    conf.cov_conf.cov_orientation = Eigen::Matrix3d::Identity() * 0.02;
    ICPPointCloudConfiguration pcc;
    pcc.lines_per_point_cloud = 40;
    pcc.min_line_advance = 5;
    pcc.min_distance_travelled_for_new_line = 0.05;
    pcc.min_rotation_for_new_line = 0.05;
    pcc.min_rotation_head_for_new_line = 0.01;
    try {
        ICPLocalization loc;
        loc.loadIcpConfiguration(conf);
        loc.initializePointCloud(pcc);
        loc.loadEnvironment(argv[5], conf.model_density); // Environment::unserialize + addToModel per point cloud

        // a tilting planar laser 1 m above the ground of the model scene sweeps 60 lines while the body creeps forward;
        // the believed pose is off by a small rigid error that the scan match has to remove
        const Eigen::Affine3d body2World_true(Eigen::Translation3d(1.0, -2.0, 1.0) * Eigen::AngleAxisd(0.3, Eigen::Vector3d::UnitZ()));
        const Eigen::Affine3d err(Eigen::Translation3d(0.08, -0.05, 0.04) * Eigen::AngleAxisd(0.03, Eigen::Vector3d::UnitY()));
        int offered = 0;
        for (int line = 0; line < 60; ++line) {
            const double tilt = 0.35 + 0.012 * line; // head pitches down
            const Eigen::Affine3d laser2Body(Eigen::Translation3d(0.0, 0.0, 0.2) * Eigen::AngleAxisd(tilt, Eigen::Vector3d::UnitY()));
            base::samples::LaserScan scan;
            scan.time = base::Time::fromMicroseconds(1000 * line);
            scan.start_angle = -1.2;
            scan.angular_resolution = 2.4 / 359.0;
            scan.minRange = 100;
            scan.maxRange = 30000;
            // ray-cast against the ground plane z = 0 of the model (flat model => exact ranges)
            // the body creeps forward 6 cm per line (> min_distance_travelled_for_new_line, so every line is kept)
            const Eigen::Affine3d body2World_line = body2World_true * Eigen::Affine3d(Eigen::Translation3d(0.06 * line, 0.0, 0.0));
            const Eigen::Affine3d body2World_believed = body2World_line * err;
            const Eigen::Affine3d laser2World = body2World_line * laser2Body;
            for (int k = 0; k < 360; ++k) {
                const double a = scan.start_angle + k * scan.angular_resolution;
                const Eigen::Vector3d dir = laser2World.linear() * Eigen::Vector3d(std::cos(a), std::sin(a), 0.0);
                const Eigen::Vector3d o = laser2World.translation();
                uint32_t r = 1; // error code
                if (dir.z() < -1e-6) {
                    const double t = -o.z() / dir.z();
                    if (t > 0.1 && t < 30.0) r = (uint32_t)std::lround(t * 1000.0);
                }
                scan.ranges.push_back(r);
            }
            loc.addScanLineToPointCloud(body2World_believed, body2World_believed, laser2Body, scan);
            ++offered;
        }
        const bool has = loc.hasNewPointCloud();
        ICPInputData in = loc.generatePointcloud();
        const size_t n_pts = in.pc->vertices.size();
        // dump the cloud for the python side (oracle comparison)
        FILE *f = fopen(argc > 4 ? argv[4] : "/tmp/loc_cloud.bin", "wb");
        for (size_t i = 0; i < n_pts; ++i) fwrite(in.pc->vertices[i].data(), sizeof(double), 3, f);
        fclose(f);
        ICPResult res = loc.doScanMatch(in, false);
        const Eigen::Matrix4d F = res.from.matrix(), T = res.to.matrix();
        printf("{\"has_new\": %s, \"offered\": %d, \"points\": %d, \"pairs\": %d, \"mse\": %.17g, \"n_pd\": %zu, \"cov_pos00\": %.17g, "
               "\"cov_ori00\": %.17g, \"cov_pos01\": %.17g, \"detached\": %s, \"from\": [",
               has ? "true" : "false", offered, res.points, res.pairs, res.mse, res.pairs_distance.size(), res.cov_position(0, 0),
               res.cov_orientation(0, 0), res.cov_position(0, 1), in.pc->getEnvironment() == 0 ? "true" : "false");
        for (int r = 0; r < 4; ++r) printf("%s[%.17g, %.17g, %.17g, %.17g]", r ? ", " : "", F(r, 0), F(r, 1), F(r, 2), F(r, 3));
        printf("], \"to\": [");
        for (int r = 0; r < 4; ++r) printf("%s[%.17g, %.17g, %.17g, %.17g]", r ? ", " : "", T(r, 0), T(r, 1), T(r, 2), T(r, 3));
        printf("]}\n");
        delete in.pc;
        delete in.fn;
    } catch (const std::exception &e) {
        printf("{\"error\": \"%s\"}\n", e.what());
        return 1;
    }
    return 0;
}

// env_icp -- pairwise alignment of two point clouds with the B200 trimmed ICP ("env_icp-style" config of BASELINE.json).
//
// The reference tool of this name walks an envire environment on disk and has its ICP calls commented out
// (tools/env_icp.cpp:44-51,127-143); this driver does what those calls were for, through the kept API:
//   TrimmedGPU icp; icp.addToModel(PointcloudAdapter(model, density)); icp.align(PointcloudAdapter(scan, density), ...)
// usage: env_icp <model.ply> <scan.ply> [--density d] [--overlap o | --golden alpha beta eps] [--max-iter n]
//                [--min-mse x] [--min-mse-diff x] [--init t00 .. t33 (16 numbers, row-major)] [--out aligned.ply]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>

#include "icp/icp_gpu.hpp"
#include "io/ply.hpp"

int main(int argc, char **argv)
{
    if (argc < 3) {
        fprintf(stderr, "usage: env_icp <model.ply> <scan.ply> [--density d] [--overlap o | --golden a b eps] [--max-iter n] "
                        "[--min-mse x] [--min-mse-diff x] [--init 16 numbers] [--out aligned.ply]\n");
        return 2;
    }
    double density = 1.0, overlap = 0.8, min_mse = 1e-5, min_mse_diff = 1e-8, alpha = 0, beta = 0, eps = 0; // enview defaults
    size_t max_iter = 5;
    bool golden = false;
    Eigen::Matrix4d init = Eigen::Matrix4d::Identity();
    const char *out = 0;
    for (int i = 3; i < argc; ++i) {
        if (!strcmp(argv[i], "--density")) density = atof(argv[++i]);
        else if (!strcmp(argv[i], "--overlap")) overlap = atof(argv[++i]);
        else if (!strcmp(argv[i], "--golden")) { golden = true; alpha = atof(argv[++i]); beta = atof(argv[++i]); eps = atof(argv[++i]); }
        else if (!strcmp(argv[i], "--max-iter")) max_iter = (size_t)atoll(argv[++i]);
        else if (!strcmp(argv[i], "--min-mse")) min_mse = atof(argv[++i]);
        else if (!strcmp(argv[i], "--min-mse-diff")) min_mse_diff = atof(argv[++i]);
        else if (!strcmp(argv[i], "--out")) out = argv[++i];
        else if (!strcmp(argv[i], "--init")) { for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) init(r, c) = atof(argv[++i]); }
        else { fprintf(stderr, "unknown option %s\n", argv[i]); return 2; }
    }
    try {
        std::unique_ptr<envire::Environment> env(new envire::Environment());
        envire::Pointcloud *model = new envire::Pointcloud(), *scan = new envire::Pointcloud();
        env->attachItem(model);
        env->attachItem(scan);
        envire::FrameNode *fm = new envire::FrameNode(), *fs = new envire::FrameNode(Eigen::Affine3d(init));
        env->addChild(env->getRootNode(), fm);
        env->addChild(env->getRootNode(), fs);
        model->setFrameNode(fm);
        scan->setFrameNode(fs);
        envire::io::readPly(argv[1], *model);
        envire::io::readPly(argv[2], *scan);

        envire::icp::TrimmedGPU icp;
        const auto t0 = std::chrono::steady_clock::now();
        icp.addToModel(envire::icp::PointcloudAdapter(model, density));
        const auto t1 = std::chrono::steady_clock::now();
        if (golden) icp.align(envire::icp::PointcloudAdapter(scan, density), max_iter, min_mse, min_mse_diff, alpha, beta, eps);
        else icp.align(envire::icp::PointcloudAdapter(scan, density), max_iter, min_mse, min_mse_diff, overlap);
        const auto t2 = std::chrono::steady_clock::now();
        const Eigen::Matrix4d M = fs->getTransform().matrix();
        printf("{\"model_points\": %zu, \"scan_points\": %zu, \"iter\": %zu, \"pairs\": %zu, \"mse\": %.17g, \"overlap\": %.17g, "
               "\"add_model_ms\": %.3f, \"align_ms\": %.3f, \"scan_frame\": [",
               model->vertices.size(), scan->vertices.size(), icp.getNumIterations(), icp.getPairs(), icp.getMeanSquareError(),
               icp.getOverlap(), std::chrono::duration<double, std::milli>(t1 - t0).count(),
               std::chrono::duration<double, std::milli>(t2 - t1).count());
        for (int r = 0; r < 4; ++r) printf("%s[%.17g, %.17g, %.17g, %.17g]", r ? ", " : "", M(r, 0), M(r, 1), M(r, 2), M(r, 3));
        printf("]}\n");
        if (out) {
            envire::Pointcloud aligned;
            const Eigen::Affine3d T = fs->getTransform();
            for (size_t i = 0; i < scan->vertices.size(); ++i) aligned.vertices.push_back(T * scan->vertices[i]);
            envire::io::writePly(out, aligned);
        }
    } catch (const std::exception &e) {
        fprintf(stderr, "env_icp: %s\n", e.what());
        return 1;
    }
    return 0;
}

// env_slam6d -- sequential registration of a series of scans into a growing model ("env_slam6d-style" config of
// BASELINE.json).  The reference tool of this name only exports scans to / imports poses from the external 3DTK slam6D
// program (tools/env_slam6d.cpp:63-175); this driver performs the registration itself with the kept API: scan i is
// aligned against the model built from scans < i (TrimmedGPU::align), its FrameNode is updated by applyTransform, and it
// is then added to the model (addToModel accumulates, like FindPairsKDTree::addModel, icp/icp.hpp:231-240).
// usage: env_slam6d <dir> <n_scans> [--density d] [--overlap o] [--max-iter n] [--min-mse x] [--min-mse-diff x]
//                   [--growth-margin f]
//   <dir>/scanNNN.ply   scan in its sensor frame        <dir>/scanNNN.pose   initial sensor pose: 16 numbers, row-major 4x4
// writes <dir>/scanNNN.frames (refined 4x4, row-major) and prints one JSON summary.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <vector>

#include "icp/icp_gpu.hpp"
#include "io/ply.hpp"

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: env_slam6d <dir> <n_scans> [options]\n"); return 2; }
    const std::string dir = argv[1];
    const int n_scans = atoi(argv[2]);
    double density = 1.0, overlap = 0.8, min_mse = 1e-6, min_mse_diff = 1e-9, growth_margin = 0.0;
    size_t max_iter = 30;
    for (int i = 3; i < argc; ++i) {
        if (!strcmp(argv[i], "--density")) density = atof(argv[++i]);
        else if (!strcmp(argv[i], "--overlap")) overlap = atof(argv[++i]);
        else if (!strcmp(argv[i], "--max-iter")) max_iter = (size_t)atoll(argv[++i]);
        else if (!strcmp(argv[i], "--min-mse")) min_mse = atof(argv[++i]);
        else if (!strcmp(argv[i], "--min-mse-diff")) min_mse_diff = atof(argv[++i]);
        else if (!strcmp(argv[i], "--growth-margin")) growth_margin = atof(argv[++i]);
    }
    try {
        std::unique_ptr<envire::Environment> env(new envire::Environment());
        envire::icp::TrimmedGPU icp;
        // leave room around the first scans so that later scans are merged into the index instead of rebuilding it
        icp.setOption("growth_margin", growth_margin);
        double align_ms = 0, add_ms = 0;
        size_t total_iter = 0, model_points = 0;
        for (int s = 0; s < n_scans; ++s) {
            char name[64];
            snprintf(name, sizeof(name), "/scan%03d", s);
            envire::Pointcloud *pc = new envire::Pointcloud();
            env->attachItem(pc);
            envire::io::readPly(dir + name + ".ply", *pc);
            Eigen::Matrix4d P = Eigen::Matrix4d::Identity();
            std::ifstream pf((dir + name + ".pose").c_str());
            for (int r = 0; r < 4 && pf; ++r) for (int c = 0; c < 4; ++c) pf >> P(r, c);
            envire::FrameNode *fn = new envire::FrameNode(Eigen::Affine3d(P));
            env->addChild(env->getRootNode(), fn);
            pc->setFrameNode(fn);
            const auto t0 = std::chrono::steady_clock::now();
            if (s > 0) { // register against everything seen so far
                icp.align(envire::icp::PointcloudAdapter(pc, density), max_iter, min_mse, min_mse_diff, overlap);
                total_iter += icp.getNumIterations();
            }
            const auto t1 = std::chrono::steady_clock::now();
            icp.addToModel(envire::icp::PointcloudAdapter(pc, density)); // adapter picks up the refined FrameNode
            const auto t2 = std::chrono::steady_clock::now();
            align_ms += std::chrono::duration<double, std::milli>(t1 - t0).count();
            add_ms += std::chrono::duration<double, std::milli>(t2 - t1).count();
            model_points += pc->vertices.size();
            const Eigen::Matrix4d M = fn->getTransform().matrix();
            std::ofstream ff((dir + name + ".frames").c_str());
            ff.precision(17);
            for (int r = 0; r < 4; ++r) { for (int c = 0; c < 4; ++c) ff << M(r, c) << (c == 3 ? "\n" : " "); }
        }
        printf("{\"scans\": %d, \"model_points\": %zu, \"total_iterations\": %zu, \"align_ms_total\": %.3f, \"model_append_ms_total\": %.3f, "
               "\"ms_per_scan\": %.3f}\n", n_scans, model_points, total_iter, align_ms, add_ms, (align_ms + add_ms) / (n_scans > 0 ? n_scans : 1));
    } catch (const std::exception &e) {
        fprintf(stderr, "env_slam6d: %s\n", e.what());
        return 1;
    }
    return 0;
}

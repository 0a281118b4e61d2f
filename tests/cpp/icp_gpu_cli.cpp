// icp_gpu_cli.cpp -- runs envire::icp::TrimmedGPU (the C++ host mirror) the way the reference's own unit test
// drives TrimmedKD (test/unit/icp.cpp:95-159): Environment + two FrameNodes + two Pointclouds, addToModel, align.
// Prints one JSON object; tests/test_gpu_golden.py compares it with tests/golden/golden.json.
// ICP_B200_DEVICES=N in the environment: N GPUs in this one process (tests/test_multi_gpu.py).
//   icp_gpu_cli <model.bin> <src.bin> <T_src 16 col-major numbers> <density> fixed <max_iter> <min_mse> <min_mse_diff> <overlap>
//   icp_gpu_cli <model.bin> <src.bin> <T_src ...>                  <density> golden <max_iter> <min_mse> <min_mse_diff> <alpha> <beta> <eps>
// .bin = raw little-endian doubles, xyz triples.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#include "icp/icp_gpu.hpp"

static std::vector<double> read_bin(const char *path)
{
    FILE *f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
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
    if (argc < 24) { fprintf(stderr, "usage: see header\n"); return 2; }
    const std::vector<double> mv = read_bin(argv[1]), sv = read_bin(argv[2]);
    Eigen::Matrix4d Ts;
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) Ts(r, c) = atof(argv[3 + c * 4 + r]);
    const double density = atof(argv[19]);
    const bool golden = !strcmp(argv[20], "golden");
    const size_t max_iter = (size_t)atoll(argv[21]);
    const double min_mse = atof(argv[22]), min_mse_diff = atof(argv[23]);

    std::unique_ptr<envire::Environment> env(new envire::Environment());
    envire::FrameNode *fm1 = new envire::FrameNode(), *fm2 = new envire::FrameNode();
    env->addChild(env->getRootNode(), fm1);
    env->addChild(env->getRootNode(), fm2);
    envire::Pointcloud *mesh = new envire::Pointcloud(), *mesh2 = new envire::Pointcloud();
    env->attachItem(mesh);
    env->attachItem(mesh2);
    for (size_t i = 0; i + 2 < mv.size(); i += 3) mesh->vertices.push_back(Eigen::Vector3d(mv[i], mv[i + 1], mv[i + 2]));
    for (size_t i = 0; i + 2 < sv.size(); i += 3) mesh2->vertices.push_back(Eigen::Vector3d(sv[i], sv[i + 1], sv[i + 2]));
    mesh->setFrameNode(fm1);
    mesh2->setFrameNode(fm2);
    fm1->setTransform(Eigen::Affine3d::Identity());
    fm2->setTransform(Eigen::Affine3d(Ts));
    const size_t modified_before = env->modifiedCount();

    if (!strcmp(argv[20], "pairs")) {
        // envire::icp::Pairs mirror: add / trim / getTransform on the (model, source) points taken as pairs
        try {
            envire::icp::TrimmedGPU owner;
            envire::icp::Pairs pairs(owner.handle());
            const size_t n = std::min(mesh->vertices.size(), mesh2->vertices.size());
            for (size_t i = 0; i < n; ++i)
                pairs.add(mesh->vertices[i], mesh2->vertices[i], (mesh->vertices[i] - mesh2->vertices[i]).norm());
            const double tau = pairs.trim((size_t)atoll(argv[21]));
            const Eigen::Matrix4d M = pairs.getTransform().matrix();
            printf("{\"T\": [");
            for (int r = 0; r < 4; ++r) printf("%s[%.17g, %.17g, %.17g, %.17g]", r ? ", " : "", M(r, 0), M(r, 1), M(r, 2), M(r, 3));
            printf("], \"tau\": %.17g, \"mse\": %.17g, \"size\": %zu}\n", tau, pairs.getMeanSquareError(), pairs.size());
        } catch (const std::exception &e) {
            printf("{\"error\": \"%s\"}\n", e.what());
            return 1;
        }
        return 0;
    }
    if (!strcmp(argv[20], "ean_sine")) {
        // the reference's setSineWave fixture WITH normals and scan-edge attributes, through TrimmedGPUEAN
        // (what the commented-out icp_test1 of test/unit/icp.cpp:132-145 drives through TrimmedKDEAN)
        envire::Pointcloud *pcs[2] = {mesh, mesh2};
        for (int k = 0; k < 2; ++k) {
            envire::Pointcloud *pc = pcs[k];
            pc->vertices.clear();
            std::vector<envire::Pointcloud::vertex_attr> &attr = pc->getVertexData<envire::Pointcloud::vertex_attr>(envire::Pointcloud::VERTEX_ATTRIBUTES);
            std::vector<Eigen::Vector3d> &normal = pc->getVertexData<Eigen::Vector3d>(envire::Pointcloud::VERTEX_NORMAL);
            for (int i = -10; i <= 10; i++)
                for (int j = -10; j <= 10; j++) {
                    const bool edge = (i == -10 || j == -10 || i == 10 || j == 10);
                    const double x = i * .1, y = j * .1, d = std::sqrt(x * x + y * y);
                    const double zd = -.2 * 10.0 * std::sin(10.0 * d), nx = std::sqrt(1.0 / (1.0 + zd * zd)), ny = zd * nx;
                    pc->vertices.push_back(Eigen::Vector3d(x, y, .2 * std::cos(10.0 * d)));
                    attr.push_back(edge << 1);
                    normal.push_back(Eigen::AngleAxisd(std::atan2(y, x), Eigen::Vector3d::UnitZ()) * Eigen::Vector3d(-ny, 0, nx));
                }
        }
        try {
            envire::icp::TrimmedGPUEAN icp;
            icp.addToModel(envire::icp::PointcloudEdgeAndNormalAdapter(mesh, 1.0));
            icp.align(envire::icp::PointcloudEdgeAndNormalAdapter(mesh2, density), max_iter, min_mse, min_mse_diff, atof(argv[24]));
            const Eigen::Matrix4d M = fm2->getTransform().matrix();
            const std::vector<double> pd = icp.getPairsDistance();
            double pd_sum = 0;
            for (size_t i = 0; i < pd.size(); ++i) pd_sum += pd[i];
            printf("{\"fn_T\": [");
            for (int r = 0; r < 4; ++r) printf("%s[%.17g, %.17g, %.17g, %.17g]", r ? ", " : "", M(r, 0), M(r, 1), M(r, 2), M(r, 3));
            printf("], \"iter\": %zu, \"pairs\": %zu, \"mse\": %.17g, \"overlap\": %.17g, \"n_pairs_distance\": %zu, \"pd_sum\": %.17g}\n",
                   icp.getNumIterations(), icp.getPairs(), icp.getMeanSquareError(), icp.getOverlap(), pd.size(), pd_sum);
        } catch (const std::exception &e) {
            printf("{\"error\": \"%s\"}\n", e.what());
            return 1;
        }
        return 0;
    }
    try {
        // ICP_B200_DEVICES=N: TrimmedGPU(0, N) -- N GPUs in this one process (N = 1: the same entry points, one device)
        const char *nd = getenv("ICP_B200_DEVICES");
        std::unique_ptr<envire::icp::TrimmedGPU> owner(nd && atoi(nd) >= 1 ? new envire::icp::TrimmedGPU(0, atoi(nd))
                                                                           : new envire::icp::TrimmedGPU());
        envire::icp::TrimmedGPU &icp = *owner;
        icp.addToModel(envire::icp::PointcloudAdapter(mesh, 1.0));
        if (golden)
            icp.align(envire::icp::PointcloudAdapter(mesh2, density), max_iter, min_mse, min_mse_diff, atof(argv[24]), atof(argv[25]), atof(argv[26]));
        else
            icp.align(envire::icp::PointcloudAdapter(mesh2, density), max_iter, min_mse, min_mse_diff, atof(argv[24]));
        const Eigen::Matrix4d M = fm2->getTransform().matrix();
        const std::vector<double> pd = icp.getPairsDistance();
        double pd_sum = 0;
        bool sorted = true;
        for (size_t i = 0; i < pd.size(); ++i) { pd_sum += pd[i]; if (i && pd[i] < pd[i - 1]) sorted = false; }
        printf("{\"fn_T\": [");
        for (int r = 0; r < 4; ++r) printf("%s[%.17g, %.17g, %.17g, %.17g]", r ? ", " : "", M(r, 0), M(r, 1), M(r, 2), M(r, 3));
        printf("], \"iter\": %zu, \"pairs\": %zu, \"mse\": %.17g, \"mse_diff\": %.17g, \"overlap\": %.17g, \"n_pairs_distance\": %zu, "
               "\"pd_sum\": %.17g, \"pd_sorted\": %s, \"pd_last\": %.17g, \"set_transform_events\": %zu}\n",
               icp.getNumIterations(), icp.getPairs(), icp.getMeanSquareError(), icp.getMeanSquareErrorDiff(), icp.getOverlap(),
               pd.size(), pd_sum, sorted ? "true" : "false", pd.empty() ? 0.0 : pd.back(), env->modifiedCount() - modified_before);
    } catch (const std::exception &e) {
        printf("{\"error\": \"%s\"}\n", e.what());
        return 1;
    }
    return 0;
}

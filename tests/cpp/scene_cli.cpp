// Reads an envire scene directory (scene.yml + PLY files) with slam-envire_b200/io/scene.hpp and prints what it found as
// JSON; with a second argument the scene is written back there first and THAT copy is read (writer -> reader round trip).
//   scene_cli <scene_dir> [<rewrite_dir>]
#include <cstdio>
#include <envire/Core.hpp>

int main(int argc, char **argv)
{
    if (argc < 2) return 2;
    try {
        envire::Environment *env = envire::Environment::unserialize(argv[1]);
        if (argc > 2) {
            env->serialize(argv[2]);
            delete env;
            env = envire::Environment::unserialize(argv[2]);
        }
        printf("{\"root\": \"%s\", \"frames\": [", env->getRootNode()->getUniqueId().c_str());
        std::vector<envire::FrameNode *> fns = env->getItems<envire::FrameNode>();
        for (size_t i = 0; i < fns.size(); ++i) {
            const Eigen::Matrix4d m = fns[i]->getTransform().matrix();
            const Eigen::Matrix4d g = env->relativeTransform(fns[i], env->getRootNode()).matrix();
            printf("%s{\"id\": \"%s\", \"parent\": \"%s\", \"transform\": [", i ? ", " : "", fns[i]->getUniqueId().c_str(),
                   fns[i]->getParent() ? fns[i]->getParent()->getUniqueId().c_str() : "");
            for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) printf("%s%.17g", (r || c) ? ", " : "", m(r, c));
            printf("], \"to_root\": [");
            for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) printf("%s%.17g", (r || c) ? ", " : "", g(r, c));
            printf("]}");
        }
        printf("], \"clouds\": [");
        std::vector<envire::Pointcloud *> pcs = env->getItems<envire::Pointcloud>();
        for (size_t i = 0; i < pcs.size(); ++i) {
            double sx = 0, sy = 0, sz = 0;
            for (size_t k = 0; k < pcs[i]->vertices.size(); ++k) { sx += pcs[i]->vertices[k].x(); sy += pcs[i]->vertices[k].y(); sz += pcs[i]->vertices[k].z(); }
            printf("%s{\"id\": \"%s\", \"frame\": \"%s\", \"n\": %zu, \"sum\": [%.17g, %.17g, %.17g], \"normals\": %s}", i ? ", " : "",
                   pcs[i]->getUniqueId().c_str(), pcs[i]->getFrameNode() ? pcs[i]->getFrameNode()->getUniqueId().c_str() : "",
                   pcs[i]->vertices.size(), sx, sy, sz, pcs[i]->hasData(envire::Pointcloud::VERTEX_NORMAL) ? "true" : "false");
        }
        printf("]}\n");
        delete env;
    } catch (const std::exception &e) {
        printf("{\"error\": \"%s\"}\n", e.what());
        return 1;
    }
    return 0;
}

// scene.hpp -- minimal reader/writer for the scene directory envire serializes an Environment into
// (SURVEY.md section 8f, row f3): <dir>/scene.yml plus one PLY per point cloud.
//
// Format (reference src/core/Serialization.cpp:330-445 writes it with libyaml, :500-600 reads it back):
//   objects:                                   every EnvironmentItem, in id order
//    - class: 'envire::FrameNode'              EnvironmentItem::getClassName()
//      id: /1                                  EnvironmentItem::serialize (src/core/Environment.cpp:107-111); older files: `id: 1`
//      label: ''
//      transform: [r00, r01, r02, tx,  r10, ...]   16 numbers ROW-major (Serialization.cpp:176-189), may span lines
//    - class: 'envire::Pointcloud'             vertices in <class>_<id with '/' -> '_'>.ply (src/core/Layer.cpp:126-136,
//      id: /3                                  src/maps/Pointcloud.cpp:35-59), written by src/tools/PlyFile.cpp:173-246
//   links:
//    - {type: frameNodeTree, child: /2, parent: /1}       FrameNode tree
//    - {type: cartesianMapGraph, map: /3, node: /2}       which frame a map lives in
// Only what the ICP path needs is read: frame nodes with their transforms, the frame tree, point clouds with their
// frames.  Other item classes and link types (layer tree, operator graph) are skipped.  Host-side I/O only.
#ifndef ICP_B200_IO_SCENE_HPP
#define ICP_B200_IO_SCENE_HPP

#include <envire/core/FrameNode.hpp>
#include <envire/maps/Pointcloud.hpp>

#include <algorithm>
#include <cstdio>
#include <filesystem>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "ply.hpp"

namespace envire {
namespace io {
namespace detail {
typedef std::map<std::string, std::string> YamlMap; // one block-mapping entry of a sequence: key -> scalar / flow text

inline std::string yamlTrim(const std::string &s)
{
    size_t a = 0, b = s.size();
    while (a < b && (s[a] == ' ' || s[a] == '\t' || s[a] == '\r')) ++a;
    while (b > a && (s[b - 1] == ' ' || s[b - 1] == '\t' || s[b - 1] == '\r')) --b;
    std::string t = s.substr(a, b - a);
    if (t.size() >= 2 && ((t[0] == '\'' && t[t.size() - 1] == '\'') || (t[0] == '"' && t[t.size() - 1] == '"'))) t = t.substr(1, t.size() - 2);
    return t;
}
/** "key: value" pairs of a flow mapping "{a: 1, b: 2}" */
inline void yamlFlowMap(const std::string &text, YamlMap &out)
{
    std::string body = text.substr(1, text.rfind('}') - 1);
    std::istringstream is(body);
    std::string item;
    while (std::getline(is, item, ',')) {
        const size_t c = item.find(':');
        if (c != std::string::npos) out[yamlTrim(item.substr(0, c))] = yamlTrim(item.substr(c + 1));
    }
}
/** the subset of YAML the reference's emitter produces for a scene: two top-level keys holding sequences of mappings */
inline void parseSceneYaml(std::istream &is, std::vector<YamlMap> &objects, std::vector<YamlMap> &links)
{
    std::vector<YamlMap> *cur = 0;
    std::string line, pending_key;
    int depth = 0; // open '[' of a flow sequence that continues on the next lines
    while (std::getline(is, line)) {
        const size_t hash = line.find(" #");
        if (hash != std::string::npos && depth == 0) line = line.substr(0, hash);
        if (yamlTrim(line).empty() || yamlTrim(line) == "---") continue;
        if (depth > 0) { // continuation of a multi-line flow sequence
            cur->back()[pending_key] += " " + yamlTrim(line);
            for (size_t i = 0; i < line.size(); ++i) depth += (line[i] == '[') - (line[i] == ']');
            continue;
        }
        if (line[0] != ' ' && line[0] != '-') { // top-level key
            const std::string key = yamlTrim(line.substr(0, line.find(':')));
            cur = key == "objects" ? &objects : (key == "links" ? &links : 0);
            continue;
        }
        if (!cur) continue;
        std::string t = yamlTrim(line);
        if (t[0] == '-') { // a new sequence entry
            cur->push_back(YamlMap());
            t = yamlTrim(t.substr(1));
            if (t.empty()) continue;
            if (t[0] == '{') {
                yamlFlowMap(t, cur->back());
                continue;
            }
        }
        if (cur->empty()) continue;
        const size_t c = t.find(':');
        if (c == std::string::npos) continue;
        pending_key = yamlTrim(t.substr(0, c));
        const std::string val = yamlTrim(t.substr(c + 1));
        cur->back()[pending_key] = val;
        for (size_t i = 0; i < val.size(); ++i) depth += (val[i] == '[') - (val[i] == ']');
    }
}
inline std::string normId(const std::string &id) { return (!id.empty() && id[0] != '/') ? "/" + id : id; } // Environment.cpp:118-120
inline std::string mapFileName(const std::string &cls, const std::string &id)
{
    std::string u = id;
    std::replace(u.begin(), u.end(), '/', '_');
    return cls + u; // Layer::getMapFileName (src/core/Layer.cpp:131-136)
}
} // namespace detail

/** Load <dir>/scene.yml: frame nodes, their tree, point clouds (vertices from their PLY files) and their frames. */
inline Environment *readScene(const std::string &dir)
{
    using namespace detail;
    std::ifstream is((dir + "/scene.yml").c_str());
    if (!is) throw std::runtime_error("scene: cannot open " + dir + "/scene.yml");
    std::vector<YamlMap> objects, links;
    parseSceneYaml(is, objects, links);
    Environment *env = new Environment();
    std::map<std::string, FrameNode *> nodes;
    std::map<std::string, Pointcloud *> clouds;
    for (size_t i = 0; i < objects.size(); ++i) {
        YamlMap &o = objects[i];
        const std::string cls = o["class"], id = normId(o["id"]);
        if (cls == "envire::FrameNode") {
            std::string t = o["transform"];
            std::replace(t.begin(), t.end(), '[', ' ');
            std::replace(t.begin(), t.end(), ']', ' ');
            std::replace(t.begin(), t.end(), ',', ' ');
            std::istringstream ts(t);
            Eigen::Matrix4d m;
            for (int r = 0; r < 4; ++r)
                for (int c = 0; c < 4; ++c)
                    if (!(ts >> m(r, c))) throw std::runtime_error("scene: frame node " + id + " has no 4x4 transform");
            FrameNode *fn = new FrameNode(Transform(m));
            fn->setUniqueId(id);
            env->attachItem(fn);
            nodes[id] = fn;
        } else if (cls == "envire::Pointcloud" || cls == "envire::TriMesh") {
            Pointcloud *pc = new Pointcloud();
            pc->setUniqueId(id);
            readPly(dir + "/" + mapFileName(cls, id) + ".ply", *pc);
            env->attachItem(pc);
            clouds[id] = pc;
        }
    }
    std::map<FrameNode *, bool> is_child;
    for (size_t i = 0; i < links.size(); ++i) {
        YamlMap &l = links[i];
        if (l["type"] == "frameNodeTree") {
            FrameNode *c = nodes[normId(l["child"])], *p = nodes[normId(l["parent"])];
            if (!c || !p) throw std::runtime_error("scene: frameNodeTree link names an unknown frame node");
            env->addChild(p, c);
            is_child[c] = true;
        } else if (l["type"] == "cartesianMapGraph") {
            Pointcloud *m = clouds.count(normId(l["map"])) ? clouds[normId(l["map"])] : 0;
            FrameNode *n = nodes[normId(l["node"])];
            if (m && n) m->setFrameNode(n);
        }
    }
    // the frame node nobody is the child of is the root (the reference replaces its default root the same way)
    FrameNode *root = 0;
    for (std::map<std::string, FrameNode *>::iterator it = nodes.begin(); it != nodes.end(); ++it)
        if (!is_child.count(it->second) && !root) root = it->second;
    if (root) {
        for (std::map<std::string, FrameNode *>::iterator it = nodes.begin(); it != nodes.end(); ++it)
            if (it->second != root && !is_child.count(it->second)) env->addChild(root, it->second); // orphans hang off the root
        env->setRootNode(root);
    }
    for (std::map<std::string, Pointcloud *>::iterator it = clouds.begin(); it != clouds.end(); ++it)
        if (!it->second->getFrameNode()) it->second->setFrameNode(env->getRootNode());
    return env;
}

/** Write `env` as <dir>/scene.yml + PLY files in the layout above. */
inline void writeScene(Environment &env, const std::string &dir)
{
    using namespace detail;
    std::error_code ec;
    std::filesystem::create_directories(dir, ec); // like FileSerialization: the scene directory is created on demand
    std::ofstream os((dir + "/scene.yml").c_str());
    if (!os) throw std::runtime_error("scene: cannot write " + dir + "/scene.yml");
    os.precision(17);
    os << "objects:\n";
    const std::vector<EnvironmentItem *> &items = env.items();
    for (size_t i = 0; i < items.size(); ++i) {
        EnvironmentItem *it = items[i];
        if (FrameNode *fn = dynamic_cast<FrameNode *>(it)) {
            os << " - class: 'envire::FrameNode'\n   id: " << fn->getUniqueId() << "\n   label: ''\n   transform: [";
            const Eigen::Matrix4d m = fn->getTransform().matrix();
            for (int r = 0; r < 4; ++r)
                for (int c = 0; c < 4; ++c) os << m(r, c) << ((r == 3 && c == 3) ? "]\n" : (c == 3 ? ",\n               " : ", "));
        } else if (Pointcloud *pc = dynamic_cast<Pointcloud *>(it)) {
            os << " - class: 'envire::Pointcloud'\n   id: " << pc->getUniqueId() << "\n   label: ''\n";
            writePly(dir + "/" + mapFileName("envire::Pointcloud", pc->getUniqueId()) + ".ply", *pc);
        }
    }
    os << "links:\n";
    for (size_t i = 0; i < items.size(); ++i)
        if (FrameNode *fn = dynamic_cast<FrameNode *>(items[i]))
            if (fn->getParent())
                os << " - type: frameNodeTree\n   child: " << fn->getUniqueId() << "\n   parent: " << fn->getParent()->getUniqueId() << "\n";
    for (size_t i = 0; i < items.size(); ++i)
        if (Pointcloud *pc = dynamic_cast<Pointcloud *>(items[i]))
            if (pc->getFrameNode())
                os << " - type: cartesianMapGraph\n   map: " << pc->getUniqueId() << "\n   node: " << pc->getFrameNode()->getUniqueId() << "\n";
}
} // namespace io

inline Environment *Environment::unserialize(const std::string &dir) { return io::readScene(dir); }
inline void Environment::serialize(const std::string &dir) { io::writeScene(*this, dir); }
} // namespace envire
#endif

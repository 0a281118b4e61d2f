// envire-mini Pointcloud: the boundary type of the ICP path (reference src/maps/Pointcloud.hpp:10-85).
// Only what icp/icp.hpp reads: vertices, per-vertex metadata vectors, frame node, environment.
#pragma once
#include <envire/core/FrameNode.hpp>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace envire {
class Pointcloud : public CartesianMap {
public:
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW
    typedef int vertex_attr;
    static inline const std::string VERTEX_COLOR = "vertex_color";
    static inline const std::string VERTEX_NORMAL = "vertex_normal";
    static inline const std::string VERTEX_ATTRIBUTES = "vertex_attributes";
    static inline const std::string VERTEX_VARIANCE = "vertex_variance";
    enum attr_flag { SCAN_EDGE = 0x01 };

    std::vector<Eigen::Vector3d> vertices;
    std::string getClassName() const { return "envire::Pointcloud"; }
    /** a detached copy of the vertices and their metadata (reference: EnvironmentItem::clone via the copy constructor) */
    Pointcloud *clone() const
    {
        Pointcloud *c = new Pointcloud();
        c->vertices = vertices;
        c->data_ = data_; // shared, immutable in the ICP path
        return c;
    }

    template <typename T> std::vector<T> &getVertexData(const std::string &key)
    {
        std::shared_ptr<void> &slot = data_[key];
        if (!slot) slot = std::shared_ptr<void>(new std::vector<T>(), [](void *p) { delete static_cast<std::vector<T> *>(p); });
        return *static_cast<std::vector<T> *>(slot.get());
    }
    bool hasData(const std::string &key) const { return data_.count(key) != 0; }
private:
    std::map<std::string, std::shared_ptr<void> > data_;
};
typedef Pointcloud TriMesh; // the reference's tests fill TriMesh-typed constants on Pointclouds (test/unit/icp.cpp:23-27)
} // namespace envire

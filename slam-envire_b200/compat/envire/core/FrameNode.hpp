// envire-mini: the part of envire's core runtime the ICP boundary touches (4 calls).
//   Environment::relativeTransform  (reference src/core/Environment.cpp:744-782)
//   FrameNode::getTransform / setTransform -> itemModified (src/core/FrameNode.cpp:90-102)
//   Environment::{getRootNode, addChild, attachItem}, EnvironmentItem::getEnvironment
// Scene-graph bookkeeping, events, serialization and the operator graph are out of scope.
#pragma once
#include <Eigen/Geometry>
#include <map>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

namespace envire {
typedef Eigen::Affine3d Transform;
class Environment;
class FrameNode;

class EnvironmentItem {
public:
    virtual ~EnvironmentItem() {}
    Environment *getEnvironment() const { return env_; }
    bool isAttached() const { return env_ != 0; }
    /** "/<number>" once attached (reference src/core/Environment.cpp:263-290); "" before */
    const std::string &getUniqueId() const { return unique_id_; }
    void setUniqueId(const std::string &id) { unique_id_ = id; }
    virtual std::string getClassName() const { return "envire::EnvironmentItem"; }
protected:
    friend class Environment;
    Environment *env_ = 0;
    std::string unique_id_;
};

class FrameNode : public EnvironmentItem {
public:
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW
    FrameNode() : frame_(Transform::Identity()), parent_(0) {}
    explicit FrameNode(const Transform &t) : frame_(t), parent_(0) {}
    std::string getClassName() const { return "envire::FrameNode"; }
    bool isRoot() const { return parent_ == 0; }
    const FrameNode *getParent() const { return parent_; }
    FrameNode *getParent() { return parent_; }
    Transform const &getTransform() const { return frame_; }
    inline void setTransform(Transform const &transform); // fires itemModified when attached
    inline Transform relativeTransform(const FrameNode *to) const;
private:
    friend class Environment;
    Transform frame_;
    FrameNode *parent_;
};

class CartesianMap : public EnvironmentItem {
public:
    FrameNode *getFrameNode() const { return fn_; }
    void setFrameNode(FrameNode *fn) { fn_ = fn; }
private:
    FrameNode *fn_ = 0;
};

class Environment {
public:
    Environment() : root_(new FrameNode()), modified_(0), last_id_(0) { attachItem(root_); }
    ~Environment() { for (size_t i = 0; i < owned_.size(); ++i) delete owned_[i]; }
    FrameNode *getRootNode() { return root_; }
    void attachItem(EnvironmentItem *item)
    {
        item->env_ = this;
        if (item->unique_id_.empty()) { // like the reference: the next free number, prefixed with '/'
            item->unique_id_ = "/" + std::to_string(last_id_++);
        } else {
            const size_t slash = item->unique_id_.rfind('/');
            const long v = atol(item->unique_id_.c_str() + (slash == std::string::npos ? 0 : slash + 1));
            if (v >= last_id_) last_id_ = v + 1;
        }
        owned_.push_back(item);
    }
    const std::vector<EnvironmentItem *> &items() const { return owned_; }
    /** the scene written by Environment::serialize / the reference's FileSerialization: <dir>/scene.yml + one PLY per
     * point cloud (src/core/Serialization.cpp:330-445,500-600); defined in io/scene.hpp */
    static Environment *unserialize(const std::string &dir);
    void serialize(const std::string &dir);
    /** after loading a scene: the frame node without a parent is the root (the default root is dropped) */
    void setRootNode(FrameNode *fn)
    {
        if (fn == root_) return;
        for (size_t i = 0; i < owned_.size(); ++i)
            if (owned_[i] == root_) { owned_.erase(owned_.begin() + i); break; }
        delete root_;
        root_ = fn;
    }
    /** hands the item back to the caller (who then owns it), like Environment::detachItem */
    void detachItem(EnvironmentItem *item)
    {
        for (size_t i = 0; i < owned_.size(); ++i)
            if (owned_[i] == item) { owned_.erase(owned_.begin() + i); break; }
        item->env_ = 0;
    }
    template <class T> std::vector<T *> getItems()
    {
        std::vector<T *> out;
        for (size_t i = 0; i < owned_.size(); ++i)
            if (T *t = dynamic_cast<T *>(owned_[i])) out.push_back(t);
        return out;
    }
    void addChild(FrameNode *parent, FrameNode *child)
    {
        if (!child->isAttached()) attachItem(child);
        child->parent_ = parent;
    }
    void setFrameNode(CartesianMap *map, FrameNode *fn) { map->setFrameNode(fn); }
    // product of the transforms up to the common root: tg^-1 * fg
    Transform relativeTransform(const FrameNode *from, const FrameNode *to)
    {
        if (from == to) return Transform::Identity();
        Transform fg = Transform::Identity(), tg = Transform::Identity();
        const FrameNode *f = from, *t = to;
        while (!f->isRoot()) { fg = f->getTransform() * fg; f = f->getParent(); }
        while (!t->isRoot()) { tg = t->getTransform() * tg; t = t->getParent(); }
        if (f != t) throw std::runtime_error("relativeTransform: FrameNodes don't have a common root.");
        return tg.inverse() * fg;
    }
    void itemModified(EnvironmentItem *) { ++modified_; } // event fan-out is out of scope; counted for tests
    size_t modifiedCount() const { return modified_; }
private:
    FrameNode *root_;
    std::vector<EnvironmentItem *> owned_;
    size_t modified_;
    long last_id_;
};

inline void FrameNode::setTransform(Transform const &transform)
{
    frame_ = transform;
    if (isAttached()) env_->itemModified(this);
}
inline Transform FrameNode::relativeTransform(const FrameNode *to) const { return env_->relativeTransform(this, to); }
} // namespace envire

// Restated libkdtree++ (KDTree::KDTree<K,Val>) for the oracle/_ref build (TEST INFRASTRUCTURE).
// libkdtree++ is an external, unpinned dependency of the reference (manifest.xml:16, cmake/FindKDTree.cmake);
// this restates the behaviour its call sites rely on (icp/icp.hpp:237,248,256,260):
//   insert(v)            descend on coordinate (level % K); strictly-less goes left; never rebalanced
//   find_nearest(v,max)  (iterator, distance) of the exact nearest value with Euclidean distance <= max,
//                        distance = sqrt(sum (a[i]-b[i])^2); (end(), max) when there is none
// Equidistant candidates: lowest insertion order wins (unspecified in the library).
#pragma once
#include <cmath>
#include <cstddef>
#include <deque>
#include <utility>
#include <vector>

namespace KDTree {
template <size_t K, typename Val> class KDTree {
public:
    typedef double distance_type;
    typedef typename std::deque<Val>::const_iterator const_iterator;
    typedef const_iterator iterator;
    const_iterator end() const { return vals_.end(); }
    const_iterator begin() const { return vals_.begin(); }
    size_t size() const { return vals_.size(); }
    void clear() { vals_.clear(); left_.clear(); right_.clear(); depth_ = 0; }
    void insert(const Val &v)
    {
        const int id = (int)vals_.size();
        vals_.push_back(v);
        left_.push_back(-1);
        right_.push_back(-1);
        if (id == 0) { depth_ = 1; return; }
        int cur = 0;
        size_t level = 0;
        for (;;) {
            const size_t dim = level % K;
            int &next = (v[dim] < vals_[cur][dim]) ? left_[cur] : right_[cur];
            ++level;
            if (next < 0) { next = id; break; }
            cur = next;
        }
        if (level + 1 > depth_) depth_ = level + 1;
    }
    template <typename SearchVal> std::pair<const_iterator, distance_type> find_nearest(const SearchVal &q, distance_type max) const
    {
        if (vals_.empty()) return std::make_pair(end(), max);
        distance_type best = max;
        int best_id = -1;
        std::vector<std::pair<int, long> > stack;
        stack.reserve(2 * depth_ + 4);
        stack.push_back(std::make_pair(0, 0L));
        while (!stack.empty()) {
            int node = stack.back().first;
            long level = stack.back().second;
            stack.pop_back();
            if (level < 0) { // pending far side of `node`
                level = -(level + 1);
                const size_t dim = (size_t)level % K;
                const double diff = q[dim] - vals_[node][dim];
                if (!(std::fabs(diff) <= best)) continue;
                node = diff < 0.0 ? right_[node] : left_[node];
                ++level;
            }
            while (node >= 0) {
                const Val &nv = vals_[node];
                double d2 = 0.0;
                for (size_t i = 0; i < K; ++i) { const double d = q[i] - nv[i]; d2 += d * d; }
                const double d = std::sqrt(d2);
                if ((d < best || (d == best && (best_id < 0 || node < best_id))) && d <= max) { best = d; best_id = node; }
                const size_t dim = (size_t)level % K;
                const double diff = q[dim] - nv[dim];
                stack.push_back(std::make_pair(node, -(level + 1)));
                node = diff < 0.0 ? left_[node] : right_[node];
                ++level;
            }
        }
        if (best_id < 0) return std::make_pair(end(), max);
        return std::make_pair(vals_.begin() + best_id, best);
    }
    size_t index_of(const_iterator it) const { return (size_t)(it - vals_.begin()); }
private:
    std::deque<Val> vals_;
    std::vector<int> left_, right_;
    size_t depth_ = 0;
};
} // namespace KDTree

// ref_driver.cpp -- C entry points around the REFERENCE's own trimmed ICP (TEST INFRASTRUCTURE).
//
// Compiles /root/reference/icp/icp.cpp and icp/icp.hpp unmodified, where they lie, against restated
// third-party headers (oracle/shim: libkdtree++, Eigen::EigenSolver, boost::bind/function; and the
// envire-mini / mini-Eigen boundary headers of slam-envire_b200/compat).  Pairs::{add,trim,getTransform},
// PointcloudAdapter, FindPairsKDTree, GoldenBracket and Trimmed::{align,_align} are therefore the
// reference's code; only the libraries underneath are restated.  Used to pin icp_oracle.c and to
// generate tests/golden/*.json (tests/golden/make_golden.py).  Never linked into the product.
#include <cstring>
#include <vector>

#include "icp/icp.hpp"

namespace {
struct Ref {
    envire::Environment env;
    envire::icp::TrimmedKD icp;
    envire::icp::TrimmedKDEAN icp_ean;
};
envire::Pointcloud *make_cloud(Ref *r, const double *xyz, size_t n, const double T[16])
{
    envire::Pointcloud *pc = new envire::Pointcloud();
    r->env.attachItem(pc);
    Eigen::Matrix4d M;
    for (int c = 0; c < 4; ++c)
        for (int rr = 0; rr < 4; ++rr) M(rr, c) = T[c * 4 + rr];
    envire::FrameNode *fn = new envire::FrameNode(Eigen::Affine3d(M));
    r->env.addChild(r->env.getRootNode(), fn);
    pc->setFrameNode(fn);
    pc->vertices.reserve(n);
    for (size_t i = 0; i < n; ++i) pc->vertices.push_back(Eigen::Vector3d(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));
    return pc;
}
void add_ean_data(envire::Pointcloud *pc, const double *normals, const int *attrs, size_t n)
{
    std::vector<Eigen::Vector3d> &nv = pc->getVertexData<Eigen::Vector3d>(envire::Pointcloud::VERTEX_NORMAL);
    std::vector<envire::Pointcloud::vertex_attr> &av = pc->getVertexData<envire::Pointcloud::vertex_attr>(envire::Pointcloud::VERTEX_ATTRIBUTES);
    for (size_t i = 0; i < n; ++i) {
        nv.push_back(Eigen::Vector3d(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]));
        av.push_back(attrs[i]);
    }
}
template <class ICP> void fill_out_t(ICP &icp, envire::Pointcloud *pc, double *out, double *pd, size_t cap, size_t *n_pd)
{
    const Eigen::Matrix4d M = pc->getFrameNode()->getTransform().matrix();
    for (int c = 0; c < 4; ++c)
        for (int rr = 0; rr < 4; ++rr) out[c * 4 + rr] = M(rr, c);
    out[16] = (double)icp.getNumIterations();
    out[17] = (double)icp.getPairs();
    out[18] = icp.getMeanSquareError();
    out[19] = icp.getMeanSquareErrorDiff();
    out[20] = icp.getOverlap();
    const std::vector<double> v = icp.getPairsDistance();
    *n_pd = v.size();
    for (size_t i = 0; i < v.size() && i < cap; ++i) pd[i] = v[i];
}
void fill_out(Ref *r, envire::Pointcloud *pc, double *out, double *pd, size_t cap, size_t *n_pd)
{
    const Eigen::Matrix4d M = pc->getFrameNode()->getTransform().matrix();
    for (int c = 0; c < 4; ++c)
        for (int rr = 0; rr < 4; ++rr) out[c * 4 + rr] = M(rr, c);
    out[16] = (double)r->icp.getNumIterations();
    out[17] = (double)r->icp.getPairs();
    out[18] = r->icp.getMeanSquareError();
    out[19] = r->icp.getMeanSquareErrorDiff();
    out[20] = r->icp.getOverlap();
    const std::vector<double> v = r->icp.getPairsDistance();
    *n_pd = v.size();
    for (size_t i = 0; i < v.size() && i < cap; ++i) pd[i] = v[i];
}
} // namespace

extern "C" {
void *ref_new() { return new Ref(); }
void ref_free(void *h) { delete static_cast<Ref *>(h); }
void ref_clear_model(void *h) { static_cast<Ref *>(h)->icp.clearModel(); }
void ref_add_model(void *h, const double *xyz, size_t n, const double T[16], double density)
{
    Ref *r = static_cast<Ref *>(h);
    r->icp.addToModel(envire::icp::PointcloudAdapter(make_cloud(r, xyz, n, T), density));
}
// out: [0..15] the source cloud's FrameNode transform after applyTransform (col-major), [16] iter, [17] pairs,
// [18] mse, [19] mse_diff, [20] overlap
void ref_align(void *h, const double *xyz, size_t n, const double T[16], double density, size_t max_iter, double min_mse,
               double min_mse_diff, double overlap, double *out, double *pd, size_t cap, size_t *n_pd)
{
    Ref *r = static_cast<Ref *>(h);
    envire::Pointcloud *pc = make_cloud(r, xyz, n, T);
    r->icp.align(envire::icp::PointcloudAdapter(pc, density), max_iter, min_mse, min_mse_diff, overlap);
    fill_out(r, pc, out, pd, cap, n_pd);
}
void ref_align_golden(void *h, const double *xyz, size_t n, const double T[16], double density, size_t max_iter,
                      double min_mse, double min_mse_diff, double alpha, double beta, double eps, double *out,
                      double *pd, size_t cap, size_t *n_pd)
{
    Ref *r = static_cast<Ref *>(h);
    envire::Pointcloud *pc = make_cloud(r, xyz, n, T);
    r->icp.align(envire::icp::PointcloudAdapter(pc, density), max_iter, min_mse, min_mse_diff, alpha, beta, eps);
    fill_out(r, pc, out, pd, cap, n_pd);
}
// the same through TrimmedKDEAN = Trimmed<PointcloudEdgeAndNormalAdapter, FindPairsKDEAN> (icp/icp.hpp:513-521)
void ref_add_model_ean(void *h, const double *xyz, const double *normals, const int *attrs, size_t n, const double T[16],
                       double density)
{
    Ref *r = static_cast<Ref *>(h);
    envire::Pointcloud *pc = make_cloud(r, xyz, n, T);
    add_ean_data(pc, normals, attrs, n);
    r->icp_ean.addToModel(envire::icp::PointcloudEdgeAndNormalAdapter(pc, density));
}
void ref_align_ean(void *h, const double *xyz, const double *normals, const int *attrs, size_t n, const double T[16],
                   double density, size_t max_iter, double min_mse, double min_mse_diff, double overlap, double *out,
                   double *pd, size_t cap, size_t *n_pd)
{
    Ref *r = static_cast<Ref *>(h);
    envire::Pointcloud *pc = make_cloud(r, xyz, n, T);
    add_ean_data(pc, normals, attrs, n);
    r->icp_ean.align(envire::icp::PointcloudEdgeAndNormalAdapter(pc, density), max_iter, min_mse, min_mse_diff, overlap);
    fill_out_t(r->icp_ean, pc, out, pd, cap, n_pd);
}
// Pairs::add x n, trim(n_po), getTransform() on caller-supplied pairs.  Returns the kept count (or -1 on throw).
long ref_pairs(const double *x, const double *p, const double *d, size_t n, size_t n_po, double *T16, double *mse,
               double *tau)
{
    envire::icp::Pairs pairs;
    for (size_t i = 0; i < n; ++i)
        pairs.add(Eigen::Vector3d(x[3 * i], x[3 * i + 1], x[3 * i + 2]), Eigen::Vector3d(p[3 * i], p[3 * i + 1], p[3 * i + 2]),
                  d[i]);
    *tau = pairs.trim(n_po);
    try {
        const Eigen::Matrix4d M = pairs.getTransform().matrix();
        for (int c = 0; c < 4; ++c)
            for (int rr = 0; rr < 4; ++rr) T16[c * 4 + rr] = M(rr, c);
        *mse = pairs.getMeanSquareError();
    } catch (const std::runtime_error &) {
        return -1;
    }
    return (long)pairs.size();
}
}

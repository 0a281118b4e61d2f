// icp_gpu.hpp -- host mirror of envire's icp/ interface over the B200 C ABI.
//
// envire::icp::TrimmedGPU has the public surface of envire::icp::TrimmedKD
// (reference icp/icp.hpp:385-439, typedef at :522): addToModel, clearModel, align (fixed overlap), align (golden
// section over the overlap), getNumIterations, getMeanSquareError, getMeanSquareErrorDiff, getOverlap, getPairs,
// getPairsDistance -- same argument order, same getter semantics, adapters passed by value, exceptions where the
// reference throws.  The compute (model index, correspondences, trim, pose, loop) runs in libicp_b200.so
// (include/icp_b200.h); what stays on the host is exactly what the reference also does once per align on the host:
// the adapter's frame lookup (Environment::relativeTransform), applyTransform -> FrameNode::setTransform, and the
// golden-section driver.  There is no CPU fallback: constructing a TrimmedGPU without a CUDA device throws.
//
// Build against real Eigen3 + envire when they are installed; otherwise add slam-envire_b200/compat to the include
// path (mini-Eigen + envire-mini with the same names).
#ifndef ICP_B200_ICP_GPU_HPP
#define ICP_B200_ICP_GPU_HPP

#include <Eigen/Core>
#include <Eigen/Geometry>
#include <envire/core/EnvironmentItem.hpp>
#include <envire/core/FrameNode.hpp>
#include <envire/maps/Pointcloud.hpp>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <limits>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/icp_b200.h"

namespace envire {
namespace icp {

/** envire::icp::Pairs (icp/icp.hpp:32-76, icp/icp.cpp:12-119): a set of point pairs with their distances; trim() keeps the
 * n_po closest and returns the largest kept distance, getTransform() returns the reference's (non-Horn) rigid fit of p onto
 * x.  Same public members (x, p, mse) and methods; the sort / sums / pose solve run on the device through
 * icpb_trim / icpb_pairs_transform (used e.g. by icp/ransac.hpp:40-50 on 3-point samples). */
namespace detail {
/** one context per process for Pairs objects that are default-constructed, as the reference's callers do
 * (icp/ransac.hpp:40); created on first use on device 0 */
inline icpb_ctx *sharedPairsContext()
{
    struct Holder {
        icpb_ctx *ctx;
        Holder() : ctx(0)
        {
            if (icpb_create(&ctx, 0) != ICPB_OK) {
                if (ctx) icpb_destroy(ctx);
                ctx = 0;
                throw std::runtime_error("envire::icp::Pairs: no CUDA device (there is no CPU fallback)");
            }
        }
        ~Holder() { if (ctx) icpb_destroy(ctx); }
    };
    static Holder h;
    return h.ctx;
}
} // namespace detail

class Pairs {
public:
    const static unsigned int MIN_PAIRS = 3;
    Pairs() : mse(0), ctx_(0), n_po_(static_cast<size_t>(-1)) {}
    explicit Pairs(icpb_ctx *ctx) : mse(0), ctx_(ctx), n_po_(static_cast<size_t>(-1)) {}
    void add(const Eigen::Vector3d &a, const Eigen::Vector3d &b, double dist)
    {
        x.push_back(a);
        p.push_back(b);
        dist_.push_back(dist);
    }
    double trim(size_t n_po)
    {
        n_po_ = n_po;
        double tau = std::numeric_limits<double>::quiet_NaN();
        size_t kept = 0, ties = 0;
        if (!ctx_) ctx_ = detail::sharedPairsContext();
        if (icpb_trim(ctx_, dist_.data(), dist_.size(), n_po, &tau, &kept, &ties) != ICPB_OK) throw std::runtime_error(icpb_last_error(ctx_));
        kept_ = kept;
        return tau;
    }
    Eigen::Affine3d getTransform()
    {
        if (size() < MIN_PAIRS) throw std::runtime_error("not enough pairs to get transform");
        double T[16], tau;
        size_t kept;
        if (!ctx_) ctx_ = detail::sharedPairsContext();
        const int rc = icpb_pairs_transform(ctx_, x.empty() ? 0 : x[0].data(), p.empty() ? 0 : p[0].data(), dist_.data(), dist_.size(),
                                            n_po_ < dist_.size() ? n_po_ : dist_.size(), T, &mse, &tau, &kept);
        if (rc == ICPB_ERR_NOT_ENOUGH_PAIRS) throw std::runtime_error("not enough pairs to get transform");
        if (rc != ICPB_OK) throw std::runtime_error(icpb_last_error(ctx_));
        Eigen::Matrix4d m;
        for (int c = 0; c < 4; ++c)
            for (int r = 0; r < 4; ++r) m(r, c) = T[c * 4 + r];
        return Eigen::Affine3d(m);
    }
    double getMeanSquareError() const { return mse; }
    size_t size() const { return n_po_ < dist_.size() ? kept_ : dist_.size(); }
    void clear()
    {
        x.clear();
        p.clear();
        dist_.clear();
        n_po_ = static_cast<size_t>(-1);
    }

    std::vector<Eigen::Vector3d> x, p;
    double mse;

private:
    icpb_ctx *ctx_;
    std::vector<double> dist_;
    size_t n_po_, kept_ = 0;
};

/** Adapter concept of Trimmed<_Adapter,_FindPairs> (icp/icp.hpp:99-171) for envire::Pointcloud.  The iteration
 * interface (reset/hasNext/next/size) is kept for code that walks an adapter itself; TrimmedGPU hands the whole
 * vertex array, the density and C_local2global to the device instead of pulling points one by one. */
class PointcloudAdapter {
public:
    PointcloudAdapter(envire::Pointcloud *cloud, double density) : cloud_(cloud), cursor_(0.0), density_(density)
    {
        envire::FrameNode *fn = cloud_->getFrameNode();
        envire::Environment *env = cloud_->getEnvironment();
        if (!fn || !env) throw std::runtime_error("PointcloudAdapter: pointcloud needs a FrameNode and an Environment");
        local2global_ = env->relativeTransform(fn, env->getRootNode());
        local2globalnew_ = local2global_;
    }
    void setOffsetTransform(const Eigen::Affine3d &C_global2globalnew) { local2globalnew_ = C_global2globalnew * local2global_; }
    void reset() { cursor_ = 0.0; }
    bool hasNext() const { return static_cast<size_t>(cursor_) < cloud_->vertices.size(); }
    Eigen::Vector3d next()
    {
        const size_t i = static_cast<size_t>(cursor_);
        cursor_ += 1.0 / density_;
        return local2globalnew_ * cloud_->vertices[i];
    }
    size_t size() const { return cloud_->vertices.size() * density_; }

    /** FrameNode update after an align (icp/icp.hpp:148-162): compose in the cloud's own frame, then take the
     * rotation() of the result so numeric run-off cannot accumulate. */
    void applyTransform(const Eigen::Affine3d &C_global2globalnew)
    {
        envire::FrameNode *fn = cloud_->getFrameNode();
        Eigen::Affine3d updated = fn->getTransform() * local2global_.inverse(Eigen::Isometry) * C_global2globalnew * local2global_;
        const Eigen::Matrix3d rot = updated.rotation();
        updated.linear() = rot;
        fn->setTransform(envire::Transform(updated));
    }

    // raw access for the device path
    const double *vertexData() const { return cloud_->vertices.empty() ? 0 : cloud_->vertices[0].data(); }
    size_t vertexCount() const { return cloud_->vertices.size(); }
    double density() const { return density_; }
    const Eigen::Affine3d &local2global() const { return local2global_; }

protected:
    envire::Pointcloud *cloud_;
    Eigen::Affine3d local2global_, local2globalnew_;
    double cursor_;
    double density_;
};

/** Adapter with per-vertex normals and scan-edge attributes (role of PointcloudEdgeAndNormalAdapter,
 * icp/icp.hpp:173-198).  Passing it to TrimmedGPU slices it to the plain adapter, exactly as passing the reference's
 * adapter to TrimmedKD does (test/unit/icp.cpp:155); TrimmedGPUEAN applies the pair filter. */
class PointcloudEdgeAndNormalAdapter : public PointcloudAdapter {
public:
    PointcloudEdgeAndNormalAdapter(envire::Pointcloud *cloud, double density) : PointcloudAdapter(cloud, density)
    {
        attrs_ = &cloud->getVertexData<envire::Pointcloud::vertex_attr>(envire::Pointcloud::VERTEX_ATTRIBUTES);
        normals_ = &cloud->getVertexData<Eigen::Vector3d>(envire::Pointcloud::VERTEX_NORMAL);
        if (attrs_->size() != cloud->vertices.size() || normals_->size() != cloud->vertices.size())
            throw std::runtime_error("PointcloudEdgeAndNormalAdapter: normals / attributes do not match the vertices");
    }
    const double *normalData() const { return normals_->empty() ? 0 : (*normals_)[0].data(); }
    const int32_t *attrData() const { return attrs_->empty() ? 0 : reinterpret_cast<const int32_t *>(attrs_->data()); }

private:
    std::vector<Eigen::Vector3d> *normals_;
    std::vector<envire::Pointcloud::vertex_attr> *attrs_;
};

namespace detail {
// how an adapter type reaches the device: plain vertices, or vertices + normals + attributes
inline int modelAdd(icpb_ctx *c, const PointcloudAdapter &a, const double T[16])
{
    return icpb_model_add(c, a.vertexData(), a.vertexCount(), T, a.density());
}
inline int modelAdd(icpb_ctx *c, const PointcloudEdgeAndNormalAdapter &a, const double T[16])
{
    return icpb_model_add_ean(c, a.vertexData(), a.normalData(), a.attrData(), a.vertexCount(), T, a.density());
}
inline int sourceUpload(icpb_ctx *c, const PointcloudAdapter &a, const double T[16])
{
    return icpb_source_upload(c, a.vertexData(), a.vertexCount(), T, a.density());
}
inline int sourceUpload(icpb_ctx *c, const PointcloudEdgeAndNormalAdapter &a, const double T[16])
{
    return icpb_source_upload_ean(c, a.vertexData(), a.normalData(), a.attrData(), a.vertexCount(), T, a.density());
}
} // namespace detail

/** Golden-section minimiser used by the overlap search (role of GoldenBracket, icp/icp.hpp:266-327): same bracket
 * rule, same constant R = 0.6180339, same termination test, so the sequence of probed overlaps is identical. */
template <class T> struct GoldenSection {
    static T minimise(const std::function<T(double)> &f, double a, double b, double c, double eps)
    {
        const double R = 0.6180339, S = 1.0 - R;
        double lo = a, hi = c, u, v;
        if (std::fabs(c - b) > std::fabs(b - a)) { u = b; v = b + S * (c - b); }
        else { v = b; u = b - S * (b - a); }
        T fu = f(u), fv = f(v);
        while (std::fabs(hi - lo) > eps * (std::fabs(u) + std::fabs(v))) {
            if (fv < fu) {
                const double probe = R * u + S * hi; // the reference forms the new point from the OLD inner point
                lo = u; u = v; v = probe;
                fu = fv; fv = f(v);
            } else {
                const double probe = R * v + S * lo;
                hi = v; v = u; u = probe;
                fv = fu; fu = f(u);
            }
        }
        return (fu < fv) ? fu : fv;
    }
};

/** Trimmed<_Adapter, ...> over the device path; _Adapter selects plain (TrimmedKD) or edge-and-normal (TrimmedKDEAN) pairing */
template <class _Adapter> class TrimmedGPUT {
    struct Result {
        Eigen::Affine3d C_global2globalnew;
        size_t iter, pairs;
        double mse, mse_diff, d_box, overlap;
        std::vector<double> pairs_distance;
        Result() : C_global2globalnew(Eigen::Affine3d::Identity()), iter(0), pairs(0), mse(0), mse_diff(0), d_box(0), overlap(0) {}
        double optFunc() const { return mse / std::pow(overlap, 3.0); } // gamma = 2 (icp/icp.hpp:349-362)
        bool operator<(const Result &o) const { return optFunc() < o.optFunc(); }
    };

public:
    explicit TrimmedGPUT(int device = 0) : ctx_(0), multi_(0)
    {
        const int rc = icpb_create(&ctx_, device);
        if (rc != ICPB_OK) {
            std::string msg = std::string("TrimmedGPU: ") + icpb_strerror(rc);
            if (ctx_) { msg += std::string(" / ") + icpb_last_error(ctx_); icpb_destroy(ctx_); ctx_ = 0; }
            throw std::runtime_error(msg);
        }
    }
    /** several GPUs of one NVLink box, this one process: devices first_device .. first_device + n_devices - 1; the
     * measurement is sharded over them, the model replicated (icpb_create_multi).  Plain adapter only. */
    TrimmedGPUT(int first_device, int n_devices) : ctx_(0), multi_(0)
    {
        static_assert(std::is_same<_Adapter, PointcloudAdapter>::value, "several devices: plain PointcloudAdapter only");
        std::vector<int> ids;
        for (int i = 0; i < n_devices; ++i) ids.push_back(first_device + i);
        const int rc = icpb_create_multi(&multi_, ids.data(), n_devices);
        if (rc != ICPB_OK) {
            const std::string msg = std::string("TrimmedGPU: ") + icpb_strerror(rc) + " / " + icpb_multi_last_error(multi_);
            icpb_destroy_multi(multi_);
            multi_ = 0;
            throw std::runtime_error(msg);
        }
        ctx_ = icpb_multi_context(multi_, 0);
    }
    ~TrimmedGPUT()
    {
        if (multi_) icpb_destroy_multi(multi_);
        else if (ctx_) icpb_destroy(ctx_);
    }
    TrimmedGPUT(const TrimmedGPUT &) = delete;
    TrimmedGPUT &operator=(const TrimmedGPUT &) = delete;

    /** golden-section search for the overlap in [alpha, beta] (icp/icp.hpp:385-401) */
    void align(_Adapter measurement, size_t max_iter, double min_mse, double min_mse_diff, double alpha, double beta, double eps)
    {
        upload(measurement); // the vertices do not change between probes: upload once, every probe restarts from identity
        std::function<Result(double)> eval = [&](double overlap) { return alignResident(max_iter, min_mse, min_mse_diff, overlap, true); };
        minResult_ = GoldenSection<Result>::minimise(eval, alpha, (alpha + beta) / 2.0, beta, eps);
        measurement.applyTransform(minResult_.C_global2globalnew);
        traceResult();
    }
    /** fixed overlap (icp/icp.hpp:413-418) */
    void align(_Adapter measurement, size_t max_iter, double min_mse, double min_mse_diff, double overlap)
    {
        upload(measurement);
        minResult_ = alignResident(max_iter, min_mse, min_mse_diff, overlap, false);
        measurement.applyTransform(minResult_.C_global2globalnew);
        traceResult();
    }
    void addToModel(_Adapter model)
    {
        // (the kept distances of the last align stay listable after a model change: icpb_pairs_distance)
        double T[16];
        toColMajor(model.local2global(), T);
        if (multi_) check(icpb_multi_model_add(multi_, model.vertexData(), model.vertexCount(), T, model.density()));
        else check(detail::modelAdd(ctx_, model, T));
    }
    void clearModel()
    {
        check(multi_ ? icpb_multi_model_clear(multi_) : icpb_model_clear(ctx_));
    }

    size_t getNumIterations() { return minResult_.iter; }
    double getMeanSquareError() { return minResult_.mse; }
    double getMeanSquareErrorDiff() { return minResult_.mse_diff; }
    double getOverlap() { return minResult_.overlap; }
    size_t getPairs() { return minResult_.pairs; }
    std::vector<double> getPairsDistance()
    {
        fetchPairsDistance(minResult_);
        return minResult_.pairs_distance;
    }

    /** GPU-only knobs (cell size, warm start, ...): see icpb_set_option */
    void setOption(const char *key, double value) { check(multi_ ? icpb_multi_set_option(multi_, key, value) : icpb_set_option(ctx_, key, value)); }
    icpb_ctx *handle() { return ctx_; } // several devices: the first device's context
    const Eigen::Affine3d &getTransform() const { return minResult_.C_global2globalnew; }

private:
    icpb_ctx *ctx_;
    icpb_multi *multi_; // several devices in this process (ctx_ then is the first device's context, owned by multi_)
    Result minResult_;
    bool pd_pending_ = false;

    /** ICP_B200_TRACE=1: one line per align on stderr (the reference keeps such a trace commented out, icp/icp.hpp:487-495) */
    void traceResult()
    {
        const char *t = getenv("ICP_B200_TRACE");
        if (!t || !*t || *t == '0') return;
        fprintf(stderr, "[icp_b200] align: iter %zu pairs %zu overlap %.17g mse %.17g mse_diff %.17g\n", minResult_.iter,
                minResult_.pairs, minResult_.overlap, minResult_.mse, minResult_.mse_diff);
    }
    void check(int rc)
    {
        if (rc != ICPB_OK)
            throw std::runtime_error(std::string("TrimmedGPU: ") + icpb_strerror(rc) + " / " + (multi_ ? icpb_multi_last_error(multi_) : icpb_last_error(ctx_)));
    }
    static void toColMajor(const Eigen::Affine3d &a, double T[16])
    {
        const Eigen::Matrix4d m = a.matrix();
        for (int c = 0; c < 4; ++c)
            for (int r = 0; r < 4; ++r) T[c * 4 + r] = m(r, c);
    }
    void upload(const _Adapter &meas)
    {
        double T[16];
        toColMajor(meas.local2global(), T);
        if (multi_) check(icpb_multi_source_upload(multi_, meas.vertexData(), meas.vertexCount(), T, meas.density()));
        else check(detail::sourceUpload(ctx_, meas, T));
    }
    Result alignResident(size_t max_iter, double min_mse, double min_mse_diff, double overlap, bool eager_pairs)
    {
        icpb_params prm;
        prm.max_iter = max_iter;
        prm.min_mse = min_mse;
        prm.min_mse_diff = min_mse_diff;
        prm.overlap = overlap;
        icpb_result r;
        check(multi_ ? icpb_multi_align_resident(multi_, &prm, &r) : icpb_align_resident(ctx_, &prm, &r));
        Result out;
        Eigen::Matrix4d m;
        for (int c = 0; c < 4; ++c)
            for (int rr = 0; rr < 4; ++rr) m(rr, c) = r.C_global2globalnew[c * 4 + rr];
        out.C_global2globalnew = Eigen::Affine3d(m);
        out.iter = r.iter;
        out.pairs = r.pairs;
        out.mse = r.mse;
        out.mse_diff = r.mse_diff;
        out.d_box = r.d_box;
        out.overlap = r.overlap;
        pd_pending_ = true;
        // golden section: later probes overwrite the device buffers, so each probe's distances are fetched right away
        if (eager_pairs) fetchPairsDistance(out);
        return out;
    }
    void fetchPairsDistance(Result &res)
    {
        if (!pd_pending_) return;
        size_t n = 0;
        check(multi_ ? icpb_multi_pairs_distance(multi_, 0, 0, &n) : icpb_pairs_distance(ctx_, 0, 0, &n));
        res.pairs_distance.resize(n);
        if (n) check(multi_ ? icpb_multi_pairs_distance(multi_, res.pairs_distance.data(), n, &n) : icpb_pairs_distance(ctx_, res.pairs_distance.data(), n, &n));
        res.pairs_distance.resize(n);
        pd_pending_ = false;
    }
};

typedef TrimmedGPUT<PointcloudAdapter> TrimmedGPU;                   // counterpart of TrimmedKD    (icp/icp.hpp:522)
typedef TrimmedGPUT<PointcloudEdgeAndNormalAdapter> TrimmedGPUEAN;   // counterpart of TrimmedKDEAN (icp/icp.hpp:521)

#ifdef ICP_B200_REPLACE_TRIMMEDKD
/** drop-in: existing callers of envire::icp::TrimmedKD (icp/icpLocalization.hpp, viz/enview/ICPHandler) pick up the GPU path */
typedef TrimmedGPU TrimmedKD;
typedef TrimmedGPUEAN TrimmedKDEAN;
#endif

} // namespace icp
} // namespace envire
#endif

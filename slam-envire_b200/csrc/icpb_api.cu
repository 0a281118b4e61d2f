// icpb_api.cu -- C ABI (include/icp_b200.h) over the sm_100a kernels.
//
// Host orchestration of the B200 trimmed-ICP path: model accumulation + grid
// build, source upload, the _align loop (enqueued iteration by iteration, the
// loop condition is evaluated on the device; the host only peeks at a pinned
// flag so it never stalls the stream), and result read-back.
// There is NO CPU fallback in this file: every compute step is a kernel.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <new>
#include <string>
#include <vector>

#include "../../include/icp_b200.h"
#include "grid_view.h"
#include "pose.h"
#include "primitives.cuh"
#include "kernels.cuh"

using namespace icpb;

// ---------------------------------------------------------------------------
// NCCL, resolved at run time (so the library loads on hosts without it and
// shares libnccl.so.2 with whatever the embedding process already loaded)
// ---------------------------------------------------------------------------
namespace {
typedef struct ncclComm *nccl_comm_t;
typedef struct {
    char internal[128];
} nccl_uid_t;
enum { NCCL_SUM = 0 };
enum { NCCL_UINT32 = 3, NCCL_UINT64 = 5, NCCL_FLOAT64 = 8 };
struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(nccl_uid_t *) = nullptr;
    int (*CommInitRank)(nccl_comm_t *, int, nccl_uid_t, int) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
};
NcclApi g_nccl;
bool nccl_load()
{
    if (g_nccl.ok) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        g_nccl.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) return false;
    g_nccl.GetUniqueId = (int (*)(nccl_uid_t *))dlsym(g_nccl.handle, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(nccl_comm_t *, int, nccl_uid_t, int))dlsym(g_nccl.handle, "ncclCommInitRank");
    g_nccl.CommDestroy = (int (*)(nccl_comm_t))dlsym(g_nccl.handle, "ncclCommDestroy");
    g_nccl.AllReduce = (int (*)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t))dlsym(
        g_nccl.handle, "ncclAllReduce");
    g_nccl.AllGather =
        (int (*)(const void *, void *, size_t, int, nccl_comm_t, cudaStream_t))dlsym(g_nccl.handle, "ncclAllGather");
    g_nccl.GetErrorString = (const char *(*)(int))dlsym(g_nccl.handle, "ncclGetErrorString");
    g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.AllReduce &&
                g_nccl.AllGather && g_nccl.GetErrorString;
    return g_nccl.ok;
}

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;
    // contents are NOT kept; capacity grows geometrically so a growing model does not pay cudaFree/cudaMalloc per add
    cudaError_t ensure(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        const size_t want = cap ? std::max(n, cap + cap / 2) : n;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc((void **)&p, want * sizeof(T));
        if (e != cudaSuccess && want > n) { // tight on memory: fall back to the exact size
            cudaGetLastError();
            e = cudaMalloc((void **)&p, n * sizeof(T));
            if (e == cudaSuccess) cap = n;
            return e;
        }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    // grow, keeping the first `used` elements
    cudaError_t ensure_keep(size_t n, size_t used, cudaStream_t s)
    {
        if (n <= cap) return cudaSuccess;
        size_t nc = cap ? cap : 1024;
        while (nc < n) nc *= 2;
        T *q = nullptr;
        cudaError_t e = cudaMalloc((void **)&q, nc * sizeof(T));
        if (e != cudaSuccess) return e;
        if (p && used) {
            e = cudaMemcpyAsync(q, p, used * sizeof(T), cudaMemcpyDeviceToDevice, s);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        }
        if (p) cudaFree(p);
        p = q;
        cap = nc;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

void cm_to_aff12(const double M[16], double a[12])
{
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) a[r * 4 + c] = M[c * 4 + r];
        a[r * 4 + 3] = M[12 + r];
    }
}
void aff12_to_cm(const double a[12], double M[16])
{
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) M[c * 4 + r] = a[r * 4 + c];
        M[12 + r] = a[r * 4 + 3];
        M[r * 4 + 3] = 0.0;
    }
    M[15] = 1.0;
}

// PointcloudAdapter iteration order (reference icp/icp.hpp:126-138):
// idx = size_t(index); index += 1.0/density; stop when idx >= n
const double *density_gather(const double *xyz, size_t n, double density, std::vector<double> &tmp, size_t *n_out)
{
    if (density == 1.0) {
        *n_out = n;
        return xyz;
    }
    tmp.clear();
    double index = 0.0;
    for (;;) {
        const size_t idx = (size_t)index;
        if (!(idx < n)) break;
        tmp.push_back(xyz[3 * idx]);
        tmp.push_back(xyz[3 * idx + 1]);
        tmp.push_back(xyz[3 * idx + 2]);
        index += 1.0 / density;
    }
    *n_out = tmp.size() / 3;
    return tmp.data();
}
// the same visiting order applied to a parallel per-vertex array of `width` items
template <typename T>
const T *density_gather_t(const T *v, size_t n, int width, double density, std::vector<T> &tmp)
{
    if (density == 1.0) return v;
    tmp.clear();
    double index = 0.0;
    for (;;) {
        const size_t idx = (size_t)index;
        if (!(idx < n)) break;
        for (int k = 0; k < width; ++k) tmp.push_back(v[width * idx + k]);
        index += 1.0 / density;
    }
    return tmp.data();
}
// NVTX range over a host-side phase (model add / grid build / merge-append / source upload / align): shows up in
// Nsight timelines; costs a few nanoseconds without a tool attached
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};
// ICPB_VERBOSE=1: wall-clock marks (with a stream sync each) through the build / append paths, printed to stderr
struct PhaseMarks {
    bool on;
    cudaStream_t s;
    const char *what;
    std::chrono::steady_clock::time_point t0, last;
    std::string line;
    PhaseMarks(const char *w, cudaStream_t st) : on(getenv("ICPB_VERBOSE") != nullptr), s(st), what(w)
    {
        if (on) {
            cudaStreamSynchronize(s);
            t0 = last = std::chrono::steady_clock::now();
        }
    }
    void mark(const char *label)
    {
        if (!on) return;
        cudaStreamSynchronize(s);
        const auto now = std::chrono::steady_clock::now();
        char buf[96];
        snprintf(buf, sizeof(buf), " %s %.0f", label, std::chrono::duration<double, std::micro>(now - last).count());
        line += buf;
        last = now;
    }
    ~PhaseMarks()
    {
        if (on) fprintf(stderr, "[icpb] %s (us):%s | total %.0f\n", what, line.c_str(), std::chrono::duration<double, std::micro>(last - t0).count());
    }
};
constexpr int MAX_DIM = 8192; // cells per axis: 13 levels above level 0; walks start at level 12 at most (ICPB_WALK_TOP), where <= 2 nodes per axis are left
constexpr int EVENT_ITERS = 512;
constexpr int FLAG_RING = 4096;
} // namespace

struct icpb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    // options
    double opt_cell_size = 0.0, opt_target = 12.0, opt_max_cells = 64.0 * 1024 * 1024, opt_run_ahead = 2.0,
           opt_warm = 1.0, opt_skip = 1.0, opt_skip_reach = 0.25, opt_skip_floor = 0.05, opt_lazy = 1.0, opt_lazy_margin = 0.05, opt_fail_build = 0.0, opt_linear_trim = 1.0, opt_fuse_moments = -1.0, opt_morton_bits = 10.0, opt_hilbert = 1.0, opt_block_levels = 1.0, opt_profile = 0.0, opt_sort = 1.0, opt_margin = 0.0, opt_merge = 1.0, opt_graph = 1.0, opt_trim_bpsm = 2.0, opt_mom_bpsm = 2.0;
    // model
    DevBuf<double> model_raw, model_nrm;
    DevBuf<uint8_t> model_edge;
    DevBuf<NPoint> npts;
    bool model_ean = false;
    DevBuf<int> stage_i;
    size_t model_n = 0;
    bool grid_dirty = true;
    DevBuf<MPoint> pts, pts2;
    DevBuf<FPoint> ptsf, ptsf2;
    DevBuf<uint32_t> pkey, pkey2, newpref;
    DevBuf<NPoint> npts2;
    GridDims gdims;
    unsigned long long merges = 0;
    DevBuf<uint32_t> cell_start, cell_box;
    DevBuf<uint32_t> masks;
    DevBuf<uint32_t> keys0, keys1, vals0, vals1, sort_scratch;
    DevBuf<double> bbox_part;
    GridView gv;
    icpb_grid_info ginfo;
    // source
    DevBuf<double> stage;
    DevBuf<double> sx, sy, sz, snx, sny, snz;
    DevBuf<uint8_t> sedge, sedge_tmp;
    bool src_ean = false;
    size_t src_n = 0, adapter_size = 0;
    double Cl2g[12];
    bool have_source = false;
    DevBuf<uint32_t> nn_pos, src_perm;
    DevBuf<double> nn_d;
    DevBuf<float> nn_lb; // per query: lower bound on the distance to every model point but its winner (search skipping)
    // iteration
    IcpState *d_state = nullptr;
    IcpState *h_state = nullptr; // pinned
    int *h_flags = nullptr;      // pinned
    uint32_t *d_ghist = nullptr;
    bool upload_pending = false; // source copies enqueued, not waited for yet
    DensitySeg h_dseg[ICPB_DENSITY_MAX_SEGS];
    DevBuf<DensitySeg> d_dseg;
    DevBuf<double> stage_full; // the whole host array of a density-stepped upload (the visited vertices are gathered from it)
    struct { // between align_enqueue and align_collect
        unsigned long long launches0 = 0;
        bool use_graph = false, linear_ok = false, pending = false;
        uint32_t n = 0;
        size_t timed_iters = 0;
    } run;
    unsigned tm_blocks = 0; // co-resident blocks of trim_moments_kernel (0: not queried yet)
    uint32_t *d_lhist = nullptr; // linear histogram of the pair distances (search epilogue -> trim_moments_kernel)
    unsigned long long *d_counter = nullptr;
    DevBuf<double> partials;
    DevBuf<double> trace;
    int trace_cap = 0;
    int last_executed = 0;
    bool last_valid = false, last_early = false;
    bool pd_valid = false; // nn_d / src_perm / the trim state of the last align are intact: its kept distances can still be listed
                           // (a model change invalidates the winners' positions -- last_valid -- but not the distances)
    // K7 / debug scratch
    DevBuf<unsigned long long> kd0, kd1, cand_key, seg_key, cell_bits; // cell_bits: occupied sixteenths per level-0 cell and axis (tight boxes) // seg_*: per-warp candidate segments of trim_moments_kernel
    DevBuf<uint32_t> cand_idx, xscratch, seg_idx, seg_count, occ_bits; // occ_bits: one bit per cell while the build tunes the cell edge
    DevBuf<uint32_t> dbg_idx;
    DevBuf<uint8_t> dbg_kept;
    // timing
    std::vector<cudaEvent_t> events;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr, ev_up0 = nullptr, ev_up1 = nullptr, ev_b0 = nullptr, ev_b1 = nullptr;
    std::vector<cudaEvent_t> ring;
    icpb_timing timing;
    std::vector<double> iter_ms;
    unsigned long long launches = 0;
    // nccl
    nccl_comm_t comm = nullptr;
    int n_ranks = 1, rank = 0;
    unsigned long long *d_gather = nullptr;
    // CUDA-graph form of the align loop (WHILE conditional node), rebuilt when any captured argument changes
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    cudaGraphNode_t graph_init_node = nullptr;
    cudaGraphConditionalHandle graph_cond = 0;
    std::vector<unsigned char> graph_key;
    // peer-memory collectives
    uint32_t *p2p_local = nullptr;
    void *p2p_mapped[P2P_MAX_RANKS] = {nullptr};
    P2PView p2p;
    bool use_p2p = false;

    int fail(cudaError_t e, const char *what, int line)
    {
        char buf[512];
        snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s:%d [%s]", (int)e, cudaGetErrorString(e), __FILE__, line,
                 what);
        err = buf;
        cudaGetLastError();
        return ICPB_ERR_CUDA;
    }
    int fail_nccl(int rc, const char *what)
    {
        char buf[512];
        snprintf(buf, sizeof(buf), "NCCL error %d (%s) in %s", rc,
                 g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?", what);
        err = buf;
        return ICPB_ERR_NCCL;
    }
};

#define CK(call)                                                    \
    do {                                                            \
        cudaError_t e_ = (call);                                    \
        if (e_ != cudaSuccess) return ctx->fail(e_, #call, __LINE__); \
    } while (0)
#define CKL() CK(cudaGetLastError())
#define NK(call)                                          \
    do {                                                  \
        int r_ = (call);                                  \
        if (r_ != 0) return ctx->fail_nccl(r_, #call);    \
    } while (0)

static unsigned grid_for(size_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

extern "C" {

const char *icpb_version(void) { return "icp_b200 0.1 (sm_100a)"; }

const char *icpb_strerror(int s)
{
    switch (s) {
    case ICPB_OK: return "ok";
    case ICPB_ERR_CUDA: return "CUDA error (no device / runtime failure)";
    case ICPB_ERR_ARG: return "invalid argument";
    case ICPB_ERR_NO_MODEL: return "no model: call addToModel first";
    case ICPB_ERR_NO_SOURCE: return "no source uploaded";
    case ICPB_ERR_NCCL: return "NCCL error";
    case ICPB_ERR_CAPACITY: return "output capacity too small";
    case ICPB_ERR_NOT_ENOUGH_PAIRS: return "not enough pairs to get transform";
    default: return "unknown status";
    }
}
const char *icpb_last_error(const icpb_ctx *ctx) { return ctx ? ctx->err.c_str() : ""; }

int icpb_create(icpb_ctx **out, int device_id)
{
    if (!out) return ICPB_ERR_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
        cudaGetLastError();
        return ICPB_ERR_CUDA; // no CPU fallback
    }
    if (device_id < 0 || device_id >= count) return ICPB_ERR_ARG;
    icpb_ctx *ctx = new icpb_ctx();
    ctx->device = device_id;
    memset(&ctx->gv, 0, sizeof(ctx->gv));
    memset(&ctx->ginfo, 0, sizeof(ctx->ginfo));
    memset(&ctx->timing, 0, sizeof(ctx->timing));
    memset(&ctx->p2p, 0, sizeof(ctx->p2p));
    ctx->p2p.world = 1;
    *out = ctx;
    CK(cudaSetDevice(device_id));
    CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CK(cudaMalloc((void **)&ctx->d_state, sizeof(IcpState)));
    CK(cudaMemset(ctx->d_state, 0, sizeof(IcpState)));
    CK(cudaMallocHost((void **)&ctx->h_state, sizeof(IcpState)));
    CK(cudaMallocHost((void **)&ctx->h_flags, FLAG_RING * sizeof(int)));
    CK(cudaMalloc((void **)&ctx->d_ghist, (SEL_MAX_BINS + 64) * sizeof(uint32_t)));
    CK(cudaMemset(ctx->d_ghist, 0, (SEL_MAX_BINS + 64) * sizeof(uint32_t)));
    CK(cudaMalloc((void **)&ctx->d_lhist, (LIN_BINS * LIN_STRIDE + 64 + LIN_BINS) * sizeof(uint32_t)));
    CK(cudaMemset(ctx->d_lhist, 0, (LIN_BINS * LIN_STRIDE + 64 + LIN_BINS) * sizeof(uint32_t)));
    CK(cudaMalloc((void **)&ctx->d_counter, 8 * sizeof(unsigned long long)));
    CK(cudaMemset(ctx->d_counter, 0, 8 * sizeof(unsigned long long)));
    CK(cudaMalloc((void **)&ctx->d_gather, 64 * sizeof(unsigned long long)));
    CK(ctx->xscratch.ensure((size_t)(1 + P2P_MAX_RANKS) * P2P_ROW_WORDS)); // candidate exchange of the multi-GPU trim
    CK(cudaMemset(ctx->xscratch.p, 0, (size_t)(1 + P2P_MAX_RANKS) * P2P_ROW_WORDS * sizeof(uint32_t)));
    CK(cudaEventCreate(&ctx->ev_begin));
    CK(cudaEventCreate(&ctx->ev_end));
    CK(cudaEventCreate(&ctx->ev_up0));
    CK(cudaEventCreate(&ctx->ev_b0));
    CK(cudaEventCreate(&ctx->ev_b1));
    CK(cudaEventCreate(&ctx->ev_up1));
    ctx->ring.resize(64);
    for (auto &e : ctx->ring) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    return ICPB_OK;
}

void icpb_destroy(icpb_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->graph_exec) cudaGraphExecDestroy(ctx->graph_exec);
    if (ctx->graph) cudaGraphDestroy(ctx->graph);
    if (ctx->comm && g_nccl.ok) g_nccl.CommDestroy(ctx->comm);
    for (int r = 0; r < P2P_MAX_RANKS; ++r)
        if (ctx->p2p_mapped[r]) cudaIpcCloseMemHandle(ctx->p2p_mapped[r]);
    if (ctx->p2p_local) cudaFree(ctx->p2p_local);
    ctx->model_raw.release();
    ctx->model_nrm.release();
    ctx->model_edge.release();
    ctx->npts.release();
    ctx->stage_i.release();
    ctx->snx.release();
    ctx->sny.release();
    ctx->snz.release();
    ctx->sedge.release();
    ctx->sedge_tmp.release();
    ctx->pts.release();
    ctx->ptsf.release();
    ctx->pts2.release();
    ctx->ptsf2.release();
    ctx->pkey.release();
    ctx->pkey2.release();
    ctx->newpref.release();
    ctx->npts2.release();
    ctx->cell_start.release();
    ctx->cell_box.release();
    ctx->cell_bits.release();
    ctx->occ_bits.release();
    ctx->masks.release();
    ctx->keys0.release();
    ctx->keys1.release();
    ctx->vals0.release();
    ctx->vals1.release();
    ctx->sort_scratch.release();
    ctx->bbox_part.release();
    ctx->stage.release();
    ctx->sx.release();
    ctx->sy.release();
    ctx->sz.release();
    ctx->nn_pos.release();
    ctx->src_perm.release();
    ctx->nn_d.release();
    ctx->nn_lb.release();
    ctx->partials.release();
    ctx->trace.release();
    ctx->kd0.release();
    ctx->cand_key.release();
    ctx->cand_idx.release();
    ctx->seg_key.release();
    ctx->seg_idx.release();
    ctx->seg_count.release();
    ctx->xscratch.release();
    ctx->kd1.release();
    ctx->dbg_idx.release();
    ctx->dbg_kept.release();
    if (ctx->d_state) cudaFree(ctx->d_state);
    if (ctx->h_state) cudaFreeHost(ctx->h_state);
    if (ctx->h_flags) cudaFreeHost(ctx->h_flags);
    if (ctx->d_ghist) cudaFree(ctx->d_ghist);
    if (ctx->d_lhist) cudaFree(ctx->d_lhist);
    if (ctx->d_counter) cudaFree(ctx->d_counter);
    if (ctx->d_gather) cudaFree(ctx->d_gather);
    for (auto e : ctx->events) cudaEventDestroy(e);
    for (auto e : ctx->ring) cudaEventDestroy(e);
    if (ctx->ev_begin) cudaEventDestroy(ctx->ev_begin);
    if (ctx->ev_end) cudaEventDestroy(ctx->ev_end);
    if (ctx->ev_up0) cudaEventDestroy(ctx->ev_up0);
    if (ctx->ev_b0) cudaEventDestroy(ctx->ev_b0);
    if (ctx->ev_b1) cudaEventDestroy(ctx->ev_b1);
    if (ctx->ev_up1) cudaEventDestroy(ctx->ev_up1);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

static double *opt_slot(icpb_ctx *ctx, const char *key)
{
    if (!strcmp(key, "cell_size")) return &ctx->opt_cell_size;
    if (!strcmp(key, "target_pts_per_cell")) return &ctx->opt_target;
    if (!strcmp(key, "max_cells")) return &ctx->opt_max_cells;
    if (!strcmp(key, "run_ahead") || !strcmp(key, "poll_every")) return &ctx->opt_run_ahead;
    if (!strcmp(key, "warm_start")) return &ctx->opt_warm;
    if (!strcmp(key, "skip")) return &ctx->opt_skip;
    if (!strcmp(key, "skip_reach_cells")) return &ctx->opt_skip_reach;
    if (!strcmp(key, "skip_floor_cells")) return &ctx->opt_skip_floor;
    if (!strcmp(key, "block_levels")) return &ctx->opt_block_levels;
    if (!strcmp(key, "lazy")) return &ctx->opt_lazy;
    if (!strcmp(key, "lazy_margin")) return &ctx->opt_lazy_margin;
    if (!strcmp(key, "linear_trim")) return &ctx->opt_linear_trim;
    if (!strcmp(key, "fuse_moments")) return &ctx->opt_fuse_moments;
    if (!strcmp(key, "morton_bits")) return &ctx->opt_morton_bits;
    if (!strcmp(key, "hilbert")) return &ctx->opt_hilbert;
    if (!strcmp(key, "test_fail_next_build")) return &ctx->opt_fail_build;
    if (!strcmp(key, "profile")) return &ctx->opt_profile;
    if (!strcmp(key, "sort_source")) return &ctx->opt_sort;
    if (!strcmp(key, "growth_margin")) return &ctx->opt_margin;
    if (!strcmp(key, "merge_append")) return &ctx->opt_merge;
    if (!strcmp(key, "graph")) return &ctx->opt_graph;
    if (!strcmp(key, "trim_blocks_per_sm")) return &ctx->opt_trim_bpsm;
    if (!strcmp(key, "moments_blocks_per_sm")) return &ctx->opt_mom_bpsm;
    return nullptr;
}
int icpb_set_option(icpb_ctx *ctx, const char *key, double value)
{
    if (!ctx || !key) return ICPB_ERR_ARG;
    double *s = opt_slot(ctx, key);
    if (!s) return ICPB_ERR_ARG;
    if (s == &ctx->opt_cell_size || s == &ctx->opt_target || s == &ctx->opt_max_cells) ctx->grid_dirty = true;
    *s = value;
    return ICPB_OK;
}
int icpb_get_option(const icpb_ctx *ctx, const char *key, double *value)
{
    if (!ctx || !key || !value) return ICPB_ERR_ARG;
    double *s = opt_slot(const_cast<icpb_ctx *>(ctx), key);
    if (!s) return ICPB_ERR_ARG;
    *value = *s;
    return ICPB_OK;
}

// ---------------------------------------------------------------------------
// model
// ---------------------------------------------------------------------------
static int update_masks(icpb_ctx *ctx, const uint32_t *keys, size_t k);
static int stage_visited(icpb_ctx *ctx, const double *xyz, size_t n, double density, size_t *m_out);

// Growing model: merge a freshly appended batch (model_raw[n_old, n_old + k)) into the existing index in one
// streaming pass instead of re-sorting everything.  Falls back (returns with *done == false) when there is no
// built grid or the batch does not fit inside the grid box.
static int try_merge_append(icpb_ctx *ctx, size_t n_old, size_t k, bool *done)
{
    NvtxRange nvtx_range("icpb: merge-append (K1b)");
    *done = false;
    if (ctx->opt_merge == 0.0 || ctx->grid_dirty || n_old == 0 || ctx->gv.n != n_old || k == 0) return ICPB_OK;
    if (n_old + k >= 0xFFFFFFF0ull) return ICPB_OK;
    cudaStream_t s = ctx->stream;
    PhaseMarks pm("merge-append", s);
    const GridView &g = ctx->gv;
    const unsigned nbb = std::min<unsigned>(1024u, grid_for(k, 256));
    CK(ctx->bbox_part.ensure((size_t)nbb * 6));
    bbox_kernel<<<nbb, 256, 0, s>>>(ctx->model_raw.p + 3 * n_old, k, ctx->bbox_part.p);
    CKL();
    std::vector<double> hb((size_t)nbb * 6);
    CK(cudaMemcpyAsync(hb.data(), ctx->bbox_part.p, hb.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    for (int a = 0; a < 3; ++a) {
        double lo = INFINITY, hi = -INFINITY;
        for (unsigned b = 0; b < nbb; ++b) {
            lo = fmin(lo, hb[(size_t)b * 6 + a]);
            hi = fmax(hi, hb[(size_t)b * 6 + 3 + a]);
        }
        // same arithmetic as icpb_cell_coord, but without clamping: the cell must exist
        if (!(floor((lo - g.o[a]) * g.inv_h) >= 0.0) || !(floor((hi - g.o[a]) * g.inv_h) <= (double)(g.dim[0][a] - 1)))
            return ICPB_OK;
    }
    const size_t ncell = (size_t)g.dim[0][0] * g.dim[0][1] * g.dim[0][2];
    pm.mark("bbox");
    cudaEvent_t e0 = ctx->ev_b0, e1 = ctx->ev_b1;
    CK(cudaEventRecord(e0, s));
    CK(ctx->keys0.ensure(k));
    CK(ctx->keys1.ensure(k));
    CK(ctx->vals0.ensure(k));
    CK(ctx->vals1.ensure(k));
    CK(ctx->sort_scratch.ensure(rs_scratch_elems(k)));
    CK(ctx->newpref.ensure(ncell + 1 + scan_scratch_elems(ncell + 1)));
    CK(ctx->pts2.ensure(n_old + k + ICPB_PAD));
    CK(ctx->ptsf2.ensure(n_old + k + ICPB_PAD));
    CK(ctx->pkey2.ensure(n_old + k));
    if (ctx->model_ean) CK(ctx->npts2.ensure(n_old + k));
    pm.mark("alloc");
    cell_keys_kernel<<<grid_for(k, 256), 256, 0, s>>>(ctx->model_raw.p + 3 * n_old, (uint32_t)k, ctx->gdims, ctx->keys0.p,
                                                      ctx->vals0.p, (uint32_t)n_old);
    CKL();
    int bits = 1;
    while (((size_t)1 << bits) < ncell) ++bits;
    const int w = radix_sort<uint32_t, true>(ctx->keys0.p, ctx->vals0.p, ctx->keys1.p, ctx->vals1.p, k, bits,
                                             ctx->sort_scratch.p, s, &ctx->launches);
    CKL();
    const uint32_t *ks = w ? ctx->keys1.p : ctx->keys0.p, *vs = w ? ctx->vals1.p : ctx->vals0.p;
    pm.mark("sort");
    CK(cudaMemsetAsync(ctx->newpref.p, 0, (ncell + 1) * sizeof(uint32_t), s));
    key_hist_kernel<<<grid_for(k, 256), 256, 0, s>>>(ks, (uint32_t)k, ctx->newpref.p);
    CKL();
    exclusive_scan_u32(ctx->newpref.p, ctx->newpref.p, ncell + 1, ctx->newpref.p + ncell + 1, s, &ctx->launches);
    CKL();
    pm.mark("prefix");
    merge_move_old_kernel<<<grid_for(n_old, 256), 256, 0, s>>>(ctx->pts.p, ctx->ptsf.p, ctx->pkey.p,
                                                               ctx->model_ean ? ctx->npts.p : nullptr, (uint32_t)n_old,
                                                               ctx->newpref.p, ctx->pts2.p, ctx->ptsf2.p, ctx->pkey2.p,
                                                               ctx->model_ean ? ctx->npts2.p : nullptr);
    merge_place_new_kernel<<<grid_for(k, 256), 256, 0, s>>>(ctx->model_raw.p, ctx->model_ean ? ctx->model_nrm.p : nullptr,
                                                            ctx->model_ean ? ctx->model_edge.p : nullptr, ks, vs,
                                                            (uint32_t)k, ctx->gdims, ctx->cell_start.p, ctx->pts2.p,
                                                            ctx->ptsf2.p, ctx->pkey2.p,
                                                            ctx->model_ean ? ctx->npts2.p : nullptr, ctx->cell_bits.p);
    add_prefix_kernel<<<grid_for(ncell + 1, 256), 256, 0, s>>>(ctx->cell_start.p, ctx->newpref.p, ncell + 1);
    CKL();
    ctx->launches += 5;
    pad_tail_kernel<<<1, 32, 0, s>>>(ctx->pts2.p, ctx->ptsf2.p, (uint32_t)(n_old + k));
    std::swap(ctx->pts, ctx->pts2);
    std::swap(ctx->ptsf, ctx->ptsf2);
    std::swap(ctx->pkey, ctx->pkey2);
    if (ctx->model_ean) std::swap(ctx->npts, ctx->npts2);
    ctx->gv.pts = ctx->pts.p;
    ctx->gv.ptsf = ctx->ptsf.p;
    ctx->gv.n = (uint32_t)(n_old + k);
    pm.mark("move");
    {
        int rc = update_masks(ctx, ks, k);
        if (rc != ICPB_OK) return rc;
    }
    pm.mark("masks");
    CK(cudaEventRecord(e1, s));
    CK(cudaStreamSynchronize(s));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    ctx->merges += 1;
    ctx->ginfo.n_points = n_old + k;
    ctx->ginfo.merges = ctx->merges;
    ctx->ginfo.build_ms = ms;
    ctx->last_valid = false;
    *done = true;
    return ICPB_OK;
}

static int model_add(icpb_ctx *ctx, const double *xyz, const double *normals, const int32_t *attrs, size_t n,
                     const double T[16], double density)
{
    NvtxRange nvtx_range("icpb: addToModel (K0)");
    if (!ctx || (!xyz && n) || !T || !(density > 0.0)) return ICPB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const bool ean = normals != nullptr;
    if (ean && !attrs) return ICPB_ERR_ARG;
    if (ctx->model_n > 0 && ean != ctx->model_ean) return ICPB_ERR_ARG; // all plain or all edge-and-normal
    PhaseMarks pm("addToModel", ctx->stream);
    size_t m = 0;
    {
        int rc = stage_visited(ctx, xyz, n, density, &m); // K0: the visited vertices, gathered on the device
        if (rc != ICPB_OK) return rc;
    }
    if (m == 0) return ICPB_OK;
    if (ctx->model_n + m >= 0xFFFFFFF0ull) return ICPB_ERR_ARG;
    pm.mark("copy");
    CK(ctx->model_raw.ensure_keep(3 * (ctx->model_n + m), 3 * ctx->model_n, ctx->stream));
    pm.mark("grow");
    Aff12 A;
    cm_to_aff12(T, A.a);
    transform_append_kernel<<<grid_for(m, 256), 256, 0, ctx->stream>>>(ctx->stage.p, m, A,
                                                                       ctx->model_raw.p + 3 * ctx->model_n);
    CKL();
    CK(cudaStreamSynchronize(ctx->stream)); // the caller's buffer may go away
    pm.mark("transform");
    if (ean) {
        std::vector<double> tn;
        std::vector<int32_t> ta;
        const double *nsrc = density_gather_t(normals, n, 3, density, tn);
        const int32_t *asrc = density_gather_t(attrs, n, 1, density, ta);
        CK(ctx->stage_i.ensure(m));
        CK(ctx->model_nrm.ensure_keep(3 * (ctx->model_n + m), 3 * ctx->model_n, ctx->stream));
        CK(ctx->model_edge.ensure_keep(ctx->model_n + m, ctx->model_n, ctx->stream));
        CK(cudaMemcpyAsync(ctx->stage.p, nsrc, 3 * m * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->stage_i.p, asrc, m * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        rotate_append_kernel<<<grid_for(m, 256), 256, 0, ctx->stream>>>(ctx->stage.p, m, A, ctx->model_nrm.p + 3 * ctx->model_n);
        edge_flags_kernel<<<grid_for(m, 256), 256, 0, ctx->stream>>>(ctx->stage_i.p, m, ctx->model_edge.p + ctx->model_n);
        CKL();
        CK(cudaStreamSynchronize(ctx->stream));
    }
    ctx->model_ean = ean;
    const size_t n_old = ctx->model_n;
    ctx->model_n += m;
    bool merged = false;
    {
        int rc = try_merge_append(ctx, n_old, m, &merged);
        if (rc != ICPB_OK) {
            ctx->grid_dirty = true; // the index may be half merged: the points are in model_raw, rebuild before the next use
            return rc;
        }
    }
    if (!merged) ctx->grid_dirty = true;
    return ICPB_OK;
}

int icpb_model_add(icpb_ctx *ctx, const double *xyz, size_t n, const double T[16], double density)
{
    return model_add(ctx, xyz, nullptr, nullptr, n, T, density);
}
int icpb_model_add_ean(icpb_ctx *ctx, const double *xyz, const double *normals, const int32_t *attrs, size_t n,
                       const double T[16], double density)
{
    if (!normals || !attrs) return ICPB_ERR_ARG;
    return model_add(ctx, xyz, normals, attrs, n, T, density);
}

int icpb_model_clear(icpb_ctx *ctx)
{
    if (!ctx) return ICPB_ERR_ARG;
    ctx->model_n = 0;
    ctx->model_ean = false;
    ctx->grid_dirty = true;
    ctx->gv.n = 0;
    return ICPB_OK;
}
int icpb_model_size(const icpb_ctx *ctx, size_t *n)
{
    if (!ctx || !n) return ICPB_ERR_ARG;
    *n = ctx->model_n;
    return ICPB_OK;
}

static void dims_for(const double ext[3], double h, int d[3])
{
    for (int a = 0; a < 3; ++a) {
        double v = floor(ext[a] / h) + 1.0;
        if (v < 1.0) v = 1.0;
        if (v > 1e9) v = 1e9;
        d[a] = (int)v;
    }
}

// Pyramid geometry of the grid in ctx->gv (dims per level, node arrays).  The node arrays are (re)allocated only when
// the geometry changed.
static int layout_masks(icpb_ctx *ctx, size_t mask_off[ICPB_MAX_LEVELS])
{
    GridView &g = ctx->gv;
    int L = 0;
    size_t mask_total = 0;
    while (g.dim[L][0] > 1 || g.dim[L][1] > 1 || g.dim[L][2] > 1) {
        for (int a = 0; a < 3; ++a) g.dim[L + 1][a] = (g.dim[L][a] + 1) >> 1;
        ++L;
        mask_off[L] = mask_total;
        mask_total += (size_t)g.dim[L][0] * g.dim[L][1] * g.dim[L][2];
    }
    g.L = L;
    CK(ctx->masks.ensure(mask_total + 16));
    CK(ctx->cell_box.ensure((size_t)g.dim[0][0] * g.dim[0][1] * g.dim[0][2] + 16));
    g.cell_box = ctx->cell_box.p;
    for (int l = 1; l <= L; ++l) g.node[l] = ctx->masks.p + mask_off[l];
    return ICPB_OK;
}

// Tight boxes of all level-0 cells from cell_bits[] (filled by gather_points_kernel) and the pyramid of node words
// (child occupancy + tight boxes) above them: one pass per level over the cells of that level.
static int build_masks(icpb_ctx *ctx)
{
    cudaStream_t s = ctx->stream;
    GridView &g = ctx->gv;
    size_t mask_off[ICPB_MAX_LEVELS] = {0};
    {
        int rc = layout_masks(ctx, mask_off);
        if (rc != ICPB_OK) return rc;
    }
    const size_t ncell = (size_t)g.dim[0][0] * g.dim[0][1] * g.dim[0][2];
    cell_words_kernel<<<grid_for(ncell, 256), 256, 0, s>>>(ctx->cell_bits.p, ncell, ctx->cell_box.p);
    CKL();
    for (int l = 1; l <= g.L; ++l) {
        const size_t cells = (size_t)g.dim[l][0] * g.dim[l][1] * g.dim[l][2];
        const uint32_t *below = l == 1 ? ctx->cell_box.p : g.node[l - 1];
        node_levelN_kernel<<<grid_for(cells, 256), 256, 0, s>>>(below, g.dim[l - 1][0], g.dim[l - 1][1], g.dim[l - 1][2],
                                                                g.dim[l][0], g.dim[l][1], g.dim[l][2],
                                                                const_cast<uint32_t *>(g.node[l]));
        CKL();
    }
    ctx->launches += 1 + (unsigned long long)g.L;
    return ICPB_OK;
}

// Growing model: only the cells a batch of k new points fell into (keys[], level-0 cell keys) and their ancestors
// change -- boxes only grow (cell_bits[] has the new points ORed in already).  O(k * levels) instead of a pass over
// every cell of every level.
static int update_masks(icpb_ctx *ctx, const uint32_t *keys, size_t k)
{
    cudaStream_t s = ctx->stream;
    GridView &g = ctx->gv;
    cell_words_update_kernel<<<grid_for(k, 256), 256, 0, s>>>(ctx->cell_bits.p, keys, (uint32_t)k, ctx->cell_box.p);
    CKL();
    for (int l = 1; l <= g.L; ++l) {
        const uint32_t *below = l == 1 ? ctx->cell_box.p : g.node[l - 1];
        node_update_kernel<<<grid_for(k, 256), 256, 0, s>>>(keys, (uint32_t)k, g.dim[0][0], g.dim[0][1], l, below,
                                                            g.dim[l - 1][0], g.dim[l - 1][1], g.dim[l - 1][2], g.dim[l][0],
                                                            g.dim[l][1], const_cast<uint32_t *>(g.node[l]));
        CKL();
    }
    ctx->launches += 1 + (unsigned long long)g.L;
    return ICPB_OK;
}

static int build_grid(icpb_ctx *ctx)
{
    NvtxRange nvtx_range("icpb: grid build (K1)");
    const size_t n = ctx->model_n;
    cudaStream_t s = ctx->stream;
    memset(&ctx->ginfo, 0, sizeof(ctx->ginfo));
    ctx->gv.n = 0;
    ctx->grid_dirty = true; // cleared at the very end: a build that fails half-way is repeated (and fails again) next time
    if (n == 0) {
        ctx->grid_dirty = false;
        return ICPB_OK;
    }
    PhaseMarks pm("build", s);
    cudaEvent_t e0 = ctx->ev_b0, e1 = ctx->ev_b1;
    CK(cudaEventRecord(e0, s));
    // bounding box
    const unsigned nbb = std::min<unsigned>(1024u, grid_for(n, 256));
    CK(ctx->bbox_part.ensure((size_t)nbb * 6));
    bbox_kernel<<<nbb, 256, 0, s>>>(ctx->model_raw.p, n, ctx->bbox_part.p);
    CKL();
    std::vector<double> hb((size_t)nbb * 6);
    CK(cudaMemcpyAsync(hb.data(), ctx->bbox_part.p, hb.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (unsigned b = 0; b < nbb; ++b)
        for (int a = 0; a < 3; ++a) {
            mn[a] = fmin(mn[a], hb[(size_t)b * 6 + a]);
            mx[a] = fmax(mx[a], hb[(size_t)b * 6 + 3 + a]);
        }
    double ext[3];
    for (int a = 0; a < 3; ++a) {
        ext[a] = mx[a] - mn[a];
        if (!(ext[a] >= 0.0) || !isfinite(ext[a])) return ICPB_ERR_ARG; // NaN / inf vertices
    }
    if (ctx->opt_fail_build != 0.0) { // fault injection for the tests: this build fails after its first steps
        ctx->opt_fail_build = 0.0;
        ctx->err = "injected build failure (test_fail_next_build)";
        return ICPB_ERR_CUDA;
    }
    double se[3] = {ext[0], ext[1], ext[2]};
    std::sort(se, se + 3);
    // growth margin: the grid box is the data box widened on every side, so that clouds added later (a growing
    // map) fall inside it and can be merged in without a rebuild
    const double margin = fmax(0.0, fmin(ctx->opt_margin, 8.0));
    const double dext[3] = {ext[0], ext[1], ext[2]};
    for (int a = 0; a < 3; ++a) {
        mn[a] -= margin * dext[a];
        ext[a] = dext[a] * (1.0 + 2.0 * margin);
    }
    const double emax = se[2] * (1.0 + 2.0 * margin);
    // smallest admissible cell edge: <= MAX_DIM cells per axis, <= max_cells cells in total
    auto h_floor = [&](double h) {
        if (emax > 0.0) h = fmax(h, emax / (double)(MAX_DIM - 1));
        if (!(h > 0.0)) h = 1.0;
        for (int it = 0; it < 200; ++it) {
            int d[3];
            dims_for(ext, h, d);
            if ((double)d[0] * d[1] * d[2] <= ctx->opt_max_cells && d[0] <= MAX_DIM && d[1] <= MAX_DIM && d[2] <= MAX_DIM)
                break;
            h *= 1.1;
        }
        return h;
    };
    const bool auto_h = !(ctx->opt_cell_size > 0.0);
    double h;
    if (!auto_h)
        h = h_floor(ctx->opt_cell_size);
    else {
        // first guess: points spread over the largest face of the bounding box
        double area = se[2] * se[1];
        if (!(area > 0.0)) area = se[2] * se[2];
        h = area > 0.0 ? sqrt(ctx->opt_target * area / (double)n) : 1.0;
        h = h_floor(h);
    }
    CK(ctx->keys0.ensure(n));
    CK(ctx->keys1.ensure(n));
    CK(ctx->vals0.ensure(n));
    CK(ctx->vals1.ensure(n));
    CK(ctx->sort_scratch.ensure(rs_scratch_elems(n)));
    pm.mark("bbox+alloc");

    GridDims gd;
    int dim0[3];
    int sorted_in = 0;
    unsigned long long occupied = 0;
    // The cell edge is tuned on the number of occupied cells, counted with one bit per cell (no sort): a few passes over
    // the points at most, then ONE sort with the edge that stands.
    for (int trial = 0;; ++trial) {
        dims_for(ext, h, dim0);
        for (int a = 0; a < 3; ++a) gd.o[a] = mn[a];
        gd.h = h;
        gd.inv_h = 1.0 / h;
        gd.nx = dim0[0];
        gd.ny = dim0[1];
        gd.nz = dim0[2];
        const size_t ncell = (size_t)dim0[0] * dim0[1] * dim0[2];
        CK(ctx->occ_bits.ensure(ncell / 32 + 2));
        CK(cudaMemsetAsync(ctx->occ_bits.p, 0, (ncell / 32 + 1) * sizeof(uint32_t), s));
        CK(cudaMemsetAsync(ctx->d_counter, 0, sizeof(unsigned long long), s));
        cell_occupancy_kernel<<<std::min<unsigned>(2048u, grid_for(n, 256)), 256, 0, s>>>(ctx->model_raw.p, (uint32_t)n, gd,
                                                                                          ctx->occ_bits.p, ctx->d_counter);
        CKL();
        ++ctx->launches;
        CK(cudaMemcpyAsync(&occupied, ctx->d_counter, sizeof(occupied), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        pm.mark("trial");
        if (!auto_h || trial >= 4 || occupied == 0) break;
        const double avg = (double)n / (double)occupied;
        if (avg >= 0.6 * ctx->opt_target && avg <= 1.6 * ctx->opt_target) break;
        const double h_new = h_floor(h * sqrt(ctx->opt_target / avg));
        if (fabs(h_new - h) <= 1e-3 * h) break; // clamped by the cell budget
        h = h_new;
    }
    {
        const size_t ncell = (size_t)dim0[0] * dim0[1] * dim0[2];
        int bits = 1;
        while (((size_t)1 << bits) < ncell) ++bits;
        cell_keys_kernel<<<grid_for(n, 256), 256, 0, s>>>(ctx->model_raw.p, (uint32_t)n, gd, ctx->keys0.p,
                                                          ctx->vals0.p);
        CKL();
        sorted_in = radix_sort<uint32_t, true>(ctx->keys0.p, ctx->vals0.p, ctx->keys1.p, ctx->vals1.p, n, bits,
                                               ctx->sort_scratch.p, s, &ctx->launches);
        CKL();
        pm.mark("sort");
    }
    ctx->gdims = gd;
    const uint32_t *keys_sorted = sorted_in ? ctx->keys1.p : ctx->keys0.p;
    CK(ctx->pkey.ensure(n));
    CK(cudaMemcpyAsync(ctx->pkey.p, keys_sorted, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    const uint32_t *perm = sorted_in ? ctx->vals1.p : ctx->vals0.p;
    const size_t ncell = (size_t)dim0[0] * dim0[1] * dim0[2];

    CK(ctx->pts.ensure(n + ICPB_PAD));
    CK(ctx->ptsf.ensure(n + ICPB_PAD));
    CK(ctx->cell_start.ensure(ncell + 1 + scan_scratch_elems(ncell + 1)));
    CK(cudaMemsetAsync(ctx->cell_start.p, 0, (ncell + 1) * sizeof(uint32_t), s));
    CK(ctx->cell_bits.ensure(ncell + 1));
    CK(cudaMemsetAsync(ctx->cell_bits.p, 0, ncell * sizeof(unsigned long long), s));
    gather_points_kernel<<<grid_for(n, 256), 256, 0, s>>>(ctx->model_raw.p, keys_sorted, perm, (uint32_t)n, gd,
                                                          ctx->pts.p, ctx->ptsf.p, ctx->cell_start.p, ctx->cell_bits.p);
    pad_tail_kernel<<<1, 32, 0, s>>>(ctx->pts.p, ctx->ptsf.p, (uint32_t)n);
    CKL();
    exclusive_scan_u32(ctx->cell_start.p, ctx->cell_start.p, ncell + 1, ctx->cell_start.p + ncell + 1, s,
                       &ctx->launches);
    CKL();
    if (ctx->model_ean) {
        CK(ctx->npts.ensure(n));
        gather_ean_kernel<<<grid_for(n, 256), 256, 0, s>>>(ctx->model_nrm.p, ctx->model_edge.p, perm, (uint32_t)n, ctx->npts.p);
        CKL();
    }

    pm.mark("gather+prefix");
    // pyramid of node words (occupancy + tight boxes) and the level-0 cell boxes
    GridView &g = ctx->gv;
    memset(&g, 0, sizeof(g));
    for (int a = 0; a < 3; ++a) {
        g.o[a] = mn[a];
        g.dim[0][a] = dim0[a];
    }
    g.h = h;
    g.inv_h = 1.0 / h;
    g.eps = 1e-9 * h;
    {
        int rc = build_masks(ctx);
        if (rc != ICPB_OK) return rc;
    }
    pm.mark("masks");
    const int L = g.L;
    g.cell_start = ctx->cell_start.p;
    g.pts = ctx->pts.p;
    g.ptsf = ctx->ptsf.p;
    g.n = (uint32_t)n;
    CK(cudaEventRecord(e1, s));
    CK(cudaStreamSynchronize(s));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    icpb_grid_info &gi = ctx->ginfo;
    for (int a = 0; a < 3; ++a) {
        gi.origin[a] = mn[a];
        gi.dims[a] = (uint32_t)dim0[a];
    }
    gi.cell_size = h;
    gi.levels = (uint32_t)L;
    gi.n_cells = ncell;
    gi.n_points = n;
    gi.n_occupied = occupied;
    gi.build_ms = ms;
    gi.merges = ctx->merges;
    ctx->last_valid = false; // nn_pos of an earlier align refers to the old layout
    ctx->grid_dirty = false;
    return ICPB_OK;
}

int icpb_model_build(icpb_ctx *ctx, icpb_grid_info *info)
{
    if (!ctx) return ICPB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (ctx->grid_dirty) {
        int rc = build_grid(ctx);
        if (rc != ICPB_OK) return rc;
    }
    if (info) *info = ctx->ginfo;
    return ICPB_OK;
}

// ---------------------------------------------------------------------------
// source
// ---------------------------------------------------------------------------
// the caller's host buffer may go away after an upload call returns: wait for the copies (icpb_align waits once, at the
// end of the align)
static int upload_wait(icpb_ctx *ctx)
{
    if (!ctx->upload_pending) return ICPB_OK;
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev_up0, ctx->ev_up1);
    ctx->timing.upload_ms = ms;
    ctx->upload_pending = false;
    return ICPB_OK;
}

// Host vertices -> stage (device AoS) in PointcloudAdapter's visiting order (K0).  density != 1: the whole array is
// copied and the visited vertices are gathered ON THE DEVICE from the closed form of the density stepping (density.h);
// absurd densities (more segments than the table holds) fall back to the host loop.  *m_out = visited vertices.
static int stage_visited(icpb_ctx *ctx, const double *xyz, size_t n, double density, size_t *m_out)
{
    cudaStream_t s = ctx->stream;
    if (density == 1.0 || n == 0) {
        CK(ctx->stage.ensure(3 * std::max<size_t>(n, 1)));
        if (n) CK(cudaMemcpyAsync(ctx->stage.p, xyz, 3 * n * sizeof(double), cudaMemcpyHostToDevice, s));
        *m_out = n;
        return ICPB_OK;
    }
    unsigned long long visited = 0;
    const int ns = icpb_density_segments(n, density, ctx->h_dseg, ICPB_DENSITY_MAX_SEGS, &visited);
    if (ns < 0) {
        std::vector<double> tmp;
        size_t m = 0;
        const double *src = density_gather(xyz, n, density, tmp, &m);
        CK(ctx->stage.ensure(3 * std::max<size_t>(m, 1)));
        if (m) CK(cudaMemcpyAsync(ctx->stage.p, src, 3 * m * sizeof(double), cudaMemcpyHostToDevice, s));
        CK(cudaStreamSynchronize(s)); // tmp goes away
        *m_out = m;
        return ICPB_OK;
    }
    const size_t m = (size_t)visited;
    CK(ctx->stage_full.ensure(3 * n));
    CK(ctx->stage.ensure(3 * std::max<size_t>(m, 1)));
    CK(ctx->d_dseg.ensure(ICPB_DENSITY_MAX_SEGS));
    CK(cudaMemcpyAsync(ctx->stage_full.p, xyz, 3 * n * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->d_dseg.p, ctx->h_dseg, (size_t)std::max(ns, 1) * sizeof(DensitySeg), cudaMemcpyHostToDevice, s));
    if (m) {
        density_gather_kernel<<<grid_for(m, 256), 256, 0, s>>>(ctx->stage_full.p, ctx->d_dseg.p, ns, visited, ctx->stage.p);
        CKL();
        ++ctx->launches;
    }
    *m_out = m;
    return ICPB_OK;
}

// stage (m vertices, device AoS, visiting order) -> the resident source: Morton order of the source in its own frame
// (rigid motions keep neighbours together, so one sort per upload serves every loop body; src_perm[] maps back to the
// visiting order), SoA coordinates.  Nothing is read back: the bounding box stays on the device.
static int upload_points(icpb_ctx *ctx, size_t m, bool wait)
{
    cudaStream_t s = ctx->stream;
    if (m >= 0xFFFFFFF0ull) return ICPB_ERR_ARG;
    CK(ctx->sx.ensure(std::max<size_t>(m, 1)));
    CK(ctx->sy.ensure(std::max<size_t>(m, 1)));
    CK(ctx->sz.ensure(std::max<size_t>(m, 1)));
    CK(ctx->nn_pos.ensure(std::max<size_t>(m, 1)));
    CK(ctx->nn_d.ensure(std::max<size_t>(m, 1)));
    CK(ctx->nn_lb.ensure(std::max<size_t>(m, 1)));
    CK(ctx->src_perm.ensure(std::max<size_t>(m, 1)));
    if (m) {
        if (ctx->opt_sort != 0.0 && m > 1024) {
            const unsigned nbb = std::min<unsigned>(1024u, grid_for(m, 256));
            CK(ctx->bbox_part.ensure((size_t)nbb * 6 + 8));
            CK(ctx->keys0.ensure(m));
            CK(ctx->keys1.ensure(m));
            CK(ctx->vals0.ensure(m));
            CK(ctx->vals1.ensure(m));
            CK(ctx->sort_scratch.ensure(rs_scratch_elems(m)));
            double *bb = ctx->bbox_part.p + (size_t)nbb * 6;
            bbox_kernel<<<nbb, 256, 0, s>>>(ctx->stage.p, m, ctx->bbox_part.p);
            const int mbits = (int)std::max(4.0, std::min(ctx->opt_morton_bits, 10.0)); // Morton cells per axis: 2^mbits
            bbox_final_kernel<<<1, 256, 0, s>>>(ctx->bbox_part.p, nbb, bb, (double)(1 << mbits));
            morton_keys_kernel<<<grid_for(m, 256), 256, 0, s>>>(ctx->stage.p, (uint32_t)m, bb, ctx->keys0.p, ctx->vals0.p,
                                                               ctx->opt_hilbert != 0.0 ? mbits : 0);
            CKL();
            const int w = radix_sort<uint32_t, true>(ctx->keys0.p, ctx->vals0.p, ctx->keys1.p, ctx->vals1.p, m, 3 * mbits,
                                                     ctx->sort_scratch.p, s, &ctx->launches);
            CKL();
            CK(cudaMemcpyAsync(ctx->src_perm.p, w ? ctx->vals1.p : ctx->vals0.p, m * sizeof(uint32_t),
                               cudaMemcpyDeviceToDevice, s));
            gather_soa_kernel<<<grid_for(m, 256), 256, 0, s>>>(ctx->stage.p, ctx->src_perm.p, (uint32_t)m, ctx->sx.p,
                                                              ctx->sy.p, ctx->sz.p);
            CKL();
            ctx->launches += 4;
        } else {
            aos_to_soa_kernel<<<grid_for(m, 256), 256, 0, s>>>(ctx->stage.p, m, ctx->sx.p, ctx->sy.p, ctx->sz.p);
            iota_kernel<<<grid_for(m, 256), 256, 0, s>>>(ctx->src_perm.p, (uint32_t)m);
            CKL();
            ctx->launches += 2;
        }
    }
    CK(cudaEventRecord(ctx->ev_up1, s));
    ctx->src_n = m;
    ctx->have_source = true;
    ctx->src_ean = false;
    ctx->last_valid = false;
    ctx->pd_valid = false;
    ctx->upload_pending = true;
    if (wait) return upload_wait(ctx);
    return ICPB_OK;
}

static int source_upload(icpb_ctx *ctx, const double *xyz, size_t n, const double T[16], double density, bool wait)
{
    NvtxRange nvtx_range("icpb: source upload (K0 + Hilbert sort)");
    if (!ctx || (!xyz && n) || !T || !(density > 0.0)) return ICPB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventRecord(ctx->ev_up0, ctx->stream));
    size_t m = 0;
    int rc = stage_visited(ctx, xyz, n, density, &m);
    if (rc != ICPB_OK) return rc;
    cm_to_aff12(T, ctx->Cl2g);
    ctx->adapter_size = (size_t)((double)n * density); // PointcloudAdapter::size() (icp/icp.hpp:143-146)
    return upload_points(ctx, m, wait);
}

int icpb_source_upload(icpb_ctx *ctx, const double *xyz, size_t n, const double T[16], double density)
{
    return source_upload(ctx, xyz, n, T, density, true);
}

int icpb_source_upload_ean(icpb_ctx *ctx, const double *xyz, const double *normals, const int32_t *attrs, size_t n,
                           const double T[16], double density)
{
    if (!normals || !attrs) return ICPB_ERR_ARG;
    int rc = icpb_source_upload(ctx, xyz, n, T, density);
    if (rc != ICPB_OK) return rc;
    const size_t m = ctx->src_n;
    cudaStream_t s = ctx->stream;
    std::vector<double> tn;
    std::vector<int32_t> ta;
    const double *nsrc = density_gather_t(normals, n, 3, density, tn);
    const int32_t *asrc = density_gather_t(attrs, n, 1, density, ta);
    CK(ctx->snx.ensure(std::max<size_t>(m, 1)));
    CK(ctx->sny.ensure(std::max<size_t>(m, 1)));
    CK(ctx->snz.ensure(std::max<size_t>(m, 1)));
    CK(ctx->sedge.ensure(std::max<size_t>(m, 1)));
    CK(ctx->sedge_tmp.ensure(std::max<size_t>(m, 1)));
    CK(ctx->stage_i.ensure(std::max<size_t>(m, 1)));
    if (m) { // same (Morton) order as the vertices: src_perm maps slot -> visiting index
        CK(cudaMemcpyAsync(ctx->stage.p, nsrc, 3 * m * sizeof(double), cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(ctx->stage_i.p, asrc, m * sizeof(int32_t), cudaMemcpyHostToDevice, s));
        gather_soa_kernel<<<grid_for(m, 256), 256, 0, s>>>(ctx->stage.p, ctx->src_perm.p, (uint32_t)m, ctx->snx.p, ctx->sny.p,
                                                          ctx->snz.p);
        edge_flags_kernel<<<grid_for(m, 256), 256, 0, s>>>(ctx->stage_i.p, m, ctx->sedge_tmp.p);
        gather_u8_kernel<<<grid_for(m, 256), 256, 0, s>>>(ctx->sedge_tmp.p, ctx->src_perm.p, (uint32_t)m, ctx->sedge.p);
        CKL();
        CK(cudaStreamSynchronize(s));
        ctx->launches += 3;
    }
    ctx->src_ean = true;
    return ICPB_OK;
}

int icpb_source_upload_shard(icpb_ctx *ctx, const double *xyz, size_t n_local, size_t adapter_size_global,
                             const double T[16])
{
    if (!ctx || (!xyz && n_local) || !T) return ICPB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    cm_to_aff12(T, ctx->Cl2g);
    ctx->adapter_size = adapter_size_global;
    CK(cudaEventRecord(ctx->ev_up0, ctx->stream));
    size_t m = 0;
    int rc = stage_visited(ctx, xyz, n_local, 1.0, &m);
    if (rc != ICPB_OK) return rc;
    return upload_points(ctx, m, true);
}

// ---------------------------------------------------------------------------
// trim: the K4 launch sequence (shared by align and icpb_trim)
// ---------------------------------------------------------------------------
__global__ void tie_count_kernel(const double *__restrict__ nn_d, uint32_t n, IcpState *st)
{
    if (st->done || !st->need_tie) return;
    const unsigned long long tau_bits = (unsigned long long)__double_as_longlong(st->tau);
    unsigned c = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        c += ((unsigned long long)__double_as_longlong(nn_d[i]) == tau_bits) ? 1u : 0u;
    c = __reduce_add_sync(0xFFFFFFFFu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&st->tie_local, (unsigned long long)c);
}
__global__ void tie_publish_kernel(IcpState *st, unsigned long long *slot)
{
    *slot = (st->done || !st->need_tie) ? 0ull : st->tie_local;
}
// ties are kept in global pair order: lower ranks first, then by local position
__global__ void tie_quota_kernel(IcpState *st, const unsigned long long *gathered, int rank)
{
    if (st->done || !st->need_tie) return;
    unsigned long long before = 0;
    for (int r = 0; r < rank; ++r) before += gathered[r];
    const unsigned long long local = st->tie_local;
    unsigned long long q = st->quota > before ? st->quota - before : 0ull;
    if (q > local) q = local;
    st->quota = q;
    st->cnt_eq = local;
    st->tie_local = 0;
    if (q == 0) {
        st->p_cut = -1;
        st->need_tie = 0;
    } else if (q == local) {
        st->p_cut = 0x100000000ll;
        st->need_tie = 0;
    }
}

// Blocks of trim_moments_kernel that are resident together (its blocks wait for each other): what the occupancy
// calculator grants on this device, queried once.
static unsigned coresident_trim_blocks(icpb_ctx *ctx)
{
    if (ctx->tm_blocks == 0) {
        int per_sm = 0, sms = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, trim_moments_kernel, SEL_THREADS, 0) != cudaSuccess) per_sm = 1;
        ctx->tm_blocks = (unsigned)std::min(TRIM_MAX_BLOCKS, std::max(1, std::min(per_sm, 2) * std::max(sms, 1)));
    }
    return ctx->tm_blocks;
}

// Moments inside the trim launch?  Measured on B200 (tools/graph_vs_stream.py, graph mode): at 100 k source points the
// fused launch wins (79.9 vs 82.9 us per loop body), at 1 M the separate moments_kernel does (239.2 vs 244.5 us): the
// kept pairs of the bin and the final reduction are a chain of trips to the L2 in ONE block, which a grid-wide kernel
// overlaps better once it has enough work.  fuse_moments: 1 / 0 force it, -1 (default) decides by size.
static bool fuse_moments_on(const icpb_ctx *ctx, uint32_t n)
{
    if (ctx->opt_fuse_moments < 0.0) return n <= 300000u;
    return ctx->opt_fuse_moments != 0.0;
}

// moments_done (may be null): set when the launch also accumulated the moments and ran the pose step (loop bodies after
// the first on the fused path) -- no moments_kernel must follow then
static int launch_trim(icpb_ctx *ctx, uint32_t n, bool linear = false, bool *moments_done = nullptr,
                       cudaGraphConditionalHandle cond = 0, int use_cond = 0)
{
    cudaStream_t s = ctx->stream;
    const unsigned blocks = std::max(1u, std::min(grid_for(n, SEL_THREADS * 4), 148u * (unsigned)std::max(1.0, std::min(ctx->opt_trim_bpsm, 32.0))));
    const int fused = (ctx->n_ranks == 1 || ctx->use_p2p) ? 1 : 0;
    if (moments_done) *moments_done = false;
    if (linear) { // the search kernel has binned the distances over [0, d_box]: one launch
        CK(ctx->cand_key.ensure(std::max<size_t>(n, 1)));
        CK(ctx->cand_idx.ensure(std::max<size_t>(n, 1)));
        const unsigned lin_blocks = std::min(blocks, coresident_trim_blocks(ctx));
        const int fuse = (moments_done && fuse_moments_on(ctx, n)) ? 1 : 0;
        CK(ctx->partials.ensure((size_t)lin_blocks * ICPB_NSUMS));
        const size_t seg_total = (size_t)lin_blocks * (SEL_THREADS / 32) * trim_seg_cap(n, lin_blocks);
        CK(ctx->seg_key.ensure(seg_total));
        CK(ctx->seg_idx.ensure(seg_total));
        CK(ctx->seg_count.ensure((size_t)TRIM_MAX_BLOCKS * (SEL_THREADS / 32)));
        trim_moments_kernel<<<lin_blocks, SEL_THREADS, 0, s>>>(ctx->gv, ctx->sx.p, ctx->sy.p, ctx->sz.p, ctx->nn_pos.p, ctx->nn_d.p,
                                                               ctx->src_perm.p, n, ctx->d_state, ctx->d_lhist, ctx->cand_key.p,
                                                               ctx->cand_idx.p, ctx->seg_key.p, ctx->seg_idx.p, ctx->seg_count.p, ctx->p2p,
                                                               ctx->xscratch.p, ctx->d_lhist + LIN_BINS * LIN_STRIDE + 64,
                                                               ctx->partials.p, ctx->trace.p, ctx->trace_cap, cond, use_cond, fuse);
        CKL();
        ctx->launches += 1;
        if (moments_done) *moments_done = fuse != 0;
        return ICPB_OK;
    }
    if (fused) {
        // two grid-wide digit passes, then collect the few remaining candidates and finish on them in one block
        CK(ctx->cand_key.ensure(std::max<size_t>(n, 1)));
        CK(ctx->cand_idx.ensure(std::max<size_t>(n, 1)));
        for (int p = 0; p < 2; ++p) {
            select_kernel<<<blocks, SEL_THREADS, 0, s>>>(ctx->nn_d.p, n, ctx->d_state, ctx->d_ghist, p, 1, ctx->p2p);
            CKL();
        }
        select_collect_kernel<<<blocks, SEL_THREADS, 0, s>>>(ctx->nn_d.p, ctx->src_perm.p, n, ctx->d_state, ctx->d_ghist,
                                                             ctx->cand_key.p, ctx->cand_idx.p,
                                                             (unsigned int *)(ctx->d_counter + 2), ctx->p2p, ctx->xscratch.p);
        CKL();
        ctx->launches += 3;
        return ICPB_OK;
    }
    if (!fused) // pairs found on all ranks
        NK(g_nccl.AllReduce(&ctx->d_state->found, &ctx->d_state->found, 1, NCCL_UINT64, NCCL_SUM, ctx->comm, s));
    for (int p = 0; p < SEL_D_PASSES; ++p) {
        select_kernel<<<blocks, SEL_THREADS, 0, s>>>(ctx->nn_d.p, n, ctx->d_state, ctx->d_ghist, p, fused, ctx->p2p);
        CKL();
        ++ctx->launches;
        if (!fused) {
            NK(g_nccl.AllReduce(ctx->d_ghist, ctx->d_ghist, SEL_MAX_BINS, NCCL_UINT32, NCCL_SUM, ctx->comm, s));
            select_pick_kernel<<<1, SEL_THREADS, 0, s>>>(ctx->d_state, ctx->d_ghist, p, 0);
            CKL();
            ++ctx->launches;
        }
    }
    if (!fused) { // split the tie quota over the ranks; the tie passes below are then rank-local
        tie_count_kernel<<<blocks, SEL_THREADS, 0, s>>>(ctx->nn_d.p, n, ctx->d_state);
        tie_publish_kernel<<<1, 1, 0, s>>>(ctx->d_state, ctx->d_gather + 32);
        CKL();
        NK(g_nccl.AllGather(ctx->d_gather + 32, ctx->d_gather, 1, NCCL_UINT64, ctx->comm, s));
        tie_quota_kernel<<<1, 1, 0, s>>>(ctx->d_state, ctx->d_gather, ctx->rank);
        CKL();
        ctx->launches += 3;
    }
    tie_select_kernel<<<1, 1024, 0, s>>>(ctx->nn_d.p, ctx->src_perm.p, n, ctx->d_state, ctx->d_ghist);
    CKL();
    ++ctx->launches;
    return ICPB_OK;
}

// ---------------------------------------------------------------------------
// align
// ---------------------------------------------------------------------------
// One Trimmed::_align in two steps: align_enqueue puts the whole loop on the context's stream (graph mode: one graph
// launch, nothing waited for), align_collect waits for it and reads the result.  A single context runs them back to
// back; icpb_multi_align_resident enqueues on every device first, so that the ranks' fused exchanges meet.
static int align_enqueue(icpb_ctx *ctx, const icpb_params *prm, bool graph_only)
{
    NvtxRange nvtx_range("icpb: _align enqueue (K3-K6 loop)");
    if (!ctx || !prm) return ICPB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    ctx->run.pending = false;
    ctx->pd_valid = false;
    if (!ctx->have_source) return ICPB_ERR_NO_SOURCE;
    if (ctx->grid_dirty) {
        int rc = build_grid(ctx);
        if (rc != ICPB_OK) return rc;
    }
    if (ctx->src_ean && ctx->model_n > 0 && !ctx->model_ean) return ICPB_ERR_ARG; // filter needs model normals
    cudaStream_t s = ctx->stream;
    const uint32_t n = (uint32_t)ctx->src_n;
    ctx->gv.block_levels = (int)std::max(-1.0, std::min(ctx->opt_block_levels, 1.0));
    EanView ean;
    memset(&ean, 0, sizeof(ean));
    if (ctx->src_ean && ctx->model_n > 0) {
        ean.mn = ctx->npts.p;
        ean.snx = ctx->snx.p;
        ean.sny = ctx->sny.p;
        ean.snz = ctx->snz.p;
        ean.sedge = ctx->sedge.p;
        ean.on = 1;
    }
    // n_po = measurement.size() * overlap  (icp/icp.hpp:472), size() from icp/icp.hpp:143-146
    const size_t n_po = (size_t)((double)ctx->adapter_size * prm->overlap);
    const size_t max_iter = prm->max_iter;

    const unsigned mom_blocks = std::max(1u, std::min(grid_for(n, MOM_THREADS), 148u * (unsigned)std::max(1.0, std::min(ctx->opt_mom_bpsm, 32.0))));
    CK(ctx->partials.ensure((size_t)std::max(mom_blocks, coresident_trim_blocks(ctx)) * ICPB_NSUMS));
    ctx->trace_cap = (int)std::min<size_t>(std::max<size_t>(max_iter, 1), 65536);
    CK(ctx->trace.ensure((size_t)ctx->trace_cap * ICPB_TRACE_COLS_DEV));
    const bool profile = ctx->opt_profile != 0.0;
    if (profile && ctx->events.empty()) {
        ctx->events.resize((size_t)EVENT_ITERS * 4);
        for (auto &e : ctx->events) CK(cudaEventCreate(&e));
    }
    const int fused = (ctx->n_ranks == 1 || ctx->use_p2p) ? 1 : 0;
    const int run_ahead = (int)std::max(1.0, std::min(ctx->opt_run_ahead, 32.0));
    const unsigned long long launches0 = ctx->launches;

    Aff12 A;
    memcpy(A.a, ctx->Cl2g, sizeof(A.a));
    cudaGraphConditionalHandle cond = 0;
    int use_cond = 0;
    const int use_warm = ctx->opt_warm != 0.0 ? 1 : 0;
    // one loop body of Trimmed::_align: search, trim, moments + pose
    NNTune tune;
    tune.reach_cells = (float)std::max(0.0, ctx->opt_skip_reach);
    tune.floor_cells = (float)std::max(0.0, ctx->opt_skip_floor);
    tune.skip = (use_warm && ctx->opt_skip != 0.0) ? 1 : 0;
    // lazy search needs the skip state (winner + bound per query); the edge/normal filter decides pairs per winner
    const double lazy_margin = (tune.skip && ctx->opt_lazy != 0.0 && !ean.on) ? std::max(-0.9, ctx->opt_lazy_margin) : -1.0;
    const unsigned search_blocks = grid_for(std::max<uint32_t>(n, 1), SEARCH_THREADS);
    // loop bodies after the first trim in ONE launch on the histogram the search kernel fills (one GPU, or several over
    // peer memory: two exchanges inside that launch; the NCCL transport keeps the radix passes)
    const bool linear_ok = (ctx->n_ranks == 1 || ctx->use_p2p) && ctx->opt_linear_trim != 0.0;
    auto enqueue_iteration = [&](bool linear) -> int {
        uint32_t *lh = linear ? ctx->d_lhist : nullptr;
        if (ean.on)
            search_kernel<true><<<search_blocks, SEARCH_THREADS, 0, s>>>(ctx->gv, ctx->sx.p, ctx->sy.p, ctx->sz.p, n,
                                                                         ctx->d_state, ctx->nn_pos.p, ctx->nn_d.p,
                                                                         ctx->nn_lb.p, use_warm, tune, ean, lh);
        else
            search_kernel<false><<<search_blocks, SEARCH_THREADS, 0, s>>>(ctx->gv, ctx->sx.p, ctx->sy.p, ctx->sz.p, n,
                                                                          ctx->d_state, ctx->nn_pos.p, ctx->nn_d.p,
                                                                          ctx->nn_lb.p, use_warm, tune, ean, lh);
        CKL();
        ctx->launches += 1;
        return ICPB_OK;
    };
    size_t it = 0, checked = 0, timed_iters = 0;
    const bool use_graph = ctx->opt_graph != 0.0 && !profile && fused;
    if (graph_only && !use_graph) {
        ctx->err = "several devices in one process need the graph loop (options graph = 1, profile = 0, peer-memory transport)";
        return ICPB_ERR_ARG;
    }
    if (use_graph) {
        // ---- the whole loop as ONE graph launch: init -> WHILE(!done){search, trim, moments+pose} ----
        CK(ctx->cand_key.ensure(std::max<size_t>(n, 1)));
        CK(ctx->cand_idx.ensure(std::max<size_t>(n, 1)));
        if (linear_ok) {
            const unsigned tb = std::max(1u, std::min(grid_for(n, SEL_THREADS * 4), 148u * (unsigned)std::max(1.0, std::min(ctx->opt_trim_bpsm, 32.0))));
            const unsigned lb = std::min(tb, coresident_trim_blocks(ctx));
            const size_t seg_total = (size_t)lb * (SEL_THREADS / 32) * trim_seg_cap(n, lb);
            CK(ctx->seg_key.ensure(seg_total));
            CK(ctx->seg_idx.ensure(seg_total));
            CK(ctx->seg_count.ensure((size_t)TRIM_MAX_BLOCKS * (SEL_THREADS / 32)));
        }
        std::vector<unsigned char> key;
        auto add_key = [&](const void *p, size_t bytes) { key.insert(key.end(), (const unsigned char *)p, (const unsigned char *)p + bytes); };
        const void *ptrs[] = {ctx->sx.p, ctx->sy.p, ctx->sz.p, ctx->nn_pos.p, ctx->nn_d.p, ctx->nn_lb.p, ctx->src_perm.p, ctx->partials.p,
                              ctx->trace.p, ctx->cand_key.p, ctx->cand_idx.p, ctx->d_state, ctx->d_ghist, ctx->d_lhist,
                              ctx->seg_key.p, ctx->seg_idx.p, ctx->seg_count.p};
        add_key(&n, sizeof(n));
        add_key(&ctx->gv, sizeof(ctx->gv));
        add_key(ptrs, sizeof(ptrs));
        add_key(&ean, sizeof(ean));
        add_key(&ctx->p2p, sizeof(ctx->p2p));
        add_key(&ctx->trace_cap, sizeof(ctx->trace_cap));
        add_key(&use_warm, sizeof(use_warm));
        add_key(&tune, sizeof(tune));
        add_key(&linear_ok, sizeof(linear_ok));
        add_key(&mom_blocks, sizeof(mom_blocks));
        add_key(&ctx->opt_trim_bpsm, sizeof(double));
        add_key(&ctx->opt_fuse_moments, sizeof(double));
        void *init_args[13];
        unsigned long long a_max_iter = max_iter, a_n_po = n_po, a_n = n, a_off = 0;
        double a_mse = prm->min_mse, a_diff = prm->min_mse_diff, a_ov = prm->overlap, a_lazy = lazy_margin;
        int one = 1;
        if (key != ctx->graph_key || !ctx->graph_exec) {
            if (ctx->graph_exec) cudaGraphExecDestroy(ctx->graph_exec);
            if (ctx->graph) cudaGraphDestroy(ctx->graph);
            ctx->graph_exec = nullptr;
            ctx->graph = nullptr;
            ctx->graph_key.clear();
            CK(cudaGraphCreate(&ctx->graph, 0));
            CK(cudaGraphConditionalHandleCreate(&ctx->graph_cond, ctx->graph, 1, cudaGraphCondAssignDefault));
            init_args[0] = &ctx->d_state; init_args[1] = &A; init_args[2] = &a_max_iter; init_args[3] = &a_mse;
            init_args[4] = &a_diff; init_args[5] = &a_ov; init_args[6] = &a_n_po; init_args[7] = &a_n; init_args[8] = &a_off;
            init_args[9] = &ctx->graph_cond; init_args[10] = &one; init_args[11] = &a_lazy; init_args[12] = &ctx->d_lhist;
            cudaKernelNodeParams kp;
            memset(&kp, 0, sizeof(kp));
            kp.func = (void *)init_state_kernel;
            kp.gridDim = dim3(1);
            kp.blockDim = dim3(256);
            kp.kernelParams = init_args;
            CK(cudaGraphAddKernelNode(&ctx->graph_init_node, ctx->graph, nullptr, 0, &kp));
            cond = ctx->graph_cond;
            use_cond = 1;
            auto capture_body = [&](bool linear) -> int {
                int rc = enqueue_iteration(linear);
                bool moments_done = false;
                if (rc == ICPB_OK) rc = launch_trim(ctx, n, linear, &moments_done, cond, use_cond);
                if (rc == ICPB_OK && !moments_done) {
                    moments_kernel<<<mom_blocks, MOM_THREADS, 0, s>>>(ctx->gv, ctx->sx.p, ctx->sy.p, ctx->sz.p, n, ctx->d_state,
                                                                      ctx->nn_pos.p, ctx->nn_d.p, ctx->src_perm.p,
                                                                      ctx->partials.p, ctx->trace.p, ctx->trace_cap, fused,
                                                                      ctx->p2p, cond, use_cond);
                }
                return rc;
            };
            // the first loop body (d_box = inf: radix trim) as plain nodes behind the init node, when the later bodies differ
            cudaGraphNode_t while_dep = ctx->graph_init_node;
            if (linear_ok) {
                CK(cudaStreamBeginCaptureToGraph(s, ctx->graph, &ctx->graph_init_node, nullptr, 1, cudaStreamCaptureModeThreadLocal));
                int rc0 = capture_body(false);
                cudaStreamCaptureStatus cst;
                const cudaGraphNode_t *deps = nullptr;
                size_t ndeps = 0;
                cudaError_t ci = cudaStreamGetCaptureInfo(s, &cst, nullptr, nullptr, &deps, &ndeps);
                if (ci == cudaSuccess && ndeps >= 1) while_dep = deps[ndeps - 1];
                cudaGraph_t cap0 = nullptr;
                cudaError_t ce0 = cudaStreamEndCapture(s, &cap0);
                if (rc0 != ICPB_OK) return rc0;
                if (ci != cudaSuccess || ndeps != 1) return ctx->fail(ci != cudaSuccess ? ci : cudaErrorUnknown, "capture of the first loop body", __LINE__);
                if (ce0 != cudaSuccess) return ctx->fail(ce0, "cudaStreamEndCapture", __LINE__);
            }
            cudaGraphNodeParams cp = {cudaGraphNodeTypeConditional};
            cp.type = cudaGraphNodeTypeConditional;
            cp.conditional.handle = ctx->graph_cond;
            cp.conditional.type = cudaGraphCondTypeWhile;
            cp.conditional.size = 1;
            cudaGraphNode_t cond_node;
            CK(cudaGraphAddNode(&cond_node, ctx->graph, &while_dep, 1, &cp));
            cudaGraph_t body = cp.conditional.phGraph_out[0];
            CK(cudaStreamBeginCaptureToGraph(s, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
            int rc = capture_body(linear_ok);
            cudaGraph_t captured = nullptr;
            cudaError_t ce = cudaStreamEndCapture(s, &captured);
            if (rc != ICPB_OK) return rc;
            if (ce != cudaSuccess) return ctx->fail(ce, "cudaStreamEndCapture", __LINE__);
            CK(cudaGraphInstantiate(&ctx->graph_exec, ctx->graph, 0));
            ctx->graph_key = key;
        }
        // this align's parameters go into the init node
        init_args[0] = &ctx->d_state; init_args[1] = &A; init_args[2] = &a_max_iter; init_args[3] = &a_mse;
        init_args[4] = &a_diff; init_args[5] = &a_ov; init_args[6] = &a_n_po; init_args[7] = &a_n; init_args[8] = &a_off;
        init_args[9] = &ctx->graph_cond; init_args[10] = &one; init_args[11] = &a_lazy; init_args[12] = &ctx->d_lhist;
        cudaKernelNodeParams kp;
        memset(&kp, 0, sizeof(kp));
        kp.func = (void *)init_state_kernel;
        kp.gridDim = dim3(1);
        kp.blockDim = dim3(256);
        kp.kernelParams = init_args;
        CK(cudaGraphExecKernelNodeSetParams(ctx->graph_exec, ctx->graph_init_node, &kp));
        CK(cudaEventRecord(ctx->ev_begin, s));
        CK(cudaGraphLaunch(ctx->graph_exec, s));
    } else {
    CK(cudaEventRecord(ctx->ev_begin, s));
    init_state_kernel<<<1, 256, 0, s>>>(ctx->d_state, A, max_iter, prm->min_mse, prm->min_mse_diff, prm->overlap, n_po,
                                        n, 0ull, 0, 0, lazy_margin, ctx->d_lhist);
    CKL();
    ++ctx->launches;
    for (int k = 0; k < FLAG_RING; ++k) ctx->h_flags[k] = -1;

    bool stop = false;
    // a loop body whose lazy trim did not verify is run again (icpb_iteration_finish): allow for a few more than max_iter
    const size_t max_bodies = max_iter + (lazy_margin > -1.0 ? std::min<size_t>(max_iter, 64) : 0);
    for (; it < max_bodies && !stop; ++it) {
        // never run more than run_ahead iterations ahead of the device-side loop condition
        if (it >= (size_t)run_ahead) {
            CK(cudaEventSynchronize(ctx->ring[(it - run_ahead) % ctx->ring.size()]));
        }
        while (checked < it && ctx->h_flags[checked % FLAG_RING] != -1) {
            if (ctx->h_flags[checked % FLAG_RING] == 1) stop = true;
            ctx->h_flags[checked % FLAG_RING] = -1;
            ++checked;
        }
        if (stop) break;
        if (it - checked >= (size_t)FLAG_RING - 1) CK(cudaStreamSynchronize(s));
        const bool timed = profile && it < (size_t)EVENT_ITERS;
        cudaEvent_t *ev = timed ? &ctx->events[it * 4] : nullptr;
        if (timed) CK(cudaEventRecord(ev[0], s));
        const bool linear = linear_ok && it > 0; // the first body's d_box is inf
        bool moments_done = false;
        {
            int rc = enqueue_iteration(linear);
            if (rc != ICPB_OK) return rc;
        }
        if (timed) CK(cudaEventRecord(ev[1], s));
        {
            int rc = launch_trim(ctx, n, linear, &moments_done, cond, use_cond);
            if (rc != ICPB_OK) return rc;
        }
        if (timed) CK(cudaEventRecord(ev[2], s));
        if (!moments_done) {
            moments_kernel<<<mom_blocks, MOM_THREADS, 0, s>>>(ctx->gv, ctx->sx.p, ctx->sy.p, ctx->sz.p, n, ctx->d_state,
                                                              ctx->nn_pos.p, ctx->nn_d.p, ctx->src_perm.p, ctx->partials.p,
                                                              ctx->trace.p, ctx->trace_cap, fused, ctx->p2p, cond, use_cond);
            CKL();
            ++ctx->launches;
        }
        if (!fused) {
            NK(g_nccl.AllReduce(ctx->d_state->sums, ctx->d_state->sums, ICPB_NSUMS, NCCL_FLOAT64, NCCL_SUM, ctx->comm,
                                s));
            pose_kernel<<<1, 1, 0, s>>>(ctx->d_state, ctx->trace.p, ctx->trace_cap);
            CKL();
            ++ctx->launches;
        }
        if (timed) {
            CK(cudaEventRecord(ev[3], s));
            ++timed_iters;
        }
        CK(cudaMemcpyAsync(&ctx->h_flags[it % FLAG_RING], &ctx->d_state->done, sizeof(int), cudaMemcpyDeviceToHost, s));
        CK(cudaEventRecord(ctx->ring[it % ctx->ring.size()], s));
    }
    }
    CK(cudaMemcpyAsync(ctx->h_state, ctx->d_state, sizeof(IcpState), cudaMemcpyDeviceToHost, s));
    CK(cudaEventRecord(ctx->ev_end, s));
    ctx->run.launches0 = launches0;
    ctx->run.use_graph = use_graph;
    ctx->run.linear_ok = linear_ok;
    ctx->run.n = n;
    ctx->run.timed_iters = timed_iters;
    ctx->run.pending = true;
    return ICPB_OK;
}

static int align_collect(icpb_ctx *ctx, icpb_result *out)
{
    NvtxRange nvtx_range("icpb: _align wait + result");
    if (!ctx || !out || !ctx->run.pending) return ICPB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    ctx->run.pending = false;
    ctx->pd_valid = false;
    cudaStream_t s = ctx->stream;
    const unsigned long long launches0 = ctx->run.launches0;
    const bool use_graph = ctx->run.use_graph, linear_ok = ctx->run.linear_ok;
    const uint32_t n = ctx->run.n;
    const size_t timed_iters = ctx->run.timed_iters;
    CK(cudaStreamSynchronize(s));

    const IcpState &hs = *ctx->h_state;
#ifdef ICPB_PHASE_STAMPS
    fprintf(stderr, "trim_moments_kernel phases (ns): block 0: hist+bin %lld, pass 1 + partial %lld | last block starts %lld after block 0's start: gather + selection %lld, kept pairs of the bin %lld, rows reduced %lld, pose %lld\n",
            (long long)(hs.stamps[1] - hs.stamps[0]), (long long)(hs.stamps[2] - hs.stamps[1]), (long long)(hs.stamps[3] - hs.stamps[0]),
            (long long)(hs.stamps[4] - hs.stamps[3]), (long long)(hs.stamps[6] - hs.stamps[4]), (long long)(hs.stamps[10] - hs.stamps[6]),
            (long long)(hs.stamps[11] - hs.stamps[10]));
#endif
    if (hs.comm_error) {
        ctx->err = "peer-memory collective timed out (a rank is missing)";
        return ICPB_ERR_NCCL;
    }
    aff12_to_cm(hs.C, out->C_global2globalnew);
    out->iter = (size_t)hs.iter;
    out->pairs = (size_t)hs.pairs;
    out->mse = hs.mse;
    out->mse_diff = hs.mse_diff;
    out->d_box = hs.d_box;
    out->overlap = hs.overlap;
    out->status = hs.early ? ICPB_ERR_NOT_ENOUGH_PAIRS : ICPB_OK;
    ctx->last_executed = hs.executed;
    ctx->last_early = hs.early != 0;
    ctx->last_valid = hs.executed > 0;
    ctx->pd_valid = ctx->last_valid;

    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev_begin, ctx->ev_end);
    ctx->timing.total_ms = ms;
    ctx->timing.search_ms = ctx->timing.trim_ms = ctx->timing.solve_ms = 0.0;
    const size_t exec_timed = std::min<size_t>(timed_iters, (size_t)hs.executed);
    ctx->iter_ms.assign(exec_timed * 3, 0.0);
    for (size_t k = 0; k < exec_timed; ++k) { // only iterations whose loop body really ran
        float a = 0.f, b = 0.f, c = 0.f;
        cudaEventElapsedTime(&a, ctx->events[k * 4], ctx->events[k * 4 + 1]);
        cudaEventElapsedTime(&b, ctx->events[k * 4 + 1], ctx->events[k * 4 + 2]);
        cudaEventElapsedTime(&c, ctx->events[k * 4 + 2], ctx->events[k * 4 + 3]);
        ctx->timing.search_ms += a;
        ctx->timing.trim_ms += b;
        ctx->timing.solve_ms += c;
        ctx->iter_ms[k * 3] = a;
        ctx->iter_ms[k * 3 + 1] = b;
        ctx->iter_ms[k * 3 + 2] = c;
    }
    ctx->timing.iterations = use_graph ? (unsigned long long)hs.executed : exec_timed;
    if (use_graph) // init + the kernels of every body that ran (first body: 5; later ones 2 with the fused trim + moments, 3 with the one-launch trim)
        ctx->launches = launches0 + 1ull + (hs.executed + hs.redone > 0 ? 5ull + (linear_ok ? (fuse_moments_on(ctx, n) ? 2ull : 3ull) : 5ull) * (unsigned long long)(hs.executed + hs.redone - 1) : 0ull);
    ctx->timing.kernels_launched = ctx->launches - launches0;
    return ICPB_OK;
}

int icpb_align_resident(icpb_ctx *ctx, const icpb_params *prm, icpb_result *out)
{
    if (!ctx || !prm || !out) return ICPB_ERR_ARG;
    int rc = align_enqueue(ctx, prm, false);
    if (rc != ICPB_OK) return rc;
    return align_collect(ctx, out);
}

int icpb_align(icpb_ctx *ctx, const double *xyz, size_t n, const double T[16], double density, const icpb_params *prm,
               icpb_result *out)
{
    int rc = source_upload(ctx, xyz, n, T, density, false); // the align below waits once, for the copies too
    if (rc != ICPB_OK) return rc;
    rc = icpb_align_resident(ctx, prm, out);
    if (ctx->upload_pending) {
        const int rw = upload_wait(ctx);
        if (rc == ICPB_OK) rc = rw;
    }
    return rc;
}

int icpb_last_timing(const icpb_ctx *ctx, icpb_timing *out)
{
    if (!ctx || !out) return ICPB_ERR_ARG;
    *out = ctx->timing;
    return ICPB_OK;
}

int icpb_iteration_times(const icpb_ctx *ctx, double *rows, size_t capacity_rows, size_t *n_rows)
{
    if (!ctx || !n_rows) return ICPB_ERR_ARG;
    const size_t n = ctx->iter_ms.size() / 3;
    *n_rows = n;
    if (!rows) return ICPB_OK;
    if (capacity_rows < n) return ICPB_ERR_CAPACITY;
    memcpy(rows, ctx->iter_ms.data(), n * 3 * sizeof(double));
    return ICPB_OK;
}

int icpb_pairs_distance(icpb_ctx *ctx, double *out, size_t capacity, size_t *n_out)
{
    if (!ctx || !n_out) return ICPB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    *n_out = 0;
    // _align returns before filling pairs_distance when pairs < MIN_PAIRS (icp/icp.hpp:475-476,500-504)
    // (they outlive addToModel / clearModel: the reference keeps them in minResult until the next align)
    if (!ctx->pd_valid || ctx->last_early) return ICPB_OK;
    cudaStream_t s = ctx->stream;
    const uint32_t n = (uint32_t)ctx->src_n;
    size_t kept = (size_t)ctx->h_state->pairs;
    *n_out = kept; // sharded align: an upper bound until the rank-local count is known (below)
    if (!out) return ICPB_OK;
    if (ctx->n_ranks == 1 && capacity < kept) return ICPB_ERR_CAPACITY;
    if (kept == 0) return ICPB_OK;
    CK(ctx->kd0.ensure(n));
    CK(ctx->kd1.ensure(n));
    CK(ctx->sort_scratch.ensure(rs_scratch_elems(n)));
    unsigned int *cnt = (unsigned int *)(ctx->d_counter + 1);
    CK(cudaMemsetAsync(cnt, 0, sizeof(unsigned int), s));
    compact_kept_kernel<<<grid_for(n, 256), 256, 0, s>>>(ctx->nn_d.p, ctx->src_perm.p, n, ctx->d_state, ctx->kd0.p, cnt);
    CKL();
    unsigned int h_cnt = 0;
    CK(cudaMemcpyAsync(&h_cnt, cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (ctx->n_ranks > 1) {
        // sharded align: `kept` counts the pairs of all ranks, this context holds its own shard's -- the rank-local kept
        // distances are returned (ascending); the caller concatenates / merges the ranks' lists if it wants them all
        kept = (size_t)h_cnt;
        *n_out = kept;
        if (capacity < kept) return ICPB_ERR_CAPACITY;
        if (kept == 0) return ICPB_OK;
    } else if ((size_t)h_cnt != kept) {
        ctx->err = "internal: kept-pair count mismatch";
        return ICPB_ERR_CUDA;
    }
    // ascending like std::sort on distances (icp/icp.cpp:26): non-negative doubles sort as u64
    const int where = radix_sort<unsigned long long, false>(ctx->kd0.p, nullptr, ctx->kd1.p, nullptr, kept, 64,
                                                            ctx->sort_scratch.p, s, &ctx->launches);
    CKL();
    CK(cudaMemcpyAsync(out, where ? ctx->kd1.p : ctx->kd0.p, kept * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return ICPB_OK;
}

int icpb_debug_pairs(icpb_ctx *ctx, uint32_t *model_idx, double *dist, uint8_t *kept, size_t capacity, size_t *n_out)
{
    if (!ctx || !n_out) return ICPB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    *n_out = 0;
    if (!ctx->last_valid) return ICPB_OK;
    const uint32_t n = (uint32_t)ctx->src_n;
    *n_out = n;
    if (!model_idx && !dist && !kept) return ICPB_OK;
    if (capacity < n) return ICPB_ERR_CAPACITY;
    if (n == 0) return ICPB_OK;
    cudaStream_t s = ctx->stream;
    CK(ctx->dbg_idx.ensure(n));
    CK(ctx->dbg_kept.ensure(n));
    CK(ctx->kd0.ensure(n));
    debug_pairs_kernel<<<grid_for(n, 256), 256, 0, s>>>(ctx->gv, ctx->sx.p, ctx->sy.p, ctx->sz.p, ctx->nn_pos.p, ctx->nn_d.p, ctx->src_perm.p, n,
                                                        ctx->d_state, ctx->dbg_idx.p, (double *)ctx->kd0.p,
                                                        ctx->dbg_kept.p);
    CKL();
    if (model_idx) CK(cudaMemcpyAsync(model_idx, ctx->dbg_idx.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    if (dist) CK(cudaMemcpyAsync(dist, ctx->kd0.p, n * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (kept) CK(cudaMemcpyAsync(kept, ctx->dbg_kept.p, n * sizeof(uint8_t), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return ICPB_OK;
}

int icpb_trace(icpb_ctx *ctx, double *rows, size_t capacity_rows, size_t *n_rows)
{
    if (!ctx || !n_rows) return ICPB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const size_t n = (size_t)std::min(ctx->last_executed, ctx->trace_cap);
    *n_rows = n;
    if (!rows) return ICPB_OK;
    if (capacity_rows < n) return ICPB_ERR_CAPACITY;
    if (n) {
        CK(cudaMemcpyAsync(rows, ctx->trace.p, n * ICPB_TRACE_COLS_DEV * sizeof(double), cudaMemcpyDeviceToHost,
                           ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return ICPB_OK;
}

// ---------------------------------------------------------------------------
// kernel-level entry points
// ---------------------------------------------------------------------------
int icpb_find_pairs(icpb_ctx *ctx, const double *q, size_t n, double d_box, uint32_t *model_idx, double *dist)
{
    if (!ctx || (!q && n) || !model_idx || !dist) return ICPB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (ctx->grid_dirty) {
        int rc = build_grid(ctx);
        if (rc != ICPB_OK) return rc;
    }
    if (n == 0) return ICPB_OK;
    if (n >= 0xFFFFFFF0ull) return ICPB_ERR_ARG;
    cudaStream_t s = ctx->stream;
    ctx->gv.block_levels = (int)std::max(-1.0, std::min(ctx->opt_block_levels, 1.0));
    CK(ctx->stage.ensure(3 * n));
    CK(ctx->dbg_idx.ensure(n));
    CK(ctx->kd0.ensure(n));
    CK(cudaMemcpyAsync(ctx->stage.p, q, 3 * n * sizeof(double), cudaMemcpyHostToDevice, s));
    find_pairs_kernel<<<grid_for(n, SEARCH_THREADS), SEARCH_THREADS, 0, s>>>(ctx->gv, ctx->stage.p, (uint32_t)n, d_box,
                                                                             ctx->dbg_idx.p, (double *)ctx->kd0.p);
    CKL();
    ++ctx->launches;
    CK(cudaMemcpyAsync(model_idx, ctx->dbg_idx.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(dist, ctx->kd0.p, n * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    ctx->have_source = false; // stage / scratch were reused
    ctx->last_valid = false;
    ctx->pd_valid = false;
    return ICPB_OK;
}

__global__ void trim_setup_kernel(IcpState *st, unsigned long long n_po, unsigned long long found)
{
    st->done = 0;
    st->n_po = n_po;
    st->found = found;
    st->need_tie = 0;
    st->tie_local = 0;
    st->p_cut = 0x100000000ll;
    for (int i = 0; i < 16; ++i) st->tickets[i] = 0;
}

int icpb_trim(icpb_ctx *ctx, const double *dist, size_t n, size_t n_po, double *tau, size_t *kept, size_t *n_ties_kept)
{
    if (!ctx || (!dist && n) || !tau || !kept) return ICPB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (n >= 0xFFFFFFF0ull) return ICPB_ERR_ARG;
    cudaStream_t s = ctx->stream;
    CK(ctx->nn_d.ensure(std::max<size_t>(n, 1)));
    size_t found = 0;
    for (size_t i = 0; i < n; ++i) found += isfinite(dist[i]) ? 1 : 0;
    CK(ctx->src_perm.ensure(std::max<size_t>(n, 1)));
    if (n) CK(cudaMemcpyAsync(ctx->nn_d.p, dist, n * sizeof(double), cudaMemcpyHostToDevice, s));
    if (n) iota_kernel<<<grid_for(n, 256), 256, 0, s>>>(ctx->src_perm.p, (uint32_t)n);
    trim_setup_kernel<<<1, 1, 0, s>>>(ctx->d_state, n_po, found);
    CKL();
    const int saved_ranks = ctx->n_ranks;
    const P2PView saved_p2p = ctx->p2p;
    ctx->n_ranks = 1; // stand-alone trim is always rank-local
    ctx->p2p.world = 1;
    int rc = launch_trim(ctx, (uint32_t)n);
    ctx->n_ranks = saved_ranks;
    ctx->p2p = saved_p2p;
    if (rc != ICPB_OK) return rc;
    CK(cudaMemcpyAsync(ctx->h_state, ctx->d_state, sizeof(IcpState), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    *tau = ctx->h_state->sel_k > 0 ? ctx->h_state->tau : NAN;
    *kept = (size_t)ctx->h_state->sel_k;
    if (n_ties_kept) *n_ties_kept = (size_t)ctx->h_state->quota;
    ctx->have_source = false;
    ctx->last_valid = false;
    ctx->pd_valid = false;
    return ICPB_OK;
}

__global__ void pairs_setup_kernel(IcpState *st, unsigned long long n_po, unsigned long long found)
{
    double I[12];
    icpb_aff_identity(I);
    icpb_state_init(*st, I, 1, -1.0, -1.0, 1.0, n_po); // one loop body, identity source transform
    st->found = found;
    st->tie_local = 0;
}

int icpb_pairs_transform(icpb_ctx *ctx, const double *x, const double *p, const double *dist, size_t n, size_t n_po,
                         double T_out[16], double *mse, double *tau, size_t *kept)
{
    if (!ctx || ((!x || !p || !dist) && n) || !T_out || !mse || !tau || !kept) return ICPB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (n >= 0xFFFFFFF0ull) return ICPB_ERR_ARG;
    cudaStream_t s = ctx->stream;
    const size_t m = std::max<size_t>(n, 1);
    CK(ctx->stage.ensure(3 * m));
    CK(ctx->sx.ensure(m));
    CK(ctx->sy.ensure(m));
    CK(ctx->sz.ensure(m));
    CK(ctx->nn_pos.ensure(m));
    CK(ctx->nn_d.ensure(m));
    CK(ctx->src_perm.ensure(m));
    CK(ctx->pts2.ensure(m));
    const unsigned mom_blocks = std::max(1u, std::min(grid_for(n, MOM_THREADS), 148u * 6u));
    CK(ctx->partials.ensure((size_t)mom_blocks * ICPB_NSUMS));
    if (n) {
        CK(cudaMemcpyAsync(ctx->stage.p, p, 3 * n * sizeof(double), cudaMemcpyHostToDevice, s));
        aos_to_soa_kernel<<<grid_for(n, 256), 256, 0, s>>>(ctx->stage.p, n, ctx->sx.p, ctx->sy.p, ctx->sz.p);
        CK(cudaStreamSynchronize(s)); // stage is reused below
        CK(cudaMemcpyAsync(ctx->stage.p, x, 3 * n * sizeof(double), cudaMemcpyHostToDevice, s));
        pack_points_kernel<<<grid_for(n, 256), 256, 0, s>>>(ctx->stage.p, (uint32_t)n, ctx->pts2.p);
        iota_kernel<<<grid_for(n, 256), 256, 0, s>>>(ctx->nn_pos.p, (uint32_t)n);
        iota_kernel<<<grid_for(n, 256), 256, 0, s>>>(ctx->src_perm.p, (uint32_t)n);
        CK(cudaMemcpyAsync(ctx->nn_d.p, dist, n * sizeof(double), cudaMemcpyHostToDevice, s));
        CKL();
    }
    size_t found = 0;
    for (size_t i = 0; i < n; ++i) found += isfinite(dist[i]) ? 1 : 0;
    pairs_setup_kernel<<<1, 1, 0, s>>>(ctx->d_state, n_po, found);
    CKL();
    const int saved_ranks = ctx->n_ranks;
    const P2PView saved_p2p = ctx->p2p;
    ctx->n_ranks = 1;
    ctx->p2p.world = 1;
    int rc = launch_trim(ctx, (uint32_t)n);
    GridView gv;
    memset(&gv, 0, sizeof(gv));
    gv.pts = ctx->pts2.p;
    gv.n = (uint32_t)n;
    if (rc == ICPB_OK) {
        moments_kernel<<<mom_blocks, MOM_THREADS, 0, s>>>(gv, ctx->sx.p, ctx->sy.p, ctx->sz.p, (uint32_t)n, ctx->d_state,
                                                          ctx->nn_pos.p, ctx->nn_d.p, ctx->src_perm.p, ctx->partials.p,
                                                          nullptr, 0, 1, ctx->p2p, 0, 0);
        ctx->launches += 1;
    }
    ctx->n_ranks = saved_ranks;
    ctx->p2p = saved_p2p;
    if (rc != ICPB_OK) return rc;
    CKL();
    CK(cudaMemcpyAsync(ctx->h_state, ctx->d_state, sizeof(IcpState), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const IcpState &hs = *ctx->h_state;
    aff12_to_cm(hs.C, T_out);
    *mse = hs.mse;
    *tau = hs.pairs > 0 ? hs.tau : NAN;
    *kept = (size_t)hs.pairs;
    ctx->have_source = false;
    ctx->last_valid = false;
    ctx->pd_valid = false;
    return hs.early ? ICPB_ERR_NOT_ENOUGH_PAIRS : ICPB_OK;
}

// ---------------------------------------------------------------------------
// multi-GPU
// ---------------------------------------------------------------------------
int icpb_nccl_unique_id(uint8_t id_out[ICPB_NCCL_ID_BYTES])
{
    if (!id_out) return ICPB_ERR_ARG;
    if (!nccl_load()) return ICPB_ERR_NCCL;
    nccl_uid_t id;
    if (g_nccl.GetUniqueId(&id) != 0) return ICPB_ERR_NCCL;
    memcpy(id_out, &id, ICPB_NCCL_ID_BYTES);
    return ICPB_OK;
}

int icpb_p2p_export(icpb_ctx *ctx, uint8_t handle_out[ICPB_IPC_HANDLE_BYTES])
{
    if (!ctx || !handle_out) return ICPB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    static_assert(sizeof(cudaIpcMemHandle_t) == ICPB_IPC_HANDLE_BYTES, "IPC handle size");
    if (!ctx->p2p_local) {
        CK(cudaMalloc((void **)&ctx->p2p_local, P2P_BUFFER_WORDS * sizeof(uint32_t)));
        CK(cudaMemset(ctx->p2p_local, 0, P2P_BUFFER_WORDS * sizeof(uint32_t)));
        CK(cudaDeviceSynchronize());
    }
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, ctx->p2p_local));
    memcpy(handle_out, &h, ICPB_IPC_HANDLE_BYTES);
    return ICPB_OK;
}

int icpb_p2p_connect(icpb_ctx *ctx, int n_ranks, int rank, const uint8_t *handles)
{
    if (!ctx || n_ranks < 1 || n_ranks > P2P_MAX_RANKS || rank < 0 || rank >= n_ranks || !handles) return ICPB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (!ctx->p2p_local) return ICPB_ERR_ARG; // icpb_p2p_export first
    memset(&ctx->p2p, 0, sizeof(ctx->p2p));
    ctx->p2p.world = n_ranks;
    ctx->p2p.rank = rank;
    for (int r = 0; r < n_ranks; ++r) {
        if (r == rank) {
            ctx->p2p.peer[r] = ctx->p2p_local;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * ICPB_IPC_HANDLE_BYTES, ICPB_IPC_HANDLE_BYTES);
        void *p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->p2p_mapped[r] = p;
        ctx->p2p.peer[r] = (uint32_t *)p;
    }
    ctx->n_ranks = n_ranks;
    ctx->rank = rank;
    ctx->use_p2p = n_ranks > 1;
    if (n_ranks == 1) ctx->p2p.world = 1;
    return ICPB_OK;
}

int icpb_comm_init(icpb_ctx *ctx, int n_ranks, int rank, const uint8_t id_in[ICPB_NCCL_ID_BYTES])
{
    if (!ctx || n_ranks < 1 || n_ranks > 32 || rank < 0 || rank >= n_ranks || !id_in) return ICPB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (n_ranks == 1) {
        ctx->n_ranks = 1;
        ctx->rank = 0;
        return ICPB_OK;
    }
    if (!nccl_load()) {
        ctx->err = "libnccl.so.2 not found";
        return ICPB_ERR_NCCL;
    }
    nccl_uid_t id;
    memcpy(&id, id_in, ICPB_NCCL_ID_BYTES);
    NK(g_nccl.CommInitRank(&ctx->comm, n_ranks, id, rank));
    ctx->n_ranks = n_ranks;
    ctx->rank = rank;
    return ICPB_OK;
}

// ---------------------------------------------------------------------------
// Several GPUs in ONE process (icpb_create(ctx, device_ids, n_devices)): one context per device, the
// exchange buffers reached through cudaDeviceEnablePeerAccess instead of IPC handles, the same fused exchanges.
// The source is sharded into contiguous ranges of the adapter's visiting order, the model is replicated.
// ---------------------------------------------------------------------------
struct icpb_multi {
    std::vector<icpb_ctx *> ctx;
    std::vector<double> stage; // density-stepped source (host)
    std::string err;
};

static int multi_fail(icpb_multi *m, int i, int rc)
{
    m->err = "device " + std::to_string(m->ctx[i]->device) + ": " + icpb_strerror(rc) + " / " + m->ctx[i]->err;
    return rc;
}

int icpb_create_multi(icpb_multi **out, const int *device_ids, int n_devices)
{
    if (!out || !device_ids || n_devices < 1 || n_devices > P2P_MAX_RANKS) return ICPB_ERR_ARG;
    *out = nullptr;
    for (int i = 0; i < n_devices; ++i)
        for (int j = 0; j < i; ++j)
            if (device_ids[i] == device_ids[j]) return ICPB_ERR_ARG;
    icpb_multi *m = new (std::nothrow) icpb_multi;
    if (!m) return ICPB_ERR_CUDA;
    *out = m; // handed out even on failure, for icpb_multi_last_error; the caller destroys it
    for (int i = 0; i < n_devices; ++i) {
        icpb_ctx *c = nullptr;
        const int rc = icpb_create(&c, device_ids[i]);
        if (c) m->ctx.push_back(c);
        if (rc != ICPB_OK) {
            m->err = "device " + std::to_string(device_ids[i]) + ": " + icpb_strerror(rc) + (c ? " / " + c->err : std::string());
            return rc;
        }
    }
    if (n_devices == 1) return ICPB_OK;
    for (int i = 0; i < n_devices; ++i) {
        icpb_ctx *c = m->ctx[i];
        if (cudaSetDevice(c->device) != cudaSuccess) return ICPB_ERR_CUDA;
        for (int j = 0; j < n_devices; ++j) {
            if (j == i) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, c->device, device_ids[j]);
            if (!can) {
                m->err = "device " + std::to_string(c->device) + " cannot map the memory of device " + std::to_string(device_ids[j]);
                return ICPB_ERR_CUDA;
            }
            const cudaError_t e = cudaDeviceEnablePeerAccess(device_ids[j], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                m->err = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e);
                return ICPB_ERR_CUDA;
            }
            cudaGetLastError();
        }
        if (cudaMalloc((void **)&c->p2p_local, P2P_BUFFER_WORDS * sizeof(uint32_t)) != cudaSuccess ||
            cudaMemset(c->p2p_local, 0, P2P_BUFFER_WORDS * sizeof(uint32_t)) != cudaSuccess ||
            cudaDeviceSynchronize() != cudaSuccess) {
            m->err = "exchange buffer allocation failed";
            return ICPB_ERR_CUDA;
        }
    }
    for (int i = 0; i < n_devices; ++i) {
        icpb_ctx *c = m->ctx[i];
        memset(&c->p2p, 0, sizeof(c->p2p));
        c->p2p.world = n_devices;
        c->p2p.rank = i;
        for (int j = 0; j < n_devices; ++j) c->p2p.peer[j] = m->ctx[j]->p2p_local;
        c->n_ranks = n_devices;
        c->rank = i;
        c->use_p2p = true;
    }
    return ICPB_OK;
}

void icpb_destroy_multi(icpb_multi *m)
{
    if (!m) return;
    for (icpb_ctx *c : m->ctx) icpb_destroy(c);
    delete m;
}

const char *icpb_multi_last_error(const icpb_multi *m) { return m ? m->err.c_str() : ""; }
int icpb_multi_size(const icpb_multi *m) { return m ? (int)m->ctx.size() : 0; }
icpb_ctx *icpb_multi_context(icpb_multi *m, int i) { return (m && i >= 0 && i < (int)m->ctx.size()) ? m->ctx[i] : nullptr; }

int icpb_multi_set_option(icpb_multi *m, const char *key, double value)
{
    if (!m) return ICPB_ERR_ARG;
    for (size_t i = 0; i < m->ctx.size(); ++i) {
        const int rc = icpb_set_option(m->ctx[i], key, value);
        if (rc != ICPB_OK) return multi_fail(m, (int)i, rc);
    }
    return ICPB_OK;
}

int icpb_multi_model_add(icpb_multi *m, const double *xyz, size_t n, const double T[16], double density)
{
    if (!m) return ICPB_ERR_ARG;
    for (size_t i = 0; i < m->ctx.size(); ++i) {
        const int rc = icpb_model_add(m->ctx[i], xyz, n, T, density);
        if (rc != ICPB_OK) return multi_fail(m, (int)i, rc);
    }
    return ICPB_OK;
}

int icpb_multi_model_clear(icpb_multi *m)
{
    if (!m) return ICPB_ERR_ARG;
    for (size_t i = 0; i < m->ctx.size(); ++i) {
        const int rc = icpb_model_clear(m->ctx[i]);
        if (rc != ICPB_OK) return multi_fail(m, (int)i, rc);
    }
    return ICPB_OK;
}

int icpb_multi_source_upload(icpb_multi *m, const double *xyz, size_t n, const double T[16], double density)
{
    if (!m || (!xyz && n) || !T || !(density > 0.0)) return ICPB_ERR_ARG;
    size_t visited = 0;
    const double *src = density_gather(xyz, n, density, m->stage, &visited);
    const size_t adapter_size = (size_t)((double)n * density); // PointcloudAdapter::size() (icp/icp.hpp:143-146)
    const size_t world = m->ctx.size(), base = visited / world, rem = visited % world;
    for (size_t r = 0; r < world; ++r) { // contiguous ranges: ties at the trim threshold are kept in global pair order
        const size_t lo = r * base + std::min(r, rem), cnt = base + (r < rem ? 1 : 0);
        const int rc = world == 1 ? icpb_source_upload(m->ctx[r], xyz, n, T, density)
                                  : icpb_source_upload_shard(m->ctx[r], src + 3 * lo, cnt, adapter_size, T);
        if (rc != ICPB_OK) return multi_fail(m, (int)r, rc);
    }
    return ICPB_OK;
}

int icpb_multi_align_resident(icpb_multi *m, const icpb_params *prm, icpb_result *out)
{
    if (!m || !prm || !out) return ICPB_ERR_ARG;
    const size_t world = m->ctx.size();
    if (world == 1) {
        const int rc = icpb_align_resident(m->ctx[0], prm, out);
        return rc == ICPB_OK ? rc : multi_fail(m, 0, rc);
    }
    // every device gets its whole loop (one graph launch) before any is waited for: the ranks meet inside the kernels
    int first_rc = ICPB_OK;
    size_t launched = 0;
    for (; launched < world; ++launched) {
        const int rc = align_enqueue(m->ctx[launched], prm, true);
        if (rc != ICPB_OK) {
            first_rc = multi_fail(m, (int)launched, rc);
            break;
        }
    }
    icpb_result r0;
    memset(&r0, 0, sizeof(r0));
    for (size_t r = 0; r < launched; ++r) { // ranks already running give up on a missing peer after their spin budget
        icpb_result rr;
        const int rc = align_collect(m->ctx[r], &rr);
        if (rc != ICPB_OK && first_rc == ICPB_OK) first_rc = multi_fail(m, (int)r, rc);
        if (r == 0) r0 = rr;
    }
    if (first_rc != ICPB_OK) return first_rc;
    *out = r0; // every rank ends with the same pose, pair count and error
    return ICPB_OK;
}

int icpb_multi_align(icpb_multi *m, const double *xyz, size_t n, const double T[16], double density, const icpb_params *prm,
                     icpb_result *out)
{
    const int rc = icpb_multi_source_upload(m, xyz, n, T, density);
    if (rc != ICPB_OK) return rc;
    return icpb_multi_align_resident(m, prm, out);
}

int icpb_multi_pairs_distance(icpb_multi *m, double *out_sorted, size_t capacity, size_t *n_out)
{
    if (!m || !n_out) return ICPB_ERR_ARG;
    const size_t world = m->ctx.size();
    if (world == 1) {
        const int rc = icpb_pairs_distance(m->ctx[0], out_sorted, capacity, n_out);
        return rc == ICPB_OK ? rc : multi_fail(m, 0, rc);
    }
    size_t total = 0;
    {
        const int rc = icpb_pairs_distance(m->ctx[0], nullptr, 0, &total); // the count of all ranks' kept pairs
        if (rc != ICPB_OK) return multi_fail(m, 0, rc);
    }
    *n_out = total;
    if (!out_sorted) return ICPB_OK;
    if (capacity < total) return ICPB_ERR_CAPACITY;
    std::vector<double> part(total), merged;
    size_t have = 0;
    for (size_t r = 0; r < world; ++r) { // every rank: the kept distances of its shard, ascending; merged on the host
        size_t k = 0;
        const int rc = icpb_pairs_distance(m->ctx[r], part.data(), total, &k);
        if (rc != ICPB_OK) return multi_fail(m, (int)r, rc);
        merged.resize(have + k);
        std::merge(out_sorted, out_sorted + have, part.data(), part.data() + k, merged.begin());
        have += k;
        if (have > total) return ICPB_ERR_CUDA;
        std::copy(merged.begin(), merged.end(), out_sorted);
    }
    *n_out = have;
    return ICPB_OK;
}

} // extern "C"

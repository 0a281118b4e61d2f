// cpu_harness.cpp -- runs the product's __host__ __device__ search / pose code on
// the CPU so the non-GPU test tier can check it against the oracle.
//
// This is TEST code.  It is not a CPU fallback of the product: the grid is
// built here with std::sort instead of the device radix sort, and nothing in
// slam-envire_b200/ links against it.  What it exercises is the exact same
// icpb_nn_search() / icpb_iteration_finish() source that the kernels inline
// (slam-envire_b200/csrc/grid_view.h, pose.h).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>

struct HsCounters { unsigned long long cells, cands, pops, nodes, exact, expands; };
static HsCounters g_cnt;
#define ICPB_COUNT(what) (++g_cnt.what)
#include "../../slam-envire_b200/csrc/grid_view.h"
#include "../../slam-envire_b200/csrc/pose.h"
#include "../../slam-envire_b200/csrc/density.h"

static int g_block_levels = 1;
namespace {
struct HostGrid {
    GridView g;
    std::vector<MPoint> pts;
    std::vector<FPoint> ptsf;
    std::vector<uint32_t> cell_start, cell_box;
    std::vector<std::vector<uint32_t>> masks;
    std::vector<uint32_t> pos_of_idx;
};

void build(HostGrid &G, const double *model, size_t n, double h)
{
    GridView &g = G.g;
    memset(&g, 0, sizeof(g));
    double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (size_t i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) {
            mn[a] = std::fmin(mn[a], model[3 * i + a]);
            mx[a] = std::fmax(mx[a], model[3 * i + a]);
        }
    for (int a = 0; a < 3; ++a) {
        g.o[a] = n ? mn[a] : 0.0;
        g.dim[0][a] = n ? (int)(std::floor((mx[a] - mn[a]) / h) + 1.0) : 1;
    }
    g.h = h;
    g.inv_h = 1.0 / h;
    g.eps = 1e-9 * h;
    const size_t ncell = (size_t)g.dim[0][0] * g.dim[0][1] * g.dim[0][2];
    std::vector<uint32_t> key(n), perm(n);
    for (size_t i = 0; i < n; ++i) {
        const int cx = icpb_cell_coord(model[3 * i], g.o[0], g.inv_h, g.dim[0][0]);
        const int cy = icpb_cell_coord(model[3 * i + 1], g.o[1], g.inv_h, g.dim[0][1]);
        const int cz = icpb_cell_coord(model[3 * i + 2], g.o[2], g.inv_h, g.dim[0][2]);
        key[i] = (uint32_t)(((size_t)cz * g.dim[0][1] + cy) * g.dim[0][0] + cx);
    }
    std::iota(perm.begin(), perm.end(), 0u);
    std::stable_sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
    G.pts.resize(n + ICPB_PAD);
    G.ptsf.resize(n + ICPB_PAD);
    for (size_t j = n; j < n + ICPB_PAD; ++j) { // the sentinels the device build appends (pad_tail_kernel)
        G.pts[j].x = G.pts[j].y = G.pts[j].z = 1.0e300;
        G.pts[j].idx = ~0ull;
        G.ptsf[j].x = G.ptsf[j].y = G.ptsf[j].z = 3.0e18f;
        G.ptsf[j].w = 0.0f;
    }
    G.pos_of_idx.resize(n);
    G.cell_start.assign(ncell + 1, 0);
    for (size_t j = 0; j < n; ++j) {
        const uint32_t s = perm[j];
        G.pts[j].x = model[3 * s];
        G.pts[j].y = model[3 * s + 1];
        G.pts[j].z = model[3 * s + 2];
        G.pts[j].idx = s;
        G.ptsf[j].x = (float)((model[3 * s] - g.o[0]) * g.inv_h);
        G.ptsf[j].y = (float)((model[3 * s + 1] - g.o[1]) * g.inv_h);
        G.ptsf[j].z = (float)((model[3 * s + 2] - g.o[2]) * g.inv_h);
        G.ptsf[j].w = 0.0f;
        G.pos_of_idx[s] = (uint32_t)j;
        G.cell_start[key[s] + 1]++;
    }
    for (size_t c = 0; c < ncell; ++c) G.cell_start[c + 1] += G.cell_start[c];
    G.cell_box.assign(ncell, 0u);
    for (int cz = 0; cz < g.dim[0][2]; ++cz)
        for (int cy = 0; cy < g.dim[0][1]; ++cy)
            for (int cx = 0; cx < g.dim[0][0]; ++cx) {
                const size_t c = ((size_t)cz * g.dim[0][1] + cy) * g.dim[0][0] + cx;
                if (G.cell_start[c + 1] == G.cell_start[c]) continue;
                int lo[3] = {15, 15, 15}, hi[3] = {0, 0, 0};
                const double cl[3] = {(double)cx, (double)cy, (double)cz};
                for (uint32_t p = G.cell_start[c]; p < G.cell_start[c + 1]; ++p) {
                    const double u[3] = {(G.pts[p].x - g.o[0]) * g.inv_h, (G.pts[p].y - g.o[1]) * g.inv_h, (G.pts[p].z - g.o[2]) * g.inv_h};
                    for (int a = 0; a < 3; ++a) {
                        const int q = icpb_node_nibble(u[a], cl[a], 0);
                        lo[a] = std::min(lo[a], q);
                        hi[a] = std::max(hi[a], q);
                    }
                }
                G.cell_box[c] = icpb_node_word(0u, lo, hi);
            }
    int L = 0;
    G.masks.assign(ICPB_MAX_LEVELS, {});
    while (g.dim[L][0] > 1 || g.dim[L][1] > 1 || g.dim[L][2] > 1) {
        for (int a = 0; a < 3; ++a) g.dim[L + 1][a] = (g.dim[L][a] + 1) >> 1;
        ++L;
        const int *d = g.dim[L], *b = g.dim[L - 1];
        G.masks[L].assign((size_t)d[0] * d[1] * d[2], 0);
        for (int cz = 0; cz < d[2]; ++cz)
            for (int cy = 0; cy < d[1]; ++cy)
                for (int cx = 0; cx < d[0]; ++cx) {
                    unsigned m = 0;
                    int lo[3] = {15, 15, 15}, hi[3] = {0, 0, 0};
                    for (int ch = 0; ch < 8; ++ch) {
                        const int x = 2 * cx + (ch & 1), y = 2 * cy + ((ch >> 1) & 1), z = 2 * cz + ((ch >> 2) & 1);
                        if (x >= b[0] || y >= b[1] || z >= b[2]) continue;
                        const size_t c = ((size_t)z * b[1] + y) * b[0] + x;
                        if (L == 1) {
                            if (G.cell_start[c + 1] == G.cell_start[c]) continue;
                            m |= 1u << ch;
                            const double cl[3] = {2.0 * cx, 2.0 * cy, 2.0 * cz};
                            for (uint32_t p = G.cell_start[c]; p < G.cell_start[c + 1]; ++p) {
                                const double u[3] = {(G.pts[p].x - g.o[0]) * g.inv_h, (G.pts[p].y - g.o[1]) * g.inv_h,
                                                     (G.pts[p].z - g.o[2]) * g.inv_h};
                                for (int a = 0; a < 3; ++a) {
                                    const int q = icpb_node_nibble(u[a], cl[a], 1);
                                    lo[a] = std::min(lo[a], q);
                                    hi[a] = std::max(hi[a], q);
                                }
                            }
                        } else {
                            const uint32_t w = G.masks[L - 1][c];
                            if (!(w & 255u)) continue;
                            m |= 1u << ch;
                            for (int a = 0; a < 3; ++a) {
                                lo[a] = std::min(lo[a], icpb_nibble_up(w, 8 + 4 * a, (ch >> a) & 1));
                                hi[a] = std::max(hi[a], icpb_nibble_up(w, 20 + 4 * a, (ch >> a) & 1));
                            }
                        }
                    }
                    G.masks[L][((size_t)cz * d[1] + cy) * d[0] + cx] = m ? icpb_node_word(m, lo, hi) : 0u;
                }
    }
    g.L = L;
    for (int l = 1; l <= L; ++l) g.node[l] = G.masks[l].data();
    g.cell_start = G.cell_start.data();
    g.cell_box = G.cell_box.data();
    g.pts = G.pts.data();
    g.ptsf = G.ptsf.data();
    g.n = (uint32_t)n;
    g.block_levels = g_block_levels;
}
} // namespace

extern "C" {
// density stepping in closed form (csrc/density.h): the vertices PointcloudAdapter visits, and how many segments describe them
long long hs_density_indices(size_t n, double density, unsigned long long *out, size_t capacity, int *n_segs)
{
    DensitySeg segs[ICPB_DENSITY_MAX_SEGS];
    unsigned long long visited = 0;
    const int ns = icpb_density_segments(n, density, segs, ICPB_DENSITY_MAX_SEGS, &visited);
    *n_segs = ns;
    if (ns < 0) return -1;
    if (visited > capacity) return -2;
    for (unsigned long long k = 0; k < visited; ++k) out[k] = icpb_density_index(segs[icpb_density_find(segs, ns, k)], k);
    return (long long)visited;
}
void hs_set_block_levels(int v) { g_block_levels = v; }
void hs_counters(unsigned long long *out, int reset)
{
    out[0] = g_cnt.cells; out[1] = g_cnt.cands; out[2] = g_cnt.pops; out[3] = g_cnt.nodes; out[4] = g_cnt.exact; out[5] = g_cnt.expands;
    if (reset) memset(&g_cnt, 0, sizeof(g_cnt));
}

// exact NN of every query through icpb_nn_search; warm_idx (insertion indices, may be NULL) seeds the bound
int hs_search(const double *model, size_t nm, double h, const double *q, size_t nq, double d_box,
              const uint32_t *warm_idx, uint32_t *out_idx, double *out_d, int *out_level)
{
    HostGrid G;
    build(G, model, nm, h);
    for (size_t i = 0; i < nq; ++i) {
        uint32_t warm = ICPB_NONE;
        if (warm_idx && warm_idx[i] != ICPB_NONE && warm_idx[i] < nm) warm = G.pos_of_idx[warm_idx[i]];
        uint32_t pos;
        double d2;
        int lvl;
        icpb_nn_search(G.g, q[3 * i], q[3 * i + 1], q[3 * i + 2], d_box * d_box * (1.0 + 1e-12), warm, pos, d2, lvl);
        const double d = std::sqrt(d2);
        const bool ok = pos != ICPB_NONE && d <= d_box;
        out_idx[i] = ok ? (uint32_t)G.pts[pos].idx : ICPB_NONE;
        out_d[i] = ok ? d : d_box;
        if (out_level) out_level[i] = lvl;
    }
    return 0;
}

void hs_transform_from_sums(const double *sums, double *T12, double *mse) { icpb_transform_from_sums(sums, T12, mse); }

// Sequential emulation of the device loop (search -> trim -> moments -> pose step) using the
// product's state machine.  result: C (col-major 16), then iter, pairs, mse, mse_diff, d_box, early.
// trace: rows of 24 doubles.  Returns the number of executed loop bodies.
int hs_align(const double *model, size_t nm, double h, const double *src, size_t ns, size_t adapter_size,
             const double *T_l2g_cm, size_t max_iter, double min_mse, double min_mse_diff, double overlap,
             int use_warm, double *result, double *trace, size_t trace_cap, int skip, double reach_cells, double floor_cells,
             double lazy_margin)
{
    HostGrid G;
    build(G, model, nm, h);
    IcpState st;
    memset(&st, 0, sizeof(st));
    double Cl2g[12];
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) Cl2g[r * 4 + c] = T_l2g_cm[c * 4 + r];
        Cl2g[r * 4 + 3] = T_l2g_cm[12 + r];
    }
    const size_t n_po = (size_t)((double)adapter_size * overlap);
    icpb_state_init(st, Cl2g, max_iter, min_mse, min_mse_diff, overlap, n_po, (skip && use_warm) ? lazy_margin : -1.0);
    std::vector<uint32_t> nn_pos(ns, ICPB_NONE);
    std::vector<double> nn_d(ns, 0.0);
    std::vector<float> nn_lb(ns, 0.0f);
    NNTune tune;
    tune.reach_cells = (float)reach_cells;
    tune.floor_cells = (float)floor_cells;
    while (!st.done) {
        // the set-up / search phases of search_kernel (kernels.cuh), one query after the other
        const bool warm_on = use_warm && st.warm;
        NNTune tn = tune;
        tn.skip = (skip && warm_on) ? 1 : 0;
        float dM[12];
        for (int k = 0; k < 12; ++k) dM[k] = (float)(st.M[k] - st.Mprev[k]);
        const double dbox2s = st.d_box * st.d_box * (1.0 + 1e-12);
        for (size_t i = 0; i < ns; ++i) {
            double qx, qy, qz;
            icpb_aff_apply(st.M, src[3 * i], src[3 * i + 1], src[3 * i + 2], qx, qy, qz);
            const uint32_t wpos = warm_on ? nn_pos[i] : ICPB_NONE;
            const float lbp = tn.skip ? nn_lb[i] : 0.0f;
            const float delta = tn.skip ? icpb_motion_bound(dM, (float)src[3 * i], (float)src[3 * i + 1], (float)src[3 * i + 2]) : 0.0f;
            double d2w;
            float lb;
            uint32_t pos;
            double d2;
            int lvl = 0;
            const int sk = icpb_nn_skip(G.g, qx, qy, qz, st.d_box, wpos, lbp, delta, d2w, lb);
            double standin;
            if (sk) {
                pos = wpos;
                d2 = d2w;
                st.skipped += 1;
            } else if (tn.skip && icpb_nn_defer(d2w, lb, st.d_box, st.tau_guard, standin)) {
                nn_d[i] = standin;
                nn_lb[i] = lb;
                st.found += 1;
                st.deferred += 1;
                continue;
            } else {
                const double need2 = dbox2s < d2w ? dbox2s : d2w;
                const float rs = icpb_cover_radius(G.g, need2, delta, tn);
                icpb_nn_search(G.g, qx, qy, qz, dbox2s, wpos, pos, d2, lvl, (double)rs * (double)rs, &lb);
            }
            const double d = std::sqrt(d2);
            const bool ok = pos != ICPB_NONE && d <= st.d_box;
            nn_pos[i] = pos;
            nn_d[i] = ok ? d : INFINITY;
            nn_lb[i] = lb;
            st.found += ok ? 1 : 0;
            st.far += lvl > 0 ? 1 : 0;
        }
        // trim: k-th smallest by (distance, position)
        const size_t k = std::min<size_t>(n_po, (size_t)st.found);
        st.sel_k = k;
        st.p_cut = 0x100000000ll;
        if (k == 0) {
            st.tau = NAN;
        } else {
            std::vector<uint32_t> ord(ns);
            std::iota(ord.begin(), ord.end(), 0u);
            std::nth_element(ord.begin(), ord.begin() + (k - 1), ord.end(), [&](uint32_t a, uint32_t b) {
                return nn_d[a] < nn_d[b] || (nn_d[a] == nn_d[b] && a < b);
            });
            st.tau = nn_d[ord[k - 1]];
            st.p_cut = ord[k - 1];
        }
        for (int j = 0; j < ICPB_NSUMS; ++j) st.sums[j] = 0.0;
        for (size_t i = 0; i < ns && k > 0; ++i) {
            const double d = nn_d[i];
            if (!(d < st.tau || (d == st.tau && (long long)i <= st.p_cut))) continue;
            double px, py, pz;
            icpb_aff_apply(st.M, src[3 * i], src[3 * i + 1], src[3 * i + 2], px, py, pz);
            const MPoint &m = G.pts[nn_pos[i]];
            const double p[3] = {px, py, pz}, x[3] = {m.x, m.y, m.z};
            for (int r = 0; r < 3; ++r) {
                st.sums[r] += p[r];
                st.sums[3 + r] += x[r];
                for (int c = 0; c < 3; ++c) st.sums[6 + r * 3 + c] += p[r] * x[c];
            }
            st.sums[15] += d * d;
            st.sums[16] += 1.0;
        }
        const int row = st.executed;
        icpb_iteration_finish(st, (size_t)row < trace_cap ? trace + (size_t)row * ICPB_TRACE_COLS_DEV : nullptr);
    }
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) result[c * 4 + r] = st.C[r * 4 + c];
        result[12 + r] = st.C[r * 4 + 3];
        result[r * 4 + 3] = 0.0;
    }
    result[15] = 1.0;
    result[16] = (double)st.iter;
    result[17] = (double)st.pairs;
    result[18] = st.mse;
    result[19] = st.mse_diff;
    result[20] = st.d_box;
    result[21] = (double)st.early;
    return st.executed;
}

// The skip protocol on caller-supplied query positions: `steps` sets of nq queries (the same queries moved a little
// from set to set).  Every set is answered with icpb_nn_skip / icpb_nn_search exactly like search_kernel does,
// delta = the true displacement rounded up.  out_idx/out_d: answers of every set; out_skipped: 1 where no search ran.
int hs_skip_sequence(const double *model, size_t nm, double h, const double *q, size_t nq, int steps, const double *d_box,
                     double reach_cells, double floor_cells, uint32_t *out_idx, double *out_d, uint8_t *out_skipped)
{
    HostGrid G;
    build(G, model, nm, h);
    std::vector<uint32_t> nn_pos(nq, ICPB_NONE);
    std::vector<float> nn_lb(nq, 0.0f);
    NNTune tn;
    tn.reach_cells = (float)reach_cells;
    tn.floor_cells = (float)floor_cells;
    for (int t = 0; t < steps; ++t) {
        tn.skip = t > 0;
        const double db = d_box[t], dbox2s = db * db * (1.0 + 1e-12);
        for (size_t i = 0; i < nq; ++i) {
            const double *c = q + ((size_t)t * nq + i) * 3;
            float delta = 0.0f;
            if (t > 0) {
                const double *b = c - nq * 3;
                const double m = std::sqrt((c[0] - b[0]) * (c[0] - b[0]) + (c[1] - b[1]) * (c[1] - b[1]) + (c[2] - b[2]) * (c[2] - b[2]));
                delta = (float)m * 1.00001f + 1e-30f;
            }
            double d2w, d2;
            float lb;
            uint32_t pos;
            int lvl;
            const int sk = icpb_nn_skip(G.g, c[0], c[1], c[2], db, t > 0 ? nn_pos[i] : ICPB_NONE, nn_lb[i], delta, d2w, lb);
            if (sk) {
                pos = nn_pos[i];
                d2 = d2w;
            } else {
                const double need2 = dbox2s < d2w ? dbox2s : d2w;
                const float rs = icpb_cover_radius(G.g, need2, delta, tn);
                icpb_nn_search(G.g, c[0], c[1], c[2], dbox2s, t > 0 ? nn_pos[i] : ICPB_NONE, pos, d2, lvl, (double)rs * (double)rs, &lb);
            }
            nn_pos[i] = pos;
            nn_lb[i] = lb;
            const double d = std::sqrt(d2);
            const bool ok = pos != ICPB_NONE && d <= db;
            out_idx[(size_t)t * nq + i] = ok ? (uint32_t)G.pts[pos].idx : ICPB_NONE;
            out_d[(size_t)t * nq + i] = ok ? d : db;
            out_skipped[(size_t)t * nq + i] = sk ? 1 : 0;
        }
    }
    return 0;
}
}

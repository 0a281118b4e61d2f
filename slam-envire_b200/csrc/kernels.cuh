// kernels.cuh -- the sm_100a kernels of the trimmed-ICP hot path.
//
//   K0  transform_append_kernel   model vertices -> global frame (addModel, icp/icp.hpp:231-240)
//   K1  keys / gather / cell counts / node words (occupancy + tight boxes) / cell boxes: uniform-grid index build
//   K2+K3 search_kernel           source transform fused with the exact NN search (findPairs, icp/icp.hpp:242-252)
//   K4  select_kernel             radix-select of the n_po-th smallest distance (Pairs::trim, icp/icp.cpp:23-41)
//   K5  moments_kernel            17 fp64 sums over the kept pairs (Pairs::getTransform, icp/icp.cpp:55-73)
//   K6  pose step in the last block of K5 (icp/icp.cpp:76-99, icp/icp.hpp:479-485)
//   K7  kept-distance compaction (+ radix sort) for getPairsDistance (icp/icp.hpp:500-504)
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include "grid_view.h"
#include "pose.h"
#include "density.h"

namespace icpb {

// model normal + edge flag in the sorted order of pts[] (FindPairsKDEAN nodes, icp/icp.hpp:91-97)
struct __align__(32) NPoint {
    double nx, ny, nz;
    unsigned long long edge;
};
// the edge/normal pair filter inputs of one align (all null/0 for the plain TrimmedKD path)
struct EanView {
    const NPoint *mn;                // model, sorted like GridView::pts
    const double *snx, *sny, *snz;   // source normals (local frame), in the order of sx/sy/sz
    const uint8_t *sedge;
    int on;
};
// ---------------------------------------------------------------------------
// K0: p = T * v  appended to the model (fp64 AoS in, AoS out)
// ---------------------------------------------------------------------------
struct Aff12 {
    double a[12];
};

__global__ void transform_append_kernel(const double *__restrict__ in, size_t n, Aff12 T, double *__restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double px, py, pz;
    icpb_aff_apply(T.a, in[3 * i], in[3 * i + 1], in[3 * i + 2], px, py, pz);
    out[3 * i] = px;
    out[3 * i + 1] = py;
    out[3 * i + 2] = pz;
}

// n = L * v  (normals ride with the vertices: PointcloudEdgeAndNormalAdapter::next, icp/icp.hpp:184-193)
__global__ void rotate_append_kernel(const double *__restrict__ in, size_t n, Aff12 T, double *__restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double vx = in[3 * i], vy = in[3 * i + 1], vz = in[3 * i + 2];
    out[3 * i] = (T.a[0] * vx + T.a[1] * vy) + T.a[2] * vz;
    out[3 * i + 1] = (T.a[4] * vx + T.a[5] * vy) + T.a[6] * vz;
    out[3 * i + 2] = (T.a[8] * vx + T.a[9] * vy) + T.a[10] * vz;
}
// SCAN_EDGE bit of the vertex attributes: attrs[idx] & (1 << envire::Pointcloud::SCAN_EDGE), SCAN_EDGE = 0x01
__global__ void edge_flags_kernel(const int *__restrict__ attrs, size_t n, uint8_t *__restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (attrs[i] & (1 << 0x01)) ? 1 : 0;
}

__global__ void gather_ean_kernel(const double *__restrict__ nrm, const uint8_t *__restrict__ edge,
                                  const uint32_t *__restrict__ perm, uint32_t n, NPoint *__restrict__ out)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const size_t s = perm[j];
    NPoint p;
    p.nx = nrm[3 * s];
    p.ny = nrm[3 * s + 1];
    p.nz = nrm[3 * s + 2];
    p.edge = edge[s];
    out[j] = p;
}
__global__ void gather_u8_kernel(const uint8_t *__restrict__ in, const uint32_t *__restrict__ perm, uint32_t n,
                                 uint8_t *__restrict__ out)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) out[j] = in[perm[j]];
}

// AoS double[n][3] -> SoA x[], y[], z[] (coalesced loads for the search kernel)
__global__ void aos_to_soa_kernel(const double *__restrict__ in, size_t n, double *__restrict__ x,
                                  double *__restrict__ y, double *__restrict__ z)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    x[i] = in[3 * i];
    y[i] = in[3 * i + 1];
    z[i] = in[3 * i + 2];
}

// ---------------------------------------------------------------------------
// K1: grid build
// ---------------------------------------------------------------------------
// per-block min/max of the model coordinates: out[block][6] = minx,miny,minz,maxx,maxy,maxz
__global__ void __launch_bounds__(256) bbox_kernel(const double *__restrict__ pts, size_t n, double *__restrict__ out)
{
    double mn[3] = {icpb_inf(), icpb_inf(), icpb_inf()}, mx[3] = {-icpb_inf(), -icpb_inf(), -icpb_inf()};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        for (int a = 0; a < 3; ++a) {
            const double v = pts[3 * i + a];
            mn[a] = fmin(mn[a], v);
            mx[a] = fmax(mx[a], v);
        }
    __shared__ double sm[8][6];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int a = 0; a < 3; ++a)
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = fmin(mn[a], __shfl_xor_sync(0xFFFFFFFFu, mn[a], o));
            mx[a] = fmax(mx[a], __shfl_xor_sync(0xFFFFFFFFu, mx[a], o));
        }
    if (lane == 0)
        for (int a = 0; a < 3; ++a) {
            sm[warp][a] = mn[a];
            sm[warp][3 + a] = mx[a];
        }
    __syncthreads();
    if (threadIdx.x < 6) {
        double v = sm[0][threadIdx.x];
        for (int w = 1; w < 8; ++w) v = threadIdx.x < 3 ? fmin(v, sm[w][threadIdx.x]) : fmax(v, sm[w][threadIdx.x]);
        out[(size_t)blockIdx.x * 6 + threadIdx.x] = v;
    }
}

struct GridDims {
    double o[3], h, inv_h;
    int nx, ny, nz;
};

__global__ void cell_keys_kernel(const double *__restrict__ pts, uint32_t n, GridDims g, uint32_t *__restrict__ keys,
                                 uint32_t *__restrict__ vals, uint32_t base = 0)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int cx = icpb_cell_coord(pts[3 * (size_t)i], g.o[0], g.inv_h, g.nx);
    const int cy = icpb_cell_coord(pts[3 * (size_t)i + 1], g.o[1], g.inv_h, g.ny);
    const int cz = icpb_cell_coord(pts[3 * (size_t)i + 2], g.o[2], g.inv_h, g.nz);
    keys[i] = (uint32_t)(((size_t)cz * g.ny + cy) * g.nx + cx);
    vals[i] = base + i; // insertion index
}

// occupied level-0 cells for a candidate cell edge (the build tries a few): one bit per cell, no sort
__global__ void __launch_bounds__(256) cell_occupancy_kernel(const double *__restrict__ pts, uint32_t n, GridDims g,
                                                              uint32_t *__restrict__ bitmap, unsigned long long *__restrict__ counter)
{
    unsigned c = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int cx = icpb_cell_coord(pts[3 * (size_t)i], g.o[0], g.inv_h, g.nx);
        const int cy = icpb_cell_coord(pts[3 * (size_t)i + 1], g.o[1], g.inv_h, g.ny);
        const int cz = icpb_cell_coord(pts[3 * (size_t)i + 2], g.o[2], g.inv_h, g.nz);
        const size_t key = ((size_t)cz * g.ny + cy) * g.nx + cx;
        const uint32_t bit = 1u << (key & 31u);
        if (!(atomicOr(bitmap + (key >> 5), bit) & bit)) ++c; // this thread set the bit: a cell not seen before
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(counter, (unsigned long long)c);
}

// ICPB_PAD sentinel entries behind the sorted points (grid_view.h): never the nearest neighbour of anything
__global__ void pad_tail_kernel(MPoint *__restrict__ pts, FPoint *__restrict__ ptsf, uint32_t n)
{
    if (threadIdx.x >= ICPB_PAD) return;
    MPoint m;
    m.x = m.y = m.z = 1.0e300;
    m.idx = ~0ull;
    pts[n + threadIdx.x] = m;
    FPoint f;
    f.x = f.y = f.z = 3.0e18f; // squared distances overflow to +inf
    f.w = 0.0f;
    ptsf[n + threadIdx.x] = f;
}

// Tight boxes without a pass per cell: every level-0 cell owns 48 bits (cell_bits[]), one per sixteenth of the cell edge
// and axis (bits 0-15 x, 16-31 y, 32-47 z); a point ORs in the three sixteenths it lies in.  The box of the cell is
// lowest / highest set bit per axis -- and a cloud appended later only ORs its own points in (the growing model).
__device__ __forceinline__ unsigned long long cell_point_bits(double px, double py, double pz, const GridDims &g, uint32_t key)
{
    const int cx = (int)(key % (uint32_t)g.nx), cy = (int)((key / (uint32_t)g.nx) % (uint32_t)g.ny),
              cz = (int)(key / ((uint32_t)g.nx * (uint32_t)g.ny));
    const int bx = icpb_node_nibble((px - g.o[0]) * g.inv_h, (double)cx, 0);
    const int by = icpb_node_nibble((py - g.o[1]) * g.inv_h, (double)cy, 0);
    const int bz = icpb_node_nibble((pz - g.o[2]) * g.inv_h, (double)cz, 0);
    return (1ull << bx) | (1ull << (16 + by)) | (1ull << (32 + bz));
}
// node-word form of a cell's bits: bits 8-31 the six nibbles, low byte 1 = occupied (what node_levelN_kernel looks at)
__device__ __forceinline__ uint32_t cell_word_from_bits(unsigned long long b)
{
    if (!b) return 0u;
    const unsigned x = (unsigned)b & 0xFFFFu, y = (unsigned)(b >> 16) & 0xFFFFu, z = (unsigned)(b >> 32) & 0xFFFFu;
    const int lo[3] = {__ffs((int)x) - 1, __ffs((int)y) - 1, __ffs((int)z) - 1};
    const int hi[3] = {31 - __clz((int)x), 31 - __clz((int)y), 31 - __clz((int)z)};
    return icpb_node_word(1u, lo, hi);
}
// all 32 lanes call; keys ascending within the warp (sorted points): one atomic per cell and warp
__device__ __forceinline__ void cell_bits_or_sorted(unsigned long long *__restrict__ cell_bits, uint32_t key, unsigned long long v, bool valid)
{
    const int lane = threadIdx.x & 31;
    if (!valid) key = 0xFFFFFFFFu, v = 0ull;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long ov = __shfl_up_sync(0xFFFFFFFFu, v, d);
        const uint32_t ok = __shfl_up_sync(0xFFFFFFFFu, key, d);
        if (lane >= d && ok == key) v |= ov;
    }
    const uint32_t nk = __shfl_down_sync(0xFFFFFFFFu, key, 1);
    if (valid && (lane == 31 || nk != key)) atomicOr(cell_bits + key, v);
}

// sorted layout: pts_sorted[j] = {model[perm[j]], perm[j]}; cell_count[key[j]]++; cell_bits[key[j]] |= bits of the point
__global__ void gather_points_kernel(const double *__restrict__ raw, const uint32_t *__restrict__ keys,
                                     const uint32_t *__restrict__ perm, uint32_t n, GridDims g,
                                     MPoint *__restrict__ out, FPoint *__restrict__ outf,
                                     uint32_t *__restrict__ cell_count, unsigned long long *__restrict__ cell_bits)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = j < n;
    uint32_t key = 0;
    unsigned long long bits = 0;
    if (valid) {
        const uint32_t src = perm[j];
        key = keys[j];
        MPoint m;
        m.x = raw[3 * (size_t)src];
        m.y = raw[3 * (size_t)src + 1];
        m.z = raw[3 * (size_t)src + 2];
        m.idx = src;
        out[j] = m;
        FPoint f;
        f.x = (float)((m.x - g.o[0]) * g.inv_h);
        f.y = (float)((m.y - g.o[1]) * g.inv_h);
        f.z = (float)((m.z - g.o[2]) * g.inv_h);
        f.w = 0.0f;
        outf[j] = f;
        atomicAdd(&cell_count[key], 1u);
        bits = cell_point_bits(m.x, m.y, m.z, g, key);
    }
    cell_bits_or_sorted(cell_bits, key, bits, valid);
}
// cell_box[] (the walk's and the block scan's tight boxes of the level-0 cells) from the bits, all cells
__global__ void cell_words_kernel(const unsigned long long *__restrict__ cell_bits, size_t ncell, uint32_t *__restrict__ cell_box)
{
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < ncell) cell_box[c] = cell_word_from_bits(cell_bits[c]);
}
// ... and of the cells a batch of new points fell into (keys[]: their cells, any order, duplicates welcome)
__global__ void cell_words_update_kernel(const unsigned long long *__restrict__ cell_bits, const uint32_t *__restrict__ keys,
                                         uint32_t k, uint32_t *__restrict__ cell_box)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < k) cell_box[keys[i]] = cell_word_from_bits(cell_bits[keys[i]]);
}

// ---------------------------------------------------------------------------
// K1b: append a batch to an existing index without re-sorting the model (growing model).
// The new batch is sorted by cell key on its own; old and new points are then merged in one streaming pass:
// an old point of cell c moves up by the number of new points in cells < c, the new points of cell c follow it.
// ---------------------------------------------------------------------------
__global__ void key_hist_kernel(const uint32_t *__restrict__ keys, uint32_t n, uint32_t *__restrict__ cnt)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&cnt[keys[i]], 1u);
}
__global__ void merge_move_old_kernel(const MPoint *__restrict__ pts, const FPoint *__restrict__ ptsf,
                                      const uint32_t *__restrict__ pkey, const NPoint *__restrict__ npts, uint32_t n,
                                      const uint32_t *__restrict__ newpref, MPoint *__restrict__ pts2,
                                      FPoint *__restrict__ ptsf2, uint32_t *__restrict__ pkey2, NPoint *__restrict__ npts2)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t c = pkey[p];
    const uint32_t q = p + newpref[c];
    pts2[q] = pts[p];
    ptsf2[q] = ptsf[p];
    pkey2[q] = c;
    if (npts) npts2[q] = npts[p];
}
__global__ void merge_place_new_kernel(const double *__restrict__ raw, const double *__restrict__ raw_nrm,
                                       const uint8_t *__restrict__ raw_edge, const uint32_t *__restrict__ keys,
                                       const uint32_t *__restrict__ vals, uint32_t k, GridDims g,
                                       const uint32_t *__restrict__ cell_start_old, MPoint *__restrict__ pts2,
                                       FPoint *__restrict__ ptsf2, uint32_t *__restrict__ pkey2, NPoint *__restrict__ npts2,
                                       unsigned long long *__restrict__ cell_bits)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= k) return;
    const uint32_t c = keys[j], src = vals[j];
    const uint32_t q = cell_start_old[c + 1] + j; // old points of cells <= c, plus the new ones sorted before j
    MPoint m;
    m.x = raw[3 * (size_t)src];
    m.y = raw[3 * (size_t)src + 1];
    m.z = raw[3 * (size_t)src + 2];
    m.idx = src;
    pts2[q] = m;
    FPoint f;
    f.x = (float)((m.x - g.o[0]) * g.inv_h);
    f.y = (float)((m.y - g.o[1]) * g.inv_h);
    f.z = (float)((m.z - g.o[2]) * g.inv_h);
    f.w = 0.0f;
    ptsf2[q] = f;
    pkey2[q] = c;
    atomicOr(cell_bits + c, cell_point_bits(m.x, m.y, m.z, g, c));
    if (npts2) {
        NPoint np;
        np.nx = raw_nrm[3 * (size_t)src];
        np.ny = raw_nrm[3 * (size_t)src + 1];
        np.nz = raw_nrm[3 * (size_t)src + 2];
        np.edge = raw_edge[src];
        npts2[q] = np;
    }
}
__global__ void add_prefix_kernel(uint32_t *__restrict__ cell_start, const uint32_t *__restrict__ newpref, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cell_start[i] += newpref[i];
}

// level-l node word (l >= 1) of node (cx,cy,cz) from the words one level below (level 0: the cell words, cell_box[])
__device__ __forceinline__ uint32_t node_word_from_children(const uint32_t *__restrict__ below, int nx, int ny, int nz,
                                                            int cx, int cy, int cz)
{
    unsigned m = 0;
    int lo[3] = {15, 15, 15}, hi[3] = {0, 0, 0};
    for (int ch = 0; ch < 8; ++ch) {
        const int x = 2 * cx + (ch & 1), y = 2 * cy + ((ch >> 1) & 1), z = 2 * cz + ((ch >> 2) & 1);
        if (x >= nx || y >= ny || z >= nz) continue;
        const uint32_t w = below[((size_t)z * ny + y) * nx + x];
        if (!(w & 255u)) continue;
        m |= 1u << ch;
        for (int a = 0; a < 3; ++a) {
            lo[a] = min(lo[a], icpb_nibble_up(w, 8 + 4 * a, (ch >> a) & 1));
            hi[a] = max(hi[a], icpb_nibble_up(w, 20 + 4 * a, (ch >> a) & 1));
        }
    }
    return m ? icpb_node_word(m, lo, hi) : 0u;
}
// growing model: the level-l ancestors of the cells a batch of new points fell into (keys[]: level-0 cell keys)
__global__ void node_update_kernel(const uint32_t *__restrict__ keys, uint32_t k, int nx0, int ny0, int level,
                                   const uint32_t *__restrict__ below, int nx, int ny, int nz, int mx, int my,
                                   uint32_t *__restrict__ node)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    const uint32_t key = keys[i];
    const int cx = (int)(key % (uint32_t)nx0) >> level, cy = (int)((key / (uint32_t)nx0) % (uint32_t)ny0) >> level,
              cz = (int)(key / ((uint32_t)nx0 * (uint32_t)ny0)) >> level;
    node[((size_t)cz * my + cy) * mx + cx] = node_word_from_children(below, nx, ny, nz, cx, cy, cz);
}
__global__ void node_levelN_kernel(const uint32_t *__restrict__ below, int nx, int ny, int nz, int mx, int my, int mz,
                                   uint32_t *__restrict__ node)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)mx * my * mz) return;
    const int cx = (int)(t % mx), cy = (int)((t / mx) % my), cz = (int)(t / ((size_t)mx * my));
    node[t] = node_word_from_children(below, nx, ny, nz, cx, cy, cz);
}
// ---------------------------------------------------------------------------
// K2 + K3: transform the source vertex and find its exact nearest model point
// ---------------------------------------------------------------------------
// linear binning of the pair distances over [0, d_box] (fused first step of the trim, see trim_moments_kernel)
constexpr int LIN_BINS = 2048;
#ifndef ICPB_LIN_STRIDE
#define ICPB_LIN_STRIDE 8
#endif
constexpr int LIN_STRIDE = ICPB_LIN_STRIDE; // words between bins: one 32-byte sector per bin spreads the atomics over the L2 slices
__device__ __forceinline__ double lin_scale(double d_box) { return d_box > 0.0 ? (double)LIN_BINS / d_box : 0.0; }
// monotone in d; the search epilogue and the trim use this very expression
__device__ __forceinline__ unsigned lin_bin(double d, double scale)
{
    const double b = d * scale;
    return b >= (double)(LIN_BINS - 1) ? (unsigned)(LIN_BINS - 1) : (unsigned)b;
}
// second-level split of bin `bin` into LIN_BINS sub-bins, monotone in d as well
constexpr int LIN_FINAL_CAP = 512;
__device__ __forceinline__ unsigned lin_sub_bin(double d, double scale, unsigned bin)
{
    const double b = (d * scale - (double)bin) * (double)LIN_BINS;
    return !(b > 0.0) ? 0u : (b >= (double)(LIN_BINS - 1) ? (unsigned)(LIN_BINS - 1) : (unsigned)b);
}

// answer of one query: inclusive ball (find_nearest(val, max)), then the optional edge/normal pair filter
template <bool EAN>
__device__ __forceinline__ unsigned search_finish(uint32_t i, uint32_t pos, double d2, float lb, bool pos_changed, double d_box,
                                                  const double *M, const EanView &ean, uint32_t *__restrict__ nn_pos,
                                                  double *__restrict__ nn_d, float *__restrict__ nn_lb, double &d_stored)
{
    const double d = sqrt(d2);
    bool ok = (pos != ICPB_NONE) && (d <= d_box);
    if (EAN && ok) {
        // EdgeAndNormalPairFilter (icp/icp.hpp:206-221) on the nearest neighbour: neither vertex on a scan edge,
        // normals (source normal mapped by C_local2globalnew.linear()) less than pi/8 apart
        const double ax = ean.snx[i], ay = ean.sny[i], az = ean.snz[i];
        const double bx = (M[0] * ax + M[1] * ay) + M[2] * az;
        const double by = (M[4] * ax + M[5] * ay) + M[6] * az;
        const double bz = (M[8] * ax + M[9] * ay) + M[10] * az;
        const NPoint mn = ean.mn[pos];
        const double dot = (bx * mn.nx + by * mn.ny) + bz * mn.nz;
        const bool edge = ean.sedge[i] || mn.edge;
        const double normal_angle = acos(fmin(1.0, dot));
        ok = !edge && (normal_angle < 3.14159265358979323846 / 8.0);
    }
    if (pos_changed) nn_pos[i] = pos; // the nearest neighbour as such (warm start / skip test of the next iteration) ...
    d_stored = ok ? d : icpb_inf();
    nn_d[i] = d_stored; // ... a pair only when this is finite
    nn_lb[i] = lb;
    return ok ? 1u : 0u;
}

// K2 + K3, one thread per query: transform the source vertex (C_local2globalnew * v, fp64), evaluate the previous
// winner exactly and apply the skip test (icpb_nn_skip_pt): a query whose old winner provably is still the exact
// nearest neighbour -- or that provably has no pair within d_box -- is answered on the spot; the others are searched
// right away (flat block scan / fast path / pyramid walk, grid_view.h).
// Measured and rejected (round 2, cfg2): parking the queries that need a search in a queue (shared memory per block,
// or global memory + a second kernel of long-lived warps that take 32 entries at a time, lists per kind of search) so
// that no lane idles for a skipped neighbour -- 26-28 ms of search per align against 21 ms for this form: the walkers
// of a warp diverge among themselves whatever their neighbours do, and the queue traffic comes on top.
#ifndef ICPB_SEARCH_THREADS
#define ICPB_SEARCH_THREADS 128
#endif
#ifndef ICPB_SEARCH_MIN_BLOCKS
#define ICPB_SEARCH_MIN_BLOCKS 8
#endif
constexpr int SEARCH_THREADS = ICPB_SEARCH_THREADS;
template <bool EAN>
__global__ void __launch_bounds__(SEARCH_THREADS, ICPB_SEARCH_MIN_BLOCKS)
    search_kernel(const GridView g, const double *__restrict__ sx, const double *__restrict__ sy,
                  const double *__restrict__ sz, uint32_t n, IcpState *st, uint32_t *__restrict__ nn_pos,
                  double *__restrict__ nn_d, float *__restrict__ nn_lb, int use_warm, const NNTune tune, const EanView ean,
                  uint32_t *__restrict__ lhist /* null: no fused histogram */)
{
    if (st->done) return;
    __shared__ double M[12]; // C_local2globalnew: broadcast reads instead of per-thread global loads
    __shared__ float dM[12]; // M - Mprev: how the source moved since the previous loop body
    __shared__ NNRun runs[ICPB_BLOCK_ROWS][SEARCH_THREADS]; // run lists of the flat block scan, one column per thread
    __shared__ unsigned sm_cnt[4], sm_ticket;
    if (threadIdx.x < 12) {
        M[threadIdx.x] = st->M[threadIdx.x];
        dM[threadIdx.x] = (float)(st->M[threadIdx.x] - st->Mprev[threadIdx.x]);
    }
    if (threadIdx.x < 4) sm_cnt[threadIdx.x] = 0u;
    if (threadIdx.x == 4) sm_ticket = 0u;
    __syncthreads();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned found = 0, far = 0, nskip = 0, ndefer = 0;
    double dq = icpb_inf(); // the distance entry this thread writes (inf: no pair)
    const double d_box_h = st->d_box;
    if (i < n) {
        const double d_box = st->d_box;
        const double dbox2s = d_box * d_box * (1.0 + 1e-12);
        const bool warm_on = use_warm && st->warm;
        NNTune tn = tune;
        if (!warm_on) tn.skip = 0;
        // source read with evict-first hints: the L2 is for the model, which every query gathers from
        const double vx = __ldcs(sx + i), vy = __ldcs(sy + i), vz = __ldcs(sz + i);
        const uint32_t wp = warm_on ? __ldcs(nn_pos + i) : ICPB_NONE;
        const float lbp = tn.skip ? __ldcs(nn_lb + i) : 0.0f;
        double qx, qy, qz;
        icpb_aff_apply(M, vx, vy, vz, qx, qy, qz); // C_local2globalnew * v (icp/icp.hpp:130)
        MPoint wm;
        wm.x = wm.y = wm.z = 0.0;
        wm.idx = 0;
        const bool has_w = wp < g.n;
        if (has_w) wm = g.pts[wp];
        const float delta = tn.skip ? icpb_motion_bound(dM, (float)vx, (float)vy, (float)vz) : 0.0f;
        double d2w;
        float lbn;
        const int sk = icpb_nn_skip_pt(qx, qy, qz, d_box, has_w, wm, lbp, delta, d2w, lbn);
        double standin;
        if (sk) {
            found = search_finish<EAN>(i, wp, d2w, lbn, false, d_box, M, ean, nn_pos, nn_d, nn_lb, dq);
            nskip = 1;
        } else if (!EAN && tn.skip && icpb_nn_defer(d2w, lbn, d_box, st->tau_guard, standin)) {
            nn_d[i] = dq = standin; // a pair exists (the old winner is within d_box); the trim will not keep it
            nn_lb[i] = lbn;
            found = 1;
            ndefer = 1;
        } else {
            const double need2 = dbox2s < d2w ? dbox2s : d2w;
            const float rs = icpb_cover_radius(g, need2, delta, tn);
            NNState s;
            icpb_nn_init(g, qx, qy, qz, dbox2s, s, (double)rs * (double)rs);
            if (has_w) {
                s.bd2 = d2w;
                s.bidx = wm.idx;
                s.bpos = wp;
                icpb_nn_refresh_bound(s);
            }
            far = icpb_nn_run(g, s, &runs[0][threadIdx.x], SEARCH_THREADS) > 0 ? 1u : 0u;
            found = search_finish<EAN>(i, s.bpos, s.bd2, icpb_nn_lower_bound(g, s), true, d_box, M, ean, nn_pos, nn_d, nn_lb, dq);
        }
    }
    // K4's first step, fused: the linear histogram of the pair distances over [0, d_box] (trim_moments_kernel); one atomic
    // per distinct bin and warp.  Not in the first loop body (d_box = inf), which is trimmed by the radix passes.
    if (lhist && d_box_h < icpb_inf()) {
        const unsigned bin = dq < icpb_inf() ? lin_bin(dq, lin_scale(d_box_h)) : ICPB_NONE;
        const unsigned peers = __match_any_sync(0xFFFFFFFFu, bin);
        if (bin != ICPB_NONE && (peers & ((1u << (threadIdx.x & 31)) - 1u)) == 0) atomicAdd(&lhist[bin * LIN_STRIDE], (unsigned)__popc(peers));
    }
    // counters: per block in shared memory, flushed by the block's last warp (thousands of same-address global atomics
    // per launch were a third of the launch once most queries skip); no block barrier: a finished warp frees its slot
    found = __reduce_add_sync(0xFFFFFFFFu, found);
    far = __reduce_add_sync(0xFFFFFFFFu, far);
    nskip = __reduce_add_sync(0xFFFFFFFFu, nskip);
    ndefer = __reduce_add_sync(0xFFFFFFFFu, ndefer);
    if ((threadIdx.x & 31) == 0) {
        if (found) atomicAdd(&sm_cnt[0], found);
        if (far) atomicAdd(&sm_cnt[1], far);
        if (nskip) atomicAdd(&sm_cnt[2], nskip);
        if (ndefer) atomicAdd(&sm_cnt[3], ndefer);
        __threadfence_block();
        if (atomicAdd(&sm_ticket, 1u) == SEARCH_THREADS / 32 - 1) {
            const unsigned f = atomicAdd(&sm_cnt[0], 0u), w = atomicAdd(&sm_cnt[1], 0u), k = atomicAdd(&sm_cnt[2], 0u),
                           e = atomicAdd(&sm_cnt[3], 0u);
            if (f) atomicAdd(&st->found, (unsigned long long)f);
            if (w) atomicAdd(&st->far, (unsigned long long)w);
            if (k) atomicAdd(&st->skipped, (unsigned long long)k);
            if (e) atomicAdd(&st->deferred, (unsigned long long)e);
        }
    }
}

// stand-alone form for icpb_find_pairs(): queries already in the global frame (AoS)
__global__ void __launch_bounds__(SEARCH_THREADS)
    find_pairs_kernel(const GridView g, const double *__restrict__ q, uint32_t n, double d_box,
                      uint32_t *__restrict__ idx_out, double *__restrict__ d_out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double dbox2s = d_box * d_box * (1.0 + 1e-12);
    uint32_t pos;
    double d2;
    int lvl;
    icpb_nn_search(g, q[3 * (size_t)i], q[3 * (size_t)i + 1], q[3 * (size_t)i + 2], dbox2s, ICPB_NONE, pos, d2, lvl);
    const double d = sqrt(d2);
    const bool ok = (pos != ICPB_NONE) && (d <= d_box);
    idx_out[i] = ok ? (uint32_t)g.pts[pos].idx : ICPB_NONE;
    d_out[i] = ok ? d : d_box; // find_nearest returns (end(), max) when nothing is found
}

// ---------------------------------------------------------------------------
// Fused collectives over NVLink peer memory (multi-GPU: one process per GPU over CUDA IPC, or one process with peer
// access).  Every rank owns an exchange buffer that its peers map.  A collective is "one-shot" and lives inside the
// producing kernel: ONE block per rank stores this rank's small vector into its row of EVERY peer's buffer and reads
// the rows the peers stored into its own.  No separate collective launch, no host involvement.
// Transport: every 32-bit data word travels as ONE 64-bit store {sequence number, data} -- 8-byte stores are single
// transactions over NVLink, so a word is valid exactly when its upper half carries the sequence number of this
// collective: no flags, no system-scope fences, no ordering between words (the "LL" idea of NCCL's low-latency
// protocol).  The receiver polls word by word and copies the data into a plain local buffer that the callers read.
// Two buffer halves alternate (a rank can be at most one collective ahead of a peer, because it needs that peer's
// contribution to finish).  Measured on 4 B200 (cfg2, identical shards on all ranks -- no skew): per loop body the
// fence + flag form of these exchanges cost +43 us in the trim (2 exchanges) and +11 us in the moments (1).
// ---------------------------------------------------------------------------
constexpr int P2P_MAX_RANKS = 8;
constexpr int P2P_ROW_WORDS = 12352; // a histogram (2048 bins + found count) or the candidates of the trim's bin: (key lo, key hi, pair index) each
constexpr int P2P_LIN_CAP = (P2P_ROW_WORDS - 4) / 3; // candidates one row carries in trim_moments_kernel
constexpr int P2P_CAND_CAP = 1055; // candidate keys (u64) one row carries next to their count (radix path)
constexpr size_t P2P_LL_WORDS = 2ull * 2ull * P2P_MAX_RANKS * P2P_ROW_WORDS; // two halves of MAX_RANKS rows of 64-bit {seq, data} entries
constexpr size_t P2P_RECV_WORDS = (size_t)P2P_MAX_RANKS * P2P_ROW_WORDS;     // the rows of the current collective, data only
constexpr size_t P2P_BUFFER_WORDS = P2P_LL_WORDS + P2P_RECV_WORDS + 64;
struct P2PView {
    int world, rank;
    uint32_t *peer[P2P_MAX_RANKS]; // base of every rank's exchange buffer (peer[rank] is the local one)
};
__device__ __forceinline__ unsigned long long *p2p_ll_rows(uint32_t *base, int par)
{
    return reinterpret_cast<unsigned long long *>(base) + (size_t)par * P2P_MAX_RANKS * P2P_ROW_WORDS;
}
// one entry of this rank's own buffer, once it carries sequence number seq (~10 s budget: a peer that is gone turns
// into ICPB_ERR_NCCL instead of a hung GPU)
__device__ __forceinline__ uint32_t p2p_poll(const unsigned long long *entry, uint32_t seq, IcpState *st, bool &dead)
{
    unsigned long long w = *reinterpret_cast<const volatile unsigned long long *>(entry);
    if ((uint32_t)(w >> 32) == seq || dead) return (uint32_t)w;
    const long long t0 = clock64();
    while ((uint32_t)(w >> 32) != seq) {
        if (clock64() - t0 > 20000000000ll) {
            st->comm_error = 1u;
            st->done = 1; // every later launch of this align becomes a no-op
            dead = true;
            break;
        }
        w = *reinterpret_cast<const volatile unsigned long long *>(entry);
    }
    return (uint32_t)w;
}

// Called by all threads of ONE block per rank.  src: nwords 32-bit words (global or shared memory, written before a
// barrier).  var_stride > 0: rows have their own lengths -- word 0 of a row is a count c and the row has
// 1 + var_stride * c words when c <= var_cap, else just that one word.
// Returns the local copy of the rows: row q (from rank q) starts at ret + q * P2P_ROW_WORDS; read with __ldcv.
__device__ const uint32_t *p2p_exchange(const P2PView &v, IcpState *st, const uint32_t *src, int nwords, int var_stride = 0,
                                        int var_cap = 0)
{
    const uint32_t seq = st->comm_seq + 1u;
    const int par = (int)(seq & 1u);
    __syncthreads(); // everyone has read comm_seq; the previous collective's rows are not needed any more
    const unsigned long long tag = (unsigned long long)seq << 32;
    // eight words per thread and trip: the loads of a trip are all in flight before the first dependent store / test
    // (a load -> store loop over possibly aliasing pointers would take one trip to the L2 per word)
    for (int base = 0; base < nwords; base += 8 * (int)blockDim.x) {
        uint32_t w[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = base + u * (int)blockDim.x + (int)threadIdx.x;
            w[u] = i < nwords ? src[i] : 0u;
        }
        for (int p = 0; p < v.world; ++p) {
            unsigned long long *dst = p2p_ll_rows(v.peer[p], par) + (size_t)v.rank * P2P_ROW_WORDS;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = base + u * (int)blockDim.x + (int)threadIdx.x;
                if (i < nwords) dst[i] = tag | w[u];
            }
        }
    }
    uint32_t *recv = v.peer[v.rank] + P2P_LL_WORDS;
    const unsigned long long *mine = p2p_ll_rows(v.peer[v.rank], par);
    bool dead = false;
    for (int q = 0; q < v.world; ++q) { // row by row, eight words per thread and trip
        const unsigned long long *row = mine + (size_t)q * P2P_ROW_WORDS;
        int len = nwords;
        if (var_stride > 0) {
            const uint32_t c = p2p_poll(row, seq, st, dead);
            len = 1 + (c <= (uint32_t)var_cap ? var_stride * (int)c : 0);
        }
        for (int base = 0; base < len; base += 8 * (int)blockDim.x) {
            unsigned long long e[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = base + u * (int)blockDim.x + (int)threadIdx.x;
                e[u] = i < len ? *reinterpret_cast<const volatile unsigned long long *>(row + i) : tag;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = base + u * (int)blockDim.x + (int)threadIdx.x;
                if (i >= len) continue;
                recv[(size_t)q * P2P_ROW_WORDS + i] = (uint32_t)(e[u] >> 32) == seq ? (uint32_t)e[u] : p2p_poll(row + i, seq, st, dead);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) st->comm_seq = seq;
    return recv;
}

// All-reduce (sum) of a 2048-bin histogram held eight bins per thread, plus one extra word of thread 0, straight from
// and to registers: the exchange that opens every trim is on the critical path of the loop body.  All threads of ONE
// block (256) per rank; same transport and sequence numbering as p2p_exchange.
__device__ void p2p_allreduce_hist(const P2PView &v, IcpState *st, const uint32_t (&mine)[8], uint32_t extra,
                                   uint32_t (&sum)[8], unsigned long long &extra_sum)
{
    const uint32_t seq = st->comm_seq + 1u;
    const int par = (int)(seq & 1u);
    __syncthreads();
    const unsigned long long tag = (unsigned long long)seq << 32;
    for (int p = 0; p < v.world; ++p) {
        unsigned long long *dst = p2p_ll_rows(v.peer[p], par) + (size_t)v.rank * P2P_ROW_WORDS;
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[threadIdx.x * 8 + j] = tag | mine[j];
        if (threadIdx.x == 0) dst[8 * blockDim.x] = tag | extra;
    }
    const unsigned long long *rows = p2p_ll_rows(v.peer[v.rank], par);
    bool dead = false;
#pragma unroll
    for (int j = 0; j < 8; ++j) sum[j] = 0u;
    extra_sum = 0ull;
    for (int q0 = 0; q0 < v.world; q0 += 4) { // rank order: the same sums on every rank; four ranks' loads in flight
        unsigned long long e[4][8];
#pragma unroll
        for (int dq = 0; dq < 4; ++dq) {
            const unsigned long long *row = rows + (size_t)(q0 + dq) * P2P_ROW_WORDS;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                e[dq][j] = q0 + dq < v.world ? *reinterpret_cast<const volatile unsigned long long *>(row + threadIdx.x * 8 + j) : tag;
        }
#pragma unroll
        for (int dq = 0; dq < 4; ++dq) {
            if (q0 + dq >= v.world) continue;
            const unsigned long long *row = rows + (size_t)(q0 + dq) * P2P_ROW_WORDS;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                sum[j] += (uint32_t)(e[dq][j] >> 32) == seq ? (uint32_t)e[dq][j] : p2p_poll(row + threadIdx.x * 8 + j, seq, st, dead);
            if (threadIdx.x == 0) extra_sum += p2p_poll(row + 8 * blockDim.x, seq, st, dead);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) st->comm_seq = seq;
}

// ---------------------------------------------------------------------------
// NOTE: IcpState pointers are deliberately NOT __restrict__: the last block of a kernel re-reads fields that another
// thread of the same block has just written (behind __syncthreads); with a noalias pointer the compiler may keep the
// stale value in a register.
// K4: radix select.  Distances are non-negative doubles (or +inf = no pair),
// so their bit patterns order like unsigned integers.
// ---------------------------------------------------------------------------
constexpr int SEL_THREADS = 256;
constexpr int SEL_MAX_BINS = 2048;
constexpr int SEL_D_PASSES = 6;
constexpr int SEL_T_PASSES = 3;
__constant__ int c_sel_d_shift[SEL_D_PASSES] = {52, 41, 30, 19, 8, 0};
__constant__ int c_sel_d_bits[SEL_D_PASSES] = {11, 11, 11, 11, 11, 8};
__constant__ int c_sel_t_shift[SEL_T_PASSES] = {21, 10, 0};
__constant__ int c_sel_t_bits[SEL_T_PASSES] = {11, 11, 10};

// the last block of a pass (or a stand-alone 1-block launch after the histogram
// all-reduce) picks the bucket holding the k-th smallest key and narrows the prefix
__device__ void select_pick(IcpState *st, uint32_t *ghist, int pass, bool tie_mode, uint32_t *smem /* >= 34 */,
                            const uint32_t *rows = nullptr, int world = 1, int rank = 0,
                            unsigned long long krem0 = ~0ull /* pass 0 selects this rank of the keys histogrammed, not min(n_po, found) */)
{
    const int bits = tie_mode ? c_sel_t_bits[pass] : c_sel_d_bits[pass];
    const int shift = tie_mode ? c_sel_t_shift[pass] : c_sel_d_shift[pass];
    const int bins = 1 << bits;
    const int nthr = (int)blockDim.x;
    const int per = (bins + nthr - 1) / nthr;
    unsigned long long krem;
    if (tie_mode)
        krem = (pass == 0) ? st->quota : st->tie_krem;
    else {
        if (pass == 0 && krem0 != ~0ull) {
            krem = krem0;
        } else if (pass == 0) {
            const unsigned long long k = st->n_po < st->found ? st->n_po : st->found; // min(n_po, found)
            if (threadIdx.x == 0) st->sel_k = k;
            krem = k;
        } else
            krem = st->sel_krem;
    }
    uint32_t c[8];
    uint32_t local = 0;
    for (int j = 0; j < per; ++j) {
        const int b = threadIdx.x * per + j;
        c[j] = b < bins ? ghist[b] : 0u;
        local += c[j];
    }
    uint32_t total;
    uint32_t before = block_exclusive_scan(local, smem, total);
    __syncthreads();
    for (int j = 0; j < per; ++j) { // zero for the next use
        const int b = threadIdx.x * per + j;
        if (b < bins) ghist[b] = 0u;
    }
    if (krem == 0 || krem > (unsigned long long)total) {
        // nothing to select (no pairs / n_po == 0): tau = NaN, nothing kept
        if (threadIdx.x == 0) {
            if (!tie_mode) {
                st->sel_krem = 0;
                if (pass == SEL_D_PASSES - 1 || krem == 0) {
                    st->tau = __builtin_nan("");
                    st->cnt_eq = 0;
                    st->quota = 0;
                    st->need_tie = 0;
                    st->p_cut = 0x100000000ll;
                }
            } else {
                st->tie_krem = 0;
                st->p_cut = 0x100000000ll;
            }
        }
        return;
    }
    unsigned long long run = before;
    for (int j = 0; j < per; ++j) {
        const int b = threadIdx.x * per + j;
        if (b < bins && run < krem && krem <= run + c[j]) {
            if (!tie_mode) {
                const unsigned long long prefix = (pass == 0 ? 0ull : st->sel_prefix) | ((unsigned long long)b << shift);
                st->sel_prefix = prefix;
                st->sel_krem = krem - run;
                if (pass == SEL_D_PASSES - 1) {
                    st->tau = __longlong_as_double((long long)prefix);
                    unsigned long long cnt = c[j], quota = krem - run;
                    if (rows && quota < cnt) {
                        // ties are kept in global pair order = lower ranks first: this rank's share of the quota
                        unsigned long long before = 0;
                        for (int q = 0; q < rank; ++q) before += __ldcg(rows + (size_t)q * P2P_ROW_WORDS + b);
                        const unsigned long long local = __ldcg(rows + (size_t)rank * P2P_ROW_WORDS + b);
                        quota = quota > before ? quota - before : 0ull;
                        if (quota > local) quota = local;
                        cnt = local;
                    }
                    st->cnt_eq = cnt;
                    st->quota = quota;
                    st->need_tie = (quota > 0 && quota < cnt) ? 1u : 0u;
                    st->p_cut = (rows && quota == 0 && cnt > 0 && (krem - run) < (unsigned long long)c[j]) ? -1ll : 0x100000000ll;
                }
            } else {
                const unsigned prefix = (pass == 0 ? 0u : st->tie_prefix) | ((unsigned)b << shift);
                st->tie_prefix = prefix;
                st->tie_krem = krem - run;
                if (pass == SEL_T_PASSES - 1) st->p_cut = (long long)prefix;
            }
        }
        run += c[j];
    }
}

// One histogram pass over the distance bits (key = bits(nn_d[i]); digit `pass` of c_sel_d_*).  The last block to finish
// (ticket) narrows the prefix: locally, or after the fused peer-memory all-reduce, or -- NCCL mode -- not at all
// (fused_pick == 0: select_pick_kernel runs after ncclAllReduce).
__global__ void __launch_bounds__(SEL_THREADS)
    select_kernel(const double *__restrict__ nn_d, uint32_t n, IcpState *st, uint32_t *__restrict__ ghist,
                  int pass, int fused_pick, const P2PView p2p)
{
    if (st->done) return;
    if (pass > 0 && st->sel_krem == 0) return; // nothing to select this iteration
    __shared__ uint32_t hist[SEL_MAX_BINS];
    __shared__ uint32_t scan_smem[34];
    __shared__ int is_last;
    const int bits = c_sel_d_bits[pass];
    const int shift = c_sel_d_shift[pass];
    const int bins = 1 << bits;
    for (int b = threadIdx.x; b < bins; b += SEL_THREADS) hist[b] = 0;
    __syncthreads();
    const unsigned bin_mask = (unsigned)bins - 1u;
    {
        const unsigned long long prefix = pass == 0 ? 0ull : st->sel_prefix;
        const int hs = shift + bits; // bits above the current digit must match the prefix
        // four independent loads in flight per thread before the (dependent) shared-memory atomics
        const uint32_t T = gridDim.x * blockDim.x;
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += 4u * T) {
            unsigned long long key[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t j = i + (uint32_t)u * T;
                key[u] = j < n && j >= i ? (unsigned long long)__double_as_longlong(nn_d[j]) : ~0ull;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (key[u] == ~0ull) continue; // out of range (a NaN pattern never occurs in nn_d)
                if (pass == 0 || (key[u] >> hs) == (prefix >> hs))
                    atomicAdd(&hist[(unsigned)(key[u] >> shift) & bin_mask], 1u);
            }
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < bins; b += SEL_THREADS)
        if (hist[b]) atomicAdd(&ghist[b], hist[b]);
    if (!fused_pick) return;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(&st->tickets[pass], 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (threadIdx.x == 0) st->tickets[pass] = 0;
    if (p2p.world > 1) {
        // fused all-reduce of the digit histogram (+ the found count in pass 0) over peer memory
        if (pass == 0 && threadIdx.x == 0) ghist[SEL_MAX_BINS] = (uint32_t)st->found;
        __syncthreads();
        const uint32_t *rows = p2p_exchange(p2p, st, ghist, pass == 0 ? SEL_MAX_BINS + 1 : bins);
        for (int b = threadIdx.x; b < bins; b += SEL_THREADS) {
            uint32_t v = 0;
            for (int q = 0; q < p2p.world; ++q) v += __ldcg(rows + (size_t)q * P2P_ROW_WORDS + b);
            ghist[b] = v;
        }
        if (pass == 0 && threadIdx.x == 0) {
            unsigned long long f = 0;
            for (int q = 0; q < p2p.world; ++q) f += __ldcg(rows + (size_t)q * P2P_ROW_WORDS + SEL_MAX_BINS);
            st->found = f;
        }
        __syncthreads();
        select_pick(st, ghist, pass, false, scan_smem, rows, p2p.world, p2p.rank);
        return;
    }
    select_pick(st, ghist, pass, false, scan_smem);
}

// stand-alone pick (multi-GPU: runs after the histogram all-reduce)
__global__ void __launch_bounds__(SEL_THREADS) select_pick_kernel(IcpState *st, uint32_t *ghist, int pass, int tie_mode)
{
    if (st->done) return;
    if (tie_mode && !st->need_tie) return;
    if (!tie_mode && pass > 0 && st->sel_krem == 0) return;
    __shared__ uint32_t scan_smem[34];
    select_pick(st, ghist, pass, tie_mode != 0, scan_smem);
}

// Tie-break in one launch: which of the pairs at distance == tau survive is decided by pair index (a stable
// sort by (distance, index)).  Only needed when fewer ties are kept than exist (need_tie), which is rare on
// real-valued data, so a single block walks the three index digits back to back instead of three grid launches.
__global__ void __launch_bounds__(1024)
    tie_select_kernel(const double *__restrict__ nn_d, const uint32_t *__restrict__ perm, uint32_t n, IcpState *st,
                      uint32_t *ghist)
{
    if (st->done || !st->need_tie) return;
    __shared__ uint32_t scan_smem[34];
    const unsigned long long tau_bits = (unsigned long long)__double_as_longlong(st->tau);
    for (int pass = 0; pass < SEL_T_PASSES; ++pass) {
        const int bits = c_sel_t_bits[pass], shift = c_sel_t_shift[pass], hs = shift + bits;
        const unsigned bin_mask = (1u << bits) - 1u;
        const unsigned prefix = pass == 0 ? 0u : st->tie_prefix;
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            if ((unsigned long long)__double_as_longlong(nn_d[i]) != tau_bits) continue;
            const uint32_t pi = perm[i];
            if (pass == 0 || hs >= 32 || (pi >> hs) == (prefix >> hs)) atomicAdd(&ghist[(pi >> shift) & bin_mask], 1u);
        }
        __threadfence();
        __syncthreads();
        select_pick(st, ghist, pass, true, scan_smem);
        __threadfence();
        __syncthreads();
    }
}

// Third and last launch of the trim (single GPU and peer-memory mode).  After two grid-wide digit passes the 22-bit
// prefix (exponent + 11 mantissa bits) of the n_po-th smallest distance is known and only a handful of distances
// still match it.  This kernel collects those candidates (warp-aggregated append) and its last block finishes the
// selection on them alone: the remaining four digits (with the fused peer-memory all-reduce per digit when there are
// several ranks) and, if fewer equidistant pairs are kept than exist, the three pair-index digits of the tie-break.
__global__ void __launch_bounds__(SEL_THREADS)
    select_collect_kernel(const double *__restrict__ nn_d, const uint32_t *__restrict__ perm, uint32_t n,
                          IcpState *st, uint32_t *__restrict__ ghist, unsigned long long *__restrict__ cand_key,
                          uint32_t *__restrict__ cand_idx, unsigned int *__restrict__ cand_count, const P2PView p2p,
                          uint32_t *xscratch /* (1 + P2P_MAX_RANKS) * P2P_ROW_WORDS words, zero between uses */)
{
    if (st->done) return;
    if (st->sel_krem == 0) return; // nothing to select this iteration
    __shared__ uint32_t scan_smem[34];
    __shared__ uint32_t fhist[SEL_MAX_BINS]; // histogram of the last block's finish (zero on entry, select_pick re-zeroes it)
    __shared__ int is_last, all_fit;
    for (int b = threadIdx.x; b < SEL_MAX_BINS; b += SEL_THREADS) fhist[b] = 0u;
    const int hs = c_sel_d_shift[1]; // bits >= hs are fixed by the first two passes
    const unsigned long long want = st->sel_prefix >> hs;
    const int lane = threadIdx.x & 31;
    const uint32_t T = gridDim.x * blockDim.x;
    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += 4u * T) { // warp-uniform trip count
        unsigned long long key[4];
        uint32_t idx[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            idx[u] = base + (uint32_t)u * T + threadIdx.x;
            key[u] = (idx[u] < n && idx[u] >= base) ? (unsigned long long)__double_as_longlong(nn_d[idx[u]]) : ~0ull;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const bool hit = key[u] != ~0ull && (key[u] >> hs) == want;
            const unsigned m = __ballot_sync(0xFFFFFFFFu, hit);
            if (!m) continue;
            unsigned pos = 0;
            if (lane == __ffs(m) - 1) pos = atomicAdd(cand_count, (unsigned)__popc(m));
            pos = __shfl_sync(0xFFFFFFFFu, pos, __ffs(m) - 1) + __popc(m & ((1u << lane) - 1u));
            if (hit) {
                cand_key[pos] = key[u];
                cand_idx[pos] = idx[u];
            }
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(&st->tickets[6], 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const unsigned cnt = *((volatile unsigned int *)cand_count);
    // Several ranks: instead of one all-reduce per remaining digit, the ranks exchange their (few) candidates ONCE and
    // every rank finishes the selection on the union -- identical arithmetic everywhere, three exchanges per trim
    // instead of six.  If some rank holds more candidates than a row carries, all ranks fall back to the digit loop.
    const uint32_t *crow = nullptr;
    if (p2p.world > 1) {
        uint32_t *xb = xscratch;
        const bool fits = cnt <= (unsigned)P2P_CAND_CAP;
        if (threadIdx.x == 0) xb[0] = cnt;
        if (fits)
            for (unsigned j = threadIdx.x; j < cnt; j += blockDim.x) {
                const unsigned long long k = __ldcg(cand_key + j);
                xb[1 + 2 * j] = (uint32_t)k;
                xb[2 + 2 * j] = (uint32_t)(k >> 32);
            }
        __threadfence();
        __syncthreads();
        crow = p2p_exchange(p2p, st, xb, fits ? 1 + 2 * (int)cnt : 1, 2, P2P_CAND_CAP);
        if (threadIdx.x == 0) {
            int ok = 1;
            for (int q = 0; q < p2p.world; ++q)
                if (__ldcg(crow + (size_t)q * P2P_ROW_WORDS) > (uint32_t)P2P_CAND_CAP) ok = 0;
            all_fit = ok;
        }
        __syncthreads();
        if (!all_fit) crow = nullptr;
    }
    if (crow) {
        uint32_t *tie_rows = xscratch + P2P_ROW_WORDS; // per-rank counts of the last digit (what select_pick splits the tie quota by)
        for (int pass = 2; pass < SEL_D_PASSES; ++pass) {
            const int bits = c_sel_d_bits[pass], shift = c_sel_d_shift[pass], hsp = shift + bits;
            const unsigned bin_mask = (1u << bits) - 1u;
            const unsigned long long prefix = st->sel_prefix;
            const bool lastp = pass == SEL_D_PASSES - 1;
            if (st->sel_krem == 0) break;
            for (int q = 0; q < p2p.world; ++q) {
                const uint32_t *r = crow + (size_t)q * P2P_ROW_WORDS;
                const unsigned cq = __ldcg(r);
                for (unsigned j = threadIdx.x; j < cq; j += blockDim.x) {
                    const unsigned long long k = (unsigned long long)__ldcg(r + 1 + 2 * j) | ((unsigned long long)__ldcg(r + 2 + 2 * j) << 32);
                    if ((k >> hsp) != (prefix >> hsp)) continue;
                    const unsigned b = (unsigned)(k >> shift) & bin_mask;
                    atomicAdd(&ghist[b], 1u);
                    if (lastp) atomicAdd(&tie_rows[(size_t)q * P2P_ROW_WORDS + b], 1u);
                }
            }
            __threadfence();
            __syncthreads();
            select_pick(st, ghist, pass, false, scan_smem, lastp ? tie_rows : nullptr, p2p.world, p2p.rank);
            __threadfence();
            __syncthreads();
            if (lastp)
                for (int q = 0; q < p2p.world; ++q)
                    for (int b = threadIdx.x; b < (1 << bits); b += blockDim.x) tie_rows[(size_t)q * P2P_ROW_WORDS + b] = 0u;
        }
    } else if (p2p.world > 1) {
        for (int pass = 2; pass < SEL_D_PASSES; ++pass) { // one all-reduce per remaining digit
            const int bits = c_sel_d_bits[pass], shift = c_sel_d_shift[pass], hsp = shift + bits;
            const unsigned bin_mask = (1u << bits) - 1u;
            const unsigned long long prefix = st->sel_prefix;
            if (st->sel_krem == 0) break;
            for (unsigned j = threadIdx.x; j < cnt; j += blockDim.x) {
                const unsigned long long k = __ldcg(cand_key + j);
                if ((k >> hsp) == (prefix >> hsp)) atomicAdd(&ghist[(unsigned)(k >> shift) & bin_mask], 1u);
            }
            __threadfence();
            __syncthreads();
            const int bins = 1 << bits;
            const uint32_t *rows = p2p_exchange(p2p, st, ghist, bins);
            for (int b = threadIdx.x; b < bins; b += blockDim.x) {
                uint32_t v = 0;
                for (int q = 0; q < p2p.world; ++q) v += __ldcg(rows + (size_t)q * P2P_ROW_WORDS + b);
                ghist[b] = v;
            }
            __syncthreads();
            select_pick(st, ghist, pass, false, scan_smem, rows, p2p.world, p2p.rank);
            __threadfence();
            __syncthreads();
        }
    } else {
        // one rank: the few candidates are histogrammed in shared memory -- no global atomics or fences in this tail
        for (int pass = 2; pass < SEL_D_PASSES; ++pass) {
            const int bits = c_sel_d_bits[pass], shift = c_sel_d_shift[pass], hsp = shift + bits;
            const unsigned bin_mask = (1u << bits) - 1u;
            const unsigned long long prefix = st->sel_prefix;
            if (st->sel_krem == 0) break;
            for (unsigned j = threadIdx.x; j < cnt; j += blockDim.x) {
                const unsigned long long k = __ldcg(cand_key + j);
                if ((k >> hsp) == (prefix >> hsp)) atomicAdd(&fhist[(unsigned)(k >> shift) & bin_mask], 1u);
            }
            __syncthreads();
            select_pick(st, fhist, pass, false, scan_smem);
            __syncthreads();
        }
    }
    if (st->need_tie) { // which of the equidistant pairs survive: the lowest pair indices (rank-local quota)
        const unsigned long long tau_bits = (unsigned long long)__double_as_longlong(st->tau);
        for (int pass = 0; pass < SEL_T_PASSES; ++pass) {
            const int bits = c_sel_t_bits[pass], shift = c_sel_t_shift[pass], hst = shift + bits;
            const unsigned bin_mask = (1u << bits) - 1u;
            const unsigned prefix = pass == 0 ? 0u : st->tie_prefix;
            for (unsigned j = threadIdx.x; j < cnt; j += blockDim.x) {
                if (__ldcg(cand_key + j) != tau_bits) continue;
                const uint32_t pi = perm[__ldcg(cand_idx + j)];
                if (pass == 0 || hst >= 32 || (pi >> hst) == (prefix >> hst)) atomicAdd(&fhist[(pi >> shift) & bin_mask], 1u);
            }
            __syncthreads();
            select_pick(st, fhist, pass, true, scan_smem);
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) {
        *cand_count = 0;
        st->tickets[6] = 0;
    }
}

// Exact selection across ranks (selecting block of trim_moments_kernel, source sharded over several GPUs): the s_krem-th
// smallest of ALL ranks' pairs of the chosen bin in the order (distance, rank, pair index) -- "lower ranks first, then
// local position", the global pair order of a contiguous sharding.  Exchange #2 of the loop body carries every rank's
// candidates (key lo, key hi, pair index) to every rank, so all of them finish on the union with identical arithmetic:
// sub-bin split + brute-force ranking when the sub-bin holds a handful, else a bit-by-bit search over the 98-bit
// combined key.  If some rank holds more candidates than a row carries, the bit-by-bit search runs distributed, one
// one-word exchange per bit (masses of equal distances only).
__device__ __forceinline__ bool lin_key_matches(unsigned long long hi, unsigned long long lo, unsigned long long phi,
                                                unsigned long long plo, int pass)
{
    // passes 0..62 decide bits 62..0 of the distance, passes 63..97 bits 34..0 of (rank << 32 | pair index)
    if (pass < 63) {
        const int b = 62 - pass;
        return b == 62 ? true : (hi >> (b + 1)) == (phi >> (b + 1));
    }
    const int b = 97 - pass;
    return hi == phi && (b == 34 ? true : (lo >> (b + 1)) == (plo >> (b + 1)));
}
__device__ __forceinline__ unsigned lin_key_bit(unsigned long long hi, unsigned long long lo, int pass)
{
    return pass < 63 ? (unsigned)(hi >> (62 - pass)) & 1u : (unsigned)(lo >> (97 - pass)) & 1u;
}

// The candidates of all ranks (rows of exchange #2: count, then (key lo, key hi, pair index) each) as one list; a thread
// loads eight of them per trip, all loads in flight before the first use.  off[q]: candidates of the ranks before q.
__device__ __forceinline__ void union_load8(const uint32_t *crow, const unsigned *off, int world, unsigned base,
                                            unsigned long long (&key)[8], uint32_t (&pidx)[8], uint32_t (&rnk)[8], bool (&ok)[8])
{
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const unsigned f = base + (unsigned)u * blockDim.x + threadIdx.x;
        ok[u] = f < off[world];
        key[u] = 0ull, pidx[u] = 0u, rnk[u] = 0u;
        if (ok[u]) {
            int q = 0;
            while (q + 1 < world && f >= off[q + 1]) ++q;
            const uint32_t *r = crow + (size_t)q * P2P_ROW_WORDS + 1 + 3 * (size_t)(f - off[q]);
            key[u] = (unsigned long long)__ldcg(r) | ((unsigned long long)__ldcg(r + 1) << 32);
            pidx[u] = __ldcg(r + 2);
            rnk[u] = (uint32_t)q;
        }
    }
}

__device__ void trim_linear_finish_multi(const double *__restrict__ nn_d, const uint32_t *__restrict__ perm, IcpState *st,
                                         const unsigned long long *cand_key, const uint32_t *cand_idx, unsigned cnt, unsigned bin,
                                         unsigned long long krem, const P2PView &p2p, uint32_t *xb, uint32_t *fhist,
                                         uint32_t *scan_smem, unsigned long long *f_key, uint32_t *f_perm)
{
    __shared__ unsigned m_cnt, m_sub, m_allfit, m_zero;
    __shared__ unsigned long long m_krem, m_phi, m_plo;
    __shared__ uint32_t f_rank[LIN_FINAL_CAP];
    const double scale = lin_scale(st->d_box);
    const bool fits = cnt <= (unsigned)P2P_LIN_CAP;
    __shared__ unsigned u_off[P2P_MAX_RANKS + 1];
    if (threadIdx.x == 0) xb[0] = cnt;
    if (fits)
        for (unsigned j0 = 0; j0 < cnt; j0 += 4u * blockDim.x) { // four per trip: loads, dependent loads, then stores
            unsigned long long k4[4];
            uint32_t i4[4], p4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned j = j0 + (unsigned)u * blockDim.x + threadIdx.x;
                k4[u] = j < cnt ? __ldcg(cand_key + j) : 0ull;
                i4[u] = j < cnt ? __ldcg(cand_idx + j) : 0u;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) p4[u] = perm[i4[u]];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned j = j0 + (unsigned)u * blockDim.x + threadIdx.x;
                if (j < cnt) {
                    xb[1 + 3 * j] = (uint32_t)k4[u];
                    xb[2 + 3 * j] = (uint32_t)(k4[u] >> 32);
                    xb[3 + 3 * j] = p4[u];
                }
            }
        }
    __threadfence();
    __syncthreads();
    const uint32_t *crow = p2p_exchange(p2p, st, xb, fits ? 1 + 3 * (int)cnt : 1, 3, P2P_LIN_CAP);
    if (threadIdx.x == 0) {
        unsigned ok = 1, run = 0;
        for (int q = 0; q < p2p.world; ++q) {
            const uint32_t cq = __ldcg(crow + (size_t)q * P2P_ROW_WORDS);
            if (cq > (uint32_t)P2P_LIN_CAP) ok = 0;
            u_off[q] = run;
            run += cq <= (uint32_t)P2P_LIN_CAP ? cq : 0u;
        }
        u_off[p2p.world] = run;
        m_allfit = ok;
        m_cnt = 0u;
        m_sub = ICPB_NONE;
    }
    __syncthreads();
    const unsigned u_total = u_off[p2p.world];
    unsigned long long w_key = 0, w_lo = 0; // the selected pair: distance bits, (rank << 32 | pair index)
    bool have = false;
    if (m_allfit) {
        // sub-bin split over the union
        for (unsigned base = 0; base < u_total; base += 8u * blockDim.x) {
            unsigned long long k8[8];
            uint32_t p8[8], r8[8];
            bool ok8[8];
            union_load8(crow, u_off, p2p.world, base, k8, p8, r8, ok8);
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (ok8[u]) atomicAdd(&fhist[lin_sub_bin(__longlong_as_double((long long)k8[u]), scale, bin)], 1u);
        }
        __syncthreads();
        uint32_t c2[8], local2 = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            c2[j] = fhist[threadIdx.x * 8 + j];
            local2 += c2[j];
        }
        uint32_t total2;
        const uint32_t before2 = block_exclusive_scan(local2, scan_smem, total2);
        {
            unsigned long long run = before2;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (run < krem && krem <= run + c2[j]) {
                    m_sub = threadIdx.x * 8 + j;
                    m_krem = krem - run;
                }
                run += c2[j];
            }
        }
        __syncthreads();
        for (int b = threadIdx.x; b < SEL_MAX_BINS; b += SEL_THREADS) fhist[b] = 0u;
        const unsigned sub = m_sub;
        for (unsigned base = 0; base < u_total && sub != ICPB_NONE; base += 8u * blockDim.x) {
            unsigned long long k8[8];
            uint32_t p8[8], r8[8];
            bool ok8[8];
            union_load8(crow, u_off, p2p.world, base, k8, p8, r8, ok8);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (!ok8[u] || lin_sub_bin(__longlong_as_double((long long)k8[u]), scale, bin) != sub) continue;
                const unsigned at = atomicAdd(&m_cnt, 1u);
                if (at < LIN_FINAL_CAP) {
                    f_key[at] = k8[u];
                    f_rank[at] = r8[u];
                    f_perm[at] = p8[u];
                }
            }
        }
        __syncthreads();
        const unsigned m = m_cnt;
        if (sub != ICPB_NONE && m <= LIN_FINAL_CAP) {
            if (threadIdx.x == 0) m_zero = 0u;
            __syncthreads();
            const unsigned long long want = m_krem - 1ull;
            for (unsigned t = threadIdx.x; t < m; t += blockDim.x) {
                const unsigned long long kt = f_key[t], lt = ((unsigned long long)f_rank[t] << 32) | f_perm[t];
                unsigned rank = 0;
                for (unsigned j = 0; j < m; ++j) {
                    const unsigned long long lj = ((unsigned long long)f_rank[j] << 32) | f_perm[j];
                    rank += (f_key[j] < kt || (f_key[j] == kt && lj < lt)) ? 1u : 0u;
                }
                if ((unsigned long long)rank == want) {
                    m_phi = kt;
                    m_plo = lt;
                    m_zero = 1u;
                }
            }
            __syncthreads();
            have = m_zero != 0u;
            w_key = m_phi;
            w_lo = m_plo;
        }
    }
    if (!have) {
        // bit-by-bit search for the krem-th smallest combined key: over the union when every rank's candidates arrived,
        // else over the local candidates with the counts summed across the ranks (one exchange per bit)
        const bool uni = m_allfit != 0u;
        if (threadIdx.x == 0) m_phi = 0ull, m_plo = 0ull, m_krem = krem;
        __syncthreads();
        for (int pass = 0; pass < 98; ++pass) {
            if (threadIdx.x == 0) m_zero = 0u;
            __syncthreads();
            const unsigned long long phi = m_phi, plo = m_plo;
            unsigned zeros = 0;
            if (uni) {
                for (unsigned base = 0; base < u_total; base += 8u * blockDim.x) {
                    unsigned long long k8[8];
                    uint32_t p8[8], r8[8];
                    bool ok8[8];
                    union_load8(crow, u_off, p2p.world, base, k8, p8, r8, ok8);
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const unsigned long long lo = ((unsigned long long)r8[u] << 32) | p8[u];
                        if (ok8[u] && lin_key_matches(k8[u], lo, phi, plo, pass) && lin_key_bit(k8[u], lo, pass) == 0u) ++zeros;
                    }
                }
            } else {
                for (unsigned j = threadIdx.x; j < cnt; j += blockDim.x) {
                    const unsigned long long hi = __ldcg(cand_key + j);
                    const unsigned long long lo = ((unsigned long long)p2p.rank << 32) | perm[__ldcg(cand_idx + j)];
                    if (lin_key_matches(hi, lo, phi, plo, pass) && lin_key_bit(hi, lo, pass) == 0u) ++zeros;
                }
            }
            zeros = __reduce_add_sync(0xFFFFFFFFu, zeros);
            if ((threadIdx.x & 31) == 0 && zeros) atomicAdd(&m_zero, zeros);
            __syncthreads();
            unsigned long long z = m_zero;
            if (!uni) {
                if (threadIdx.x == 0) xb[0] = m_zero;
                __threadfence();
                __syncthreads();
                const uint32_t *rr = p2p_exchange(p2p, st, xb, 1);
                z = 0;
                for (int q = 0; q < p2p.world; ++q) z += __ldcg(rr + (size_t)q * P2P_ROW_WORDS);
                __syncthreads();
            }
            if (threadIdx.x == 0) {
                if (m_krem > z) { // the wanted key has this bit set
                    m_krem -= z;
                    if (pass < 63) m_phi |= 1ull << (62 - pass);
                    else m_plo |= 1ull << (97 - pass);
                }
            }
            __syncthreads();
        }
        w_key = m_phi;
        w_lo = m_plo;
    }
    if (threadIdx.x == 0) {
        const int q_w = (int)(w_lo >> 32);
        st->tau = __longlong_as_double((long long)w_key);
        st->p_cut = p2p.rank < q_w ? 0x100000000ll : (p2p.rank == q_w ? (long long)(w_lo & 0xFFFFFFFFull) : -1ll);
        st->need_tie = 0;
        st->sel_krem = 1;
    }
    for (int b = threadIdx.x; b < SEL_MAX_BINS; b += SEL_THREADS) fhist[b] = 0u;
    __syncthreads();
}

__device__ __forceinline__ bool pair_kept(double d, uint32_t i, double tau, long long p_cut)
{
    return d < tau || (d == tau && (long long)i <= p_cut);
}

// ---------------------------------------------------------------------------
// K5 (+K6): moments over the kept pairs, deterministic two-stage reduction,
// closed-form pose step in the last block
// ---------------------------------------------------------------------------
constexpr int MOM_THREADS = 256;

__device__ __forceinline__ double shfl_xor_double(double v, int o)
{
    return __shfl_xor_sync(0xFFFFFFFFu, v, o);
}

#ifdef ICPB_PHASE_STAMPS
__device__ __forceinline__ void phase_stamp(IcpState *st, int slot)
{
    if (threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        st->stamps[slot] = t;
    }
}
#define ICPB_STAMP(cond, slot) do { if (cond) phase_stamp(st, slot); } while (0)
#else
#define ICPB_STAMP(cond, slot)
#endif
// one kept pair: p = the transformed source vertex (C_local2globalnew * v), m = its model point, d = their distance
__device__ __forceinline__ void moments_add(double (&acc)[ICPB_NSUMS], const double *M, double vx, double vy, double vz,
                                            const MPoint &m, double d)
{
    double px, py, pz;
    icpb_aff_apply(M, vx, vy, vz, px, py, pz);
    acc[0] += px;
    acc[1] += py;
    acc[2] += pz;
    acc[3] += m.x;
    acc[4] += m.y;
    acc[5] += m.z;
    acc[6] += px * m.x;
    acc[7] += px * m.y;
    acc[8] += px * m.z;
    acc[9] += py * m.x;
    acc[10] += py * m.y;
    acc[11] += py * m.z;
    acc[12] += pz * m.x;
    acc[13] += pz * m.y;
    acc[14] += pz * m.z;
    acc[15] += d * d; // mu_d += d*d (icp/icp.cpp:60-61)
    acc[16] += 1.0;
}

// The moments end in two steps.  moments_block_partial: warp shuffles -> this block's row of `partials`.
// moments_final (the block that takes the last ticket): reduces the rows in a fixed order (run-to-run deterministic),
// exchanges the 17 sums with the other ranks over peer memory, runs the pose step (K6) and drives the WHILE node of the
// graph.  It works on a shared-memory copy of the loop state (one coalesced read, issued together with the loads of the
// partials, one coalesced write-back): the pose step is a chain of ~60 state accesses by a single thread.
#ifdef ICPB_PHASE_STAMPS
constexpr int MOM_STATE_WORDS = (int)(offsetof(IcpState, stamps) / 4);
#else
constexpr int MOM_STATE_WORDS = (int)(sizeof(IcpState) / 4);
#endif
static_assert(sizeof(IcpState) % 4 == 0 && MOM_STATE_WORDS <= 2 * MOM_THREADS, "state copy: at most two words per thread");
struct MomShared {
    double sm[MOM_THREADS / 32][ICPB_NSUMS];
    double extra[ICPB_NSUMS]; // a row added behind the block partials (trim_moments_kernel: the kept pairs of the bin)
    IcpState st;
    int is_last;
};
// all threads of the block; the row lands in dst[0..16] (global or shared)
__device__ __forceinline__ void moments_block_partial(double (&acc)[ICPB_NSUMS], double *dst, MomShared &ms)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < ICPB_NSUMS; ++k) {
        double v = acc[k];
        for (int o = 16; o > 0; o >>= 1) v += shfl_xor_double(v, o);
        if (lane == 0) ms.sm[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < ICPB_NSUMS) {
        double v = 0.0;
        for (int w = 0; w < MOM_THREADS / 32; ++w) v += ms.sm[w][threadIdx.x];
        dst[threadIdx.x] = v;
    }
}
// The block that took the last ticket (all threads), after a fence: requests its copy of the loop state and its share
// of the partial rows together (one trip to the L2 for both).  Column k by warp (k % 8), lanes stride the rows.
__device__ __forceinline__ void moments_last_begin(const IcpState *st, const double *__restrict__ partials, unsigned rows,
                                                   double (&v)[3], MomShared &ms)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t *src = reinterpret_cast<const uint32_t *>(st);
    uint32_t *dst = reinterpret_cast<uint32_t *>(&ms.st);
    uint32_t w0 = 0, w1 = 0;
    if ((int)threadIdx.x < MOM_STATE_WORDS) w0 = __ldcg(src + threadIdx.x);
    if ((int)threadIdx.x + MOM_THREADS < MOM_STATE_WORDS) w1 = __ldcg(src + threadIdx.x + MOM_THREADS);
    v[0] = v[1] = v[2] = 0.0;
    for (unsigned b0 = 0; b0 < rows; b0 += 320u) {
        double x[3][10];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int k = warp + c * (MOM_THREADS / 32);
#pragma unroll
            for (int j = 0; j < 10; ++j) {
                const unsigned b = b0 + (unsigned)lane + 32u * (unsigned)j;
                x[c][j] = (k < ICPB_NSUMS && b < rows) ? __ldcg(partials + (size_t)b * ICPB_NSUMS + k) : 0.0;
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int j = 0; j < 10; ++j) v[c] += x[c][j];
    }
    if ((int)threadIdx.x < MOM_STATE_WORDS) dst[threadIdx.x] = w0;
    if ((int)threadIdx.x + MOM_THREADS < MOM_STATE_WORDS) dst[threadIdx.x + MOM_THREADS] = w1;
}
// ... and ends the loop body on the shared-memory state ms.st: final sums (+ ms.extra), exchange with the other ranks,
// pose step, loop condition, state written back.  v: from moments_last_begin.
__device__ void moments_last_finish(IcpState *st, double (&v)[3], bool with_extra, double *__restrict__ trace, int trace_cap,
                                    int fused_pose, const P2PView &p2p, cudaGraphConditionalHandle cond, int use_cond,
                                    MomShared &ms)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads(); // ms.sm is free (moments_block_partial), ms.extra and ms.st are written
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int k = warp + c * (MOM_THREADS / 32);
        double t = v[c];
        for (int o = 16; o > 0; o >>= 1) t += shfl_xor_double(t, o);
        if (lane == 0 && k < ICPB_NSUMS) ms.sm[0][k] = with_extra ? t + ms.extra[k] : t;
    }
    __syncthreads();
    ICPB_STAMP(true, 10);
    if (p2p.world > 1) {
        // fused all-reduce of the 17 moment sums over peer memory; rows are summed in rank order on every rank
        const uint32_t *prow = p2p_exchange(p2p, &ms.st, reinterpret_cast<const uint32_t *>(&ms.sm[0][0]), 2 * ICPB_NSUMS);
        double t = 0.0;
        if (threadIdx.x < ICPB_NSUMS) {
            for (int q = 0; q < p2p.world; ++q) {
                const uint32_t *r = prow + (size_t)q * P2P_ROW_WORDS + 2 * threadIdx.x;
                const unsigned long long bits = (unsigned long long)__ldcg(r) | ((unsigned long long)__ldcg(r + 1) << 32);
                t += __longlong_as_double((long long)bits);
            }
        }
        __syncthreads();
        if (threadIdx.x < ICPB_NSUMS) ms.sm[0][threadIdx.x] = t;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        for (int k = 0; k < ICPB_NSUMS; ++k) ms.st.sums[k] = ms.sm[0][k];
        ms.st.tickets[15] = 0;
        if (fused_pose) {
            const int row = ms.st.executed;
            icpb_iteration_finish(ms.st, row < trace_cap ? trace + (size_t)row * ICPB_TRACE_COLS_DEV : nullptr);
        }
        ICPB_STAMP(true, 11);
        // CUDA-graph mode: the loop condition of Trimmed::_align drives the WHILE node this kernel lives in
        if (use_cond) cudaGraphSetConditional(cond, ms.st.done ? 0u : 1u);
    }
    __syncthreads();
    {
        uint32_t *dst = reinterpret_cast<uint32_t *>(st);
        const uint32_t *src = reinterpret_cast<const uint32_t *>(&ms.st);
        for (int w = threadIdx.x; w < MOM_STATE_WORDS; w += MOM_THREADS) dst[w] = src[w];
    }
}

// ---------------------------------------------------------------------------
// K4 + K5 + K6, loop bodies after the first, ONE launch.  Every pair distance lies in [0, d_box] with d_box = 2 *
// (previous tau) finite, so the search kernel bins the distances of the pairs it finds LINEARLY over that interval in
// its epilogue (2048 bins, one warp-aggregated global atomic per distinct bin and warp).  This kernel (grid = the
// co-resident blocks of the device, so that its blocks may wait for each other):
//   every block  scans the histogram (8 KB) and finds the bin that holds the n_po-th smallest distance; then passes
//                over its share of the pairs once: a pair of a LOWER bin is kept whatever the exact tau is -- its
//                moments are accumulated on the spot (Pairs::getTransform, icp/icp.cpp:55-73) --, the pairs OF the bin
//                (a few hundred to a few thousand of 10^6) are collected;
//   the block that arrives last selects exactly among the collected keys (one more linear split into 2048 sub-bins,
//                brute-force ranking of the handful left; radix digits as the general fall-back) -> tau, p_cut, and
//                releases the others;
//   every block  adds those pairs of the bin that are kept (its own elements, in its own fixed order: the sums are
//                run-to-run deterministic), block partials, ticket;
//   last block   fixed-order reduction, exchange of the sums (multi-GPU), pose step, loop condition.
// fuse == 0 stops after the selection (stand-alone trim; moments_kernel follows as a launch of its own).
// The first loop body (d_box = inf) and the stand-alone icpb_trim keep the radix passes above.
// ---------------------------------------------------------------------------
// candidate segments of trim_moments_kernel: one per warp, large enough for all the pairs a warp visits
constexpr int TRIM_MAX_BLOCKS = 320;
constexpr int TRIM_SEG_PER_THREAD = TRIM_MAX_BLOCKS * (SEL_THREADS / 32) / SEL_THREADS; // segments one thread of the selecting block gathers
__host__ __device__ __forceinline__ uint32_t trim_seg_cap(uint32_t n, unsigned blocks)
{
    const uint32_t share = (uint32_t)(((unsigned long long)n + blocks - 1) / blocks); // largest share of a block
    return ((share + 4u * SEL_THREADS - 1u) / (4u * SEL_THREADS)) * 128u;                 // trips * 4 * 32 lanes
}
__device__ __forceinline__ bool spin_until(volatile unsigned int *flag, unsigned target, IcpState *st)
{
    const long long t0 = clock64();
    while (*flag != target) {
        if (clock64() - t0 > 24000000000ll) { // ~12 s: the block that was to publish never did; do not hang the GPU
            st->comm_error = 1u;
            return false;
        }
    }
    return true;
}

__global__ void __launch_bounds__(SEL_THREADS, 2)
    trim_moments_kernel(const GridView g, const double *__restrict__ sx, const double *__restrict__ sy,
                        const double *__restrict__ sz, const uint32_t *__restrict__ nn_pos,
                        const double *__restrict__ nn_d, const uint32_t *__restrict__ perm, uint32_t n, IcpState *st,
                        uint32_t *__restrict__ lhist, unsigned long long *__restrict__ cand_key,
                        uint32_t *__restrict__ cand_idx, unsigned long long *__restrict__ seg_key,
                        uint32_t *__restrict__ seg_idx, unsigned int *__restrict__ seg_count, const P2PView p2p,
                        uint32_t *xscratch /* P2P_ROW_WORDS words */, uint32_t *lsum /* LIN_BINS words: the histogram of all ranks */,
                        double *__restrict__ partials, double *__restrict__ trace, int trace_cap,
                        cudaGraphConditionalHandle cond, int use_cond, int fuse)
{
    // the scalars this block needs, requested together (each is a round trip to the L2)
    const int done_in = *((volatile int *)&st->done);
    const unsigned lin_target = *((volatile unsigned int *)&st->lin_epoch) + 1u;
    const unsigned long long found_in = *((volatile unsigned long long *)&st->found);
    const unsigned long long n_po = *((volatile unsigned long long *)&st->n_po);
    const double d_box = *((volatile double *)&st->d_box);
    const bool multi = p2p.world > 1;
    uint32_t c[8];
    if (!multi) {
#pragma unroll
        for (int j = 0; j < 8; ++j) c[j] = __ldcg(lhist + (threadIdx.x * 8 + j) * LIN_STRIDE);
    }
    const double m_in = threadIdx.x < 12 ? st->M[threadIdx.x] : 0.0;
    if (done_in) return;
    __shared__ uint32_t fhist[SEL_MAX_BINS];
    __shared__ uint32_t scan_smem[34];
    __shared__ int is_last;
    __shared__ unsigned s_bin;
    __shared__ unsigned long long s_krem;
    __shared__ MomShared ms;
    __shared__ double M[12]; // C_local2globalnew: broadcast reads instead of 24 registers per thread
    static_assert(LIN_BINS == SEL_THREADS * 8, "eight bins per thread");
    static_assert(MOM_THREADS == SEL_THREADS, "one block shape for the trim and the moments");
    ICPB_STAMP(blockIdx.x == 0, 0);
    if (threadIdx.x < 12) M[threadIdx.x] = m_in;
    if (multi) {
        // Several ranks (source sharded): exchange #1 of the loop body.  Block 0 sends this rank's histogram and pair count
        // to every peer over NVLink peer memory, sums what it receives and publishes the sums; the other blocks (all
        // co-resident) wait for that flag.  lin_epoch is advanced by the selecting block, i.e. after every block has
        // read it (above).
        const unsigned target = lin_target;
        if (blockIdx.x == 0) {
            {
                uint32_t h8[8], a8[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) h8[j] = __ldcg(lhist + (threadIdx.x * 8 + j) * LIN_STRIDE);
                unsigned long long f = 0;
                p2p_allreduce_hist(p2p, st, h8, (uint32_t)found_in, a8, f);
#pragma unroll
                for (int j = 0; j < 8; ++j) lsum[threadIdx.x * 8 + j] = a8[j];
                if (threadIdx.x == 0) st->found = f;
            }
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) atomicExch(&st->lin_ready, target);
        } else {
            if (threadIdx.x == 0) spin_until(&st->lin_ready, target, st);
            __syncthreads();
            __threadfence();
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) c[j] = __ldcg(lsum + threadIdx.x * 8 + j);
    }
    const unsigned long long found_all = multi ? *((volatile unsigned long long *)&st->found) : found_in;
    const unsigned long long k = n_po < found_all ? n_po : found_all; // min(n_po, found)
    if (threadIdx.x == 0) s_bin = ICPB_NONE;
    // which bin holds the k-th smallest distance (every block finds the same one)
    uint32_t local = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) local += c[j];
    uint32_t total;
    const uint32_t before = block_exclusive_scan(local, scan_smem, total);
    if (k > 0 && k <= (unsigned long long)total) {
        unsigned long long run = before;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (run < k && k <= run + c[j]) {
                s_bin = threadIdx.x * 8 + j;
                s_krem = k - run;
            }
            run += c[j];
        }
    }
    for (int b = threadIdx.x; b < SEL_MAX_BINS; b += SEL_THREADS) fhist[b] = 0u;
    __syncthreads();
    const unsigned bin = s_bin;
    const double scale = lin_scale(d_box);
    // this block's share of the pairs: a contiguous range (equal shares to within one pair)
    const uint32_t e_lo = (uint32_t)((unsigned long long)n * blockIdx.x / gridDim.x);
    const uint32_t e_hi = (uint32_t)((unsigned long long)n * (blockIdx.x + 1u) / gridDim.x);
    ICPB_STAMP(blockIdx.x == 0, 1);
    double acc[ICPB_NSUMS];
#pragma unroll
    for (int q = 0; q < ICPB_NSUMS; ++q) acc[q] = 0.0;
    // the pairs of the bin go to a segment of their own per warp, in the order the warp meets them: no atomics, and
    // the order (hence every sum formed over them later) is the same in every run
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t seg_cap = trim_seg_cap(n, gridDim.x);
    const size_t seg_base = ((size_t)blockIdx.x * (SEL_THREADS / 32) + (size_t)warp) * seg_cap;
    unsigned wcnt = 0; // warp-uniform
    if (bin != ICPB_NONE) {
        // four pairs per trip (block-uniform trip count): their coalesced loads are all in flight before the first
        // dependent gather of a model point
        for (uint32_t base = e_lo; base < e_hi; base += 4u * SEL_THREADS) {
            double d[4], vx[4], vy[4], vz[4];
            uint32_t idx[4], pos[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                idx[u] = base + (uint32_t)u * SEL_THREADS + threadIdx.x;
                const bool in = idx[u] < e_hi;
                d[u] = in ? nn_d[idx[u]] : icpb_inf();
                if (fuse) {
                    vx[u] = in ? sx[idx[u]] : 0.0;
                    vy[u] = in ? sy[idx[u]] : 0.0;
                    vz[u] = in ? sz[idx[u]] : 0.0;
                    pos[u] = in ? nn_pos[idx[u]] : ICPB_NONE;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned b = d[u] < icpb_inf() ? lin_bin(d[u], scale) : ICPB_NONE;
                if (fuse && b < bin) moments_add(acc, M, vx[u], vy[u], vz[u], g.pts[pos[u]], d[u]); // kept whatever tau is
                const bool hit = b == bin;
                const unsigned m = __ballot_sync(0xFFFFFFFFu, hit);
                if (!m) continue;
                if (hit) {
                    const size_t at = seg_base + wcnt + (unsigned)__popc(m & ((1u << lane) - 1u));
                    seg_key[at] = (unsigned long long)__double_as_longlong(d[u]);
                    seg_idx[at] = idx[u];
                }
                wcnt += (unsigned)__popc(m);
            }
        }
    }
    if (lane == 0) seg_count[blockIdx.x * (SEL_THREADS / 32) + warp] = wcnt;
    if (fuse) moments_block_partial(acc, partials + (size_t)blockIdx.x * ICPB_NSUMS, ms);
    ICPB_STAMP(blockIdx.x == 0, 2);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(&st->tickets[7], 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!is_last) return;
    // ---- the block that arrived last: selection, the kept pairs of the bin, final sums, pose ----
    // It works on a shared-memory copy of the loop state (ls) from here on; everything it needs from the L2 first --
    // the state, the segment counts, its share of the partial rows -- is requested in one go.
    ICPB_STAMP(true, 3);
    __threadfence();
    __shared__ uint32_t seg_off[TRIM_MAX_BLOCKS * (SEL_THREADS / 32) + 1];
    const unsigned nseg = gridDim.x * (SEL_THREADS / 32);
    const unsigned s0 = threadIdx.x * TRIM_SEG_PER_THREAD;
    unsigned cn[TRIM_SEG_PER_THREAD], mine = 0;
#pragma unroll
    for (int i = 0; i < TRIM_SEG_PER_THREAD; ++i) cn[i] = (bin != ICPB_NONE && s0 + i < nseg) ? __ldcg(seg_count + s0 + i) : 0u;
    double rowsum[3];
    moments_last_begin(st, partials, fuse ? gridDim.x : 0u, rowsum, ms);
    for (int b = threadIdx.x; b < LIN_BINS; b += SEL_THREADS) lhist[b * LIN_STRIDE] = 0u; // for the next loop body's search
#pragma unroll
    for (int i = 0; i < TRIM_SEG_PER_THREAD; ++i) mine += cn[i];
    IcpState *ls = &ms.st;
    // the warps' segments, in segment order, form one list: candidate j lives in the segment s with seg_off[s] <= j < seg_off[s+1]
    uint32_t total_c;
    {
        unsigned at = block_exclusive_scan(mine, scan_smem, total_c);
#pragma unroll
        for (int i = 0; i < TRIM_SEG_PER_THREAD; ++i) {
            seg_off[s0 + i] = at;
            at += cn[i];
        }
        if (threadIdx.x == SEL_THREADS - 1) seg_off[TRIM_MAX_BLOCKS * (SEL_THREADS / 32)] = at;
        if (threadIdx.x == 0) ls->sel_k = k;
    }
    __syncthreads();
    const unsigned cnt = total_c;
    // the first 4 * 256 candidates stay in registers (rk: distance bits, ri: pair); all go to the flat list as well
    unsigned long long rk[4];
    uint32_t ri[4];
    {
        constexpr int NS = TRIM_MAX_BLOCKS * (SEL_THREADS / 32);
        for (unsigned j0 = 0; j0 < cnt; j0 += 4u * SEL_THREADS) {
            unsigned long long kk[4];
            uint32_t ii[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned j = j0 + threadIdx.x + (unsigned)u * SEL_THREADS;
                kk[u] = 0ull, ii[u] = 0u;
                if (j < cnt) {
                    unsigned lo = 0, hi = NS; // last s in [0, NS] with seg_off[s] <= j (seg_off[NS] = cnt > j)
                    while (hi - lo > 1u) {
                        const unsigned mid = (lo + hi) >> 1;
                        if (seg_off[mid] <= j) lo = mid;
                        else hi = mid;
                    }
                    const size_t src = (size_t)lo * seg_cap + (j - seg_off[lo]);
                    kk[u] = __ldcg(seg_key + src);
                    ii[u] = __ldcg(seg_idx + src);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned j = j0 + threadIdx.x + (unsigned)u * SEL_THREADS;
                if (j < cnt) {
                    cand_key[j] = kk[u];
                    cand_idx[j] = ii[u];
                }
                if (j0 == 0) rk[u] = kk[u], ri[u] = ii[u];
            }
        }
        if (cnt == 0) {
#pragma unroll
            for (int u = 0; u < 4; ++u) rk[u] = 0ull, ri[u] = 0u;
        }
        if (cnt > 4u * SEL_THREADS || multi) __threadfence(); // the flat list is read right away
        __syncthreads();
    }
    // Exact selection among the collected pairs: the s_krem-th smallest by (distance, pair index) -- tau is its distance
    // and p_cut its pair index (pair_kept: d < tau, or d == tau with an index <= p_cut: a stable order, like the oracle).
    // One more linear split of the bin into 2048 sub-bins leaves a handful of pairs, ranked by brute force.
    __shared__ unsigned long long f_key[LIN_FINAL_CAP];
    __shared__ uint32_t f_perm[LIN_FINAL_CAP];
    __shared__ unsigned f_cnt, f_sub;
    __shared__ unsigned long long f_krem;
    bool done_fast = false;
    if (multi && bin != ICPB_NONE) {
        done_fast = true;
        trim_linear_finish_multi(nn_d, perm, ls, cand_key, cand_idx, cnt, bin, s_krem, p2p, xscratch, fhist, scan_smem, f_key, f_perm);
    } else if (bin != ICPB_NONE) {
        const unsigned long long krem = s_krem;
        if (threadIdx.x == 0) f_cnt = 0u, f_sub = ICPB_NONE;
        uint32_t rsub[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned j = threadIdx.x + (unsigned)u * SEL_THREADS;
            rsub[u] = lin_sub_bin(__longlong_as_double((long long)rk[u]), scale, bin);
            if (j < cnt) atomicAdd(&fhist[rsub[u]], 1u);
        }
        for (unsigned j = threadIdx.x + 4u * SEL_THREADS; j < cnt; j += blockDim.x)
            atomicAdd(&fhist[lin_sub_bin(__longlong_as_double((long long)__ldcg(cand_key + j)), scale, bin)], 1u);
        __syncthreads();
        uint32_t c2[8], local2 = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            c2[j] = fhist[threadIdx.x * 8 + j];
            local2 += c2[j];
        }
        uint32_t total2;
        const uint32_t before2 = block_exclusive_scan(local2, scan_smem, total2);
        {
            unsigned long long run = before2;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (run < krem && krem <= run + c2[j]) {
                    f_sub = threadIdx.x * 8 + j;
                    f_krem = krem - run;
                }
                run += c2[j];
            }
        }
        __syncthreads();
        for (int b = threadIdx.x; b < SEL_MAX_BINS; b += SEL_THREADS) fhist[b] = 0u;
        const unsigned sub = f_sub;
        if (sub != ICPB_NONE) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned j = threadIdx.x + (unsigned)u * SEL_THREADS;
                if (j >= cnt || rsub[u] != sub) continue;
                const unsigned at = atomicAdd(&f_cnt, 1u);
                if (at < LIN_FINAL_CAP) {
                    f_key[at] = rk[u];
                    f_perm[at] = perm[ri[u]];
                }
            }
            for (unsigned j = threadIdx.x + 4u * SEL_THREADS; j < cnt; j += blockDim.x) {
                const unsigned long long key = __ldcg(cand_key + j);
                if (lin_sub_bin(__longlong_as_double((long long)key), scale, bin) != sub) continue;
                const unsigned at = atomicAdd(&f_cnt, 1u);
                if (at < LIN_FINAL_CAP) {
                    f_key[at] = key;
                    f_perm[at] = perm[__ldcg(cand_idx + j)];
                }
            }
        }
        __syncthreads();
        const unsigned m = f_cnt;
        if (sub != ICPB_NONE && m <= LIN_FINAL_CAP) {
            const unsigned long long want = f_krem - 1ull; // rank (0-based) inside the sub-bin
            for (unsigned t = threadIdx.x; t < m; t += blockDim.x) {
                const unsigned long long kt = f_key[t];
                const uint32_t pt = f_perm[t];
                unsigned rank = 0;
                for (unsigned j = 0; j < m; ++j) rank += (f_key[j] < kt || (f_key[j] == kt && f_perm[j] < pt)) ? 1u : 0u;
                if ((unsigned long long)rank == want) {
                    ls->tau = __longlong_as_double((long long)kt);
                    ls->p_cut = (long long)pt;
                    ls->need_tie = 0;
                    ls->sel_krem = 1;
                }
            }
            done_fast = true;
        }
        __syncthreads();
    }
    // nothing to select (tau = NaN), or too many pairs in one sub-bin (masses of equal distances): the general radix
    // selection over all collected keys, then the tie-break on the pair index
    if (!done_fast) {
    for (int pass = 0; pass < SEL_D_PASSES; ++pass) {
        const int bits = c_sel_d_bits[pass], shift = c_sel_d_shift[pass], hsp = shift + bits;
        const unsigned bin_mask = (1u << bits) - 1u;
        const unsigned long long prefix = pass == 0 ? 0ull : ls->sel_prefix;
        if (pass > 0 && ls->sel_krem == 0) break;
        for (unsigned j = threadIdx.x; j < cnt; j += blockDim.x) {
            const unsigned long long key = __ldcg(cand_key + j);
            if (pass == 0 || (key >> hsp) == (prefix >> hsp)) atomicAdd(&fhist[(unsigned)(key >> shift) & bin_mask], 1u);
        }
        __syncthreads();
        select_pick(ls, fhist, pass, false, scan_smem, nullptr, 1, 0, bin != ICPB_NONE ? s_krem : 0ull);
        __syncthreads();
    }
    if (ls->need_tie) { // which of the equidistant pairs survive: the lowest pair indices
        const unsigned long long tau_bits = (unsigned long long)__double_as_longlong(ls->tau);
        for (int pass = 0; pass < SEL_T_PASSES; ++pass) {
            const int bits = c_sel_t_bits[pass], shift = c_sel_t_shift[pass], hst = shift + bits;
            const unsigned bin_mask = (1u << bits) - 1u;
            const unsigned prefix = pass == 0 ? 0u : ls->tie_prefix;
            for (unsigned j = threadIdx.x; j < cnt; j += blockDim.x) {
                if (__ldcg(cand_key + j) != tau_bits) continue;
                const uint32_t pi = perm[__ldcg(cand_idx + j)];
                if (pass == 0 || hst >= 32 || (pi >> hst) == (prefix >> hst)) atomicAdd(&fhist[(pi >> shift) & bin_mask], 1u);
            }
            __syncthreads();
            select_pick(ls, fhist, pass, true, scan_smem);
            __syncthreads();
        }
    }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        ls->tickets[7] = 0;
        if (multi) ls->lin_epoch = ls->lin_epoch + 1u;
        ICPB_STAMP(true, 4);
    }
    if (!fuse) { // stand-alone trim: the state goes back, moments_kernel follows as a launch of its own
        __syncthreads();
        uint32_t *dst = reinterpret_cast<uint32_t *>(st);
        const uint32_t *src = reinterpret_cast<const uint32_t *>(ls);
        for (int w = threadIdx.x; w < MOM_STATE_WORDS; w += MOM_THREADS) dst[w] = src[w];
        return;
    }
    // the pairs of the bin that are kept, in the list's order (deterministic); the first 1024 come from the registers
    const double tau = ls->tau;
    const long long p_cut = ls->p_cut;
#pragma unroll
    for (int q = 0; q < ICPB_NSUMS; ++q) acc[q] = 0.0;
    if (bin != ICPB_NONE && k > 0) {
        for (unsigned j0 = 0; j0 < cnt; j0 += 4u * SEL_THREADS) {
            double d[4], vx[4], vy[4], vz[4];
            uint32_t pos[4], pi[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned j = j0 + (unsigned)u * SEL_THREADS + threadIdx.x;
                const bool in = j < cnt;
                const uint32_t i = j0 == 0 ? ri[u] : (in ? __ldcg(cand_idx + j) : 0u);
                d[u] = !in ? icpb_inf() : __longlong_as_double((long long)(j0 == 0 ? rk[u] : __ldcg(cand_key + j)));
                const bool want = d[u] <= tau; // most of the bin's pairs on one side of tau need nothing more
                vx[u] = want ? sx[i] : 0.0;
                vy[u] = want ? sy[i] : 0.0;
                vz[u] = want ? sz[i] : 0.0;
                pos[u] = want ? nn_pos[i] : ICPB_NONE;
                pi[u] = want ? perm[i] : 0u;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (d[u] <= tau && pair_kept(d[u], pi[u], tau, p_cut)) moments_add(acc, M, vx[u], vy[u], vz[u], g.pts[pos[u]], d[u]);
        }
    }
    __syncthreads();
    moments_block_partial(acc, ms.extra, ms);
    ICPB_STAMP(true, 6);
    moments_last_finish(st, rowsum, true, trace, trace_cap, 1, p2p, cond, use_cond, ms);
}

__global__ void __launch_bounds__(MOM_THREADS, 2)
    moments_kernel(const GridView g, const double *__restrict__ sx, const double *__restrict__ sy,
                   const double *__restrict__ sz, uint32_t n, IcpState *st,
                   const uint32_t *__restrict__ nn_pos, const double *__restrict__ nn_d,
                   const uint32_t *__restrict__ perm, double *__restrict__ partials, double *__restrict__ trace,
                   int trace_cap, int fused_pose, const P2PView p2p, cudaGraphConditionalHandle cond, int use_cond)
{
    if (st->done) return;
    __shared__ MomShared ms;
    double acc[ICPB_NSUMS];
#pragma unroll
    for (int k = 0; k < ICPB_NSUMS; ++k) acc[k] = 0.0;
    const double tau = st->tau;
    const long long p_cut = st->p_cut;
    const bool any = st->sel_k > 0;
    __shared__ double M[12]; // broadcast reads instead of 24 registers per thread
    if (threadIdx.x < 12) M[threadIdx.x] = st->M[threadIdx.x];
    __syncthreads();
    if (any) {
        // four pairs per trip: their (coalesced) loads are all in flight before the first dependent gather of a
        // model point, and few, long-lived blocks amortise the fixed cost of a block (state reads, reduction, ticket)
        const uint32_t T = gridDim.x * blockDim.x;
        for (uint32_t base = blockIdx.x * blockDim.x + threadIdx.x; base < n; base += 4u * T) {
            double d[4], vx[4], vy[4], vz[4];
            uint32_t pos[4];
            bool in[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t j = base + (uint32_t)u * T;
                in[u] = j < n && j >= base;
                d[u] = in[u] ? nn_d[j] : icpb_inf();
                vx[u] = in[u] ? sx[j] : 0.0;
                vy[u] = in[u] ? sy[j] : 0.0;
                vz[u] = in[u] ? sz[j] : 0.0;
                pos[u] = in[u] ? nn_pos[j] : ICPB_NONE;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (!(d[u] <= tau)) continue; // also +inf (no pair) and the out-of-range slots
                if (d[u] == tau && !pair_kept(d[u], perm[base + (uint32_t)u * T], tau, p_cut)) continue;
                moments_add(acc, M, vx[u], vy[u], vz[u], g.pts[pos[u]], d[u]);
            }
        }
    }
    moments_block_partial(acc, partials + (size_t)blockIdx.x * ICPB_NSUMS, ms);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) ms.is_last = (atomicAdd(&st->tickets[15], 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!ms.is_last) return;
    __threadfence();
    double v[3];
    moments_last_begin(st, partials, gridDim.x, v, ms);
    moments_last_finish(st, v, false, trace, trace_cap, fused_pose, p2p, cond, use_cond, ms);
}

// stand-alone pose step (multi-GPU: after the all-reduce of the 17 sums)
__global__ void pose_kernel(IcpState *st, double *trace, int trace_cap)
{
    if (st->done) return;
    const int row = st->executed;
    icpb_iteration_finish(*st, row < trace_cap ? trace + (size_t)row * ICPB_TRACE_COLS_DEV : nullptr);
}

__global__ void init_state_kernel(IcpState *st, Aff12 Cl2g, unsigned long long max_iter, double min_mse,
                                  double min_mse_diff, double overlap, unsigned long long n_po,
                                  unsigned long long n_local, unsigned long long n_offset,
                                  cudaGraphConditionalHandle cond, int use_cond, double lazy_margin, uint32_t *lhist)
{
    for (int b = threadIdx.x; b < LIN_BINS; b += blockDim.x) lhist[b * LIN_STRIDE] = 0u; // the fused trim histogram starts empty
    if (threadIdx.x != 0) return;
    icpb_state_init(*st, Cl2g.a, max_iter, min_mse, min_mse_diff, overlap, n_po, lazy_margin);
    st->n_local = n_local;
    st->n_global_offset = n_offset;
    st->sel_prefix = st->sel_krem = st->sel_k = 0;
    st->cnt_eq = st->quota = 0;
    st->tie_prefix = 0;
    st->tie_krem = 0;
    st->tie_local = 0;
    for (int k = 0; k < ICPB_NSUMS; ++k) st->sums[k] = 0.0;
    if (use_cond) cudaGraphSetConditional(cond, st->done ? 0u : 1u); // max_iter == 0 etc.: the body never runs
}

// ---------------------------------------------------------------------------
// K7 + debug: kept distances / correspondences of the last executed iteration
// ---------------------------------------------------------------------------
__global__ void compact_kept_kernel(const double *__restrict__ nn_d, const uint32_t *__restrict__ perm, uint32_t n,
                                    const IcpState *st,
                                    unsigned long long *__restrict__ out_keys, unsigned int *__restrict__ counter)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool keep = false;
    double d = 0.0;
    if (i < n && st->sel_k > 0) {
        d = nn_d[i];
        keep = pair_kept(d, perm[i], st->tau, st->p_cut);
    }
    const unsigned m = __ballot_sync(0xFFFFFFFFu, keep);
    if (!m) return;
    const int lane = threadIdx.x & 31;
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(counter, (unsigned)__popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (keep) out_keys[base + __popc(m & ((1u << lane) - 1u))] = (unsigned long long)__double_as_longlong(d);
}

__global__ void debug_pairs_kernel(const GridView g, const double *__restrict__ sx, const double *__restrict__ sy,
                                   const double *__restrict__ sz, const uint32_t *__restrict__ nn_pos,
                                   const double *__restrict__ nn_d, const uint32_t *__restrict__ perm, uint32_t n,
                                   const IcpState *st, uint32_t *__restrict__ idx_out,
                                   double *__restrict__ d_out, uint8_t *__restrict__ kept_out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t pos = nn_pos[i];
    const uint32_t o = perm[i]; // back to the adapter's visiting order
    const double d = nn_d[i];
    // nn_pos keeps the nearest neighbour as such; it is a pair only when its distance is finite (within d_box, filter
    // passed).  A finite distance that is not the distance to nn_pos is the stand-in of a deferred query (lazy search).
    uint32_t idx = (pos == ICPB_NONE || !(d < icpb_inf())) ? ICPB_NONE : (uint32_t)g.pts[pos].idx;
    if (idx != ICPB_NONE) {
        const double *Ml = st->early ? st->M : st->Mprev; // the transform the last search used
        double qx, qy, qz;
        icpb_aff_apply(Ml, sx[i], sy[i], sz[i], qx, qy, qz);
        const MPoint m = g.pts[pos];
        const double dx = qx - m.x, dy = qy - m.y, dz = qz - m.z;
        if (sqrt((dx * dx + dy * dy) + dz * dz) != d) idx = 0xFFFFFFFEu; // ICPB_DEFERRED
    }
    idx_out[o] = idx;
    d_out[o] = d;
    kept_out[o] = (st->sel_k > 0 && pair_kept(d, o, st->tau, st->p_cut)) ? 1 : 0;
}

// ---------------------------------------------------------------------------
// source upload: spatial order (Hilbert curve over 1024^3 cells of the source's own bounding box; Morton as an option)
// so that a warp's 32 queries are neighbours.  Measured on cfg2: Morton 10 bits per axis 17.98 ms of search per align,
// 8 bits 18.57, 7 bits 20.04; Hilbert 10 bits 16.70 -- the order of the queries is worth more than most kernel tweaks
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t spread10(uint32_t v)
{
    v &= 1023u;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
// second stage of the bounding box, on the device: out[0..2] = min corner, out[3] = cells per axis / largest extent (the
// scale of the Morton keys) -- the upload path needs no trip to the host for it
__global__ void __launch_bounds__(256) bbox_final_kernel(const double *__restrict__ part, unsigned nb, double *__restrict__ out, double cells_per_axis)
{
    __shared__ double sm[8][6];
    double mn[3] = {icpb_inf(), icpb_inf(), icpb_inf()}, mx[3] = {-icpb_inf(), -icpb_inf(), -icpb_inf()};
    for (unsigned b = threadIdx.x; b < nb; b += blockDim.x)
        for (int a = 0; a < 3; ++a) {
            mn[a] = fmin(mn[a], part[(size_t)b * 6 + a]);
            mx[a] = fmax(mx[a], part[(size_t)b * 6 + 3 + a]);
        }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int a = 0; a < 3; ++a)
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = fmin(mn[a], __shfl_xor_sync(0xFFFFFFFFu, mn[a], o));
            mx[a] = fmax(mx[a], __shfl_xor_sync(0xFFFFFFFFu, mx[a], o));
        }
    if (lane == 0)
        for (int a = 0; a < 3; ++a) {
            sm[warp][a] = mn[a];
            sm[warp][3 + a] = mx[a];
        }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w)
            for (int a = 0; a < 3; ++a) {
                sm[0][a] = fmin(sm[0][a], sm[w][a]);
                sm[0][3 + a] = fmax(sm[0][3 + a], sm[w][3 + a]);
            }
        double ext = fmax(fmax(sm[0][3] - sm[0][0], sm[0][4] - sm[0][1]), sm[0][5] - sm[0][2]);
        if (!(ext > 0.0) || !(ext < icpb_inf())) ext = 1.0;
        out[0] = sm[0][0];
        out[1] = sm[0][1];
        out[2] = sm[0][2];
        out[3] = (cells_per_axis - 0.001) / ext;
        out[4] = cells_per_axis - 1.0;
    }
}
// K0: the vertices PointcloudAdapter visits at this density (icp/icp.hpp:126-138), gathered in visiting order; the
// visiting order in closed form: density.h
__global__ void density_gather_kernel(const double *__restrict__ full, const DensitySeg *__restrict__ segs, int nseg,
                                      unsigned long long m, double *__restrict__ out)
{
    __shared__ DensitySeg ss[64];
    const int ns = nseg < 64 ? nseg : 64;
    for (int i = threadIdx.x; i < ns; i += blockDim.x) ss[i] = segs[i];
    __syncthreads();
    const unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    const int si = icpb_density_find(nseg <= 64 ? ss : segs, nseg, k);
    const unsigned long long i = icpb_density_index(nseg <= 64 ? ss[si] : segs[si], k);
    out[3 * k] = full[3 * i];
    out[3 * k + 1] = full[3 * i + 1];
    out[3 * k + 2] = full[3 * i + 2];
}
// the same visiting order for a parallel per-vertex array of `width` 4-byte items
__global__ void density_gather_u32_kernel(const uint32_t *__restrict__ full, const DensitySeg *__restrict__ segs, int nseg,
                                          unsigned long long m, int width, uint32_t *__restrict__ out)
{
    const unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    const unsigned long long i = icpb_density_index(segs[icpb_density_find(segs, nseg, k)], k);
    for (int w = 0; w < width; ++w) out[(size_t)width * k + w] = full[(size_t)width * i + w];
}
// Hilbert index of a cell (x, y, z < 2^bits) -- Skilling's transpose form, then the bits interleaved: consecutive keys
// are face-adjacent cells (a Morton curve jumps at every power-of-two boundary)
__device__ __forceinline__ uint32_t hilbert3(uint32_t x, uint32_t y, uint32_t z, int bits)
{
    uint32_t X[3] = {x, y, z};
    const uint32_t M = 1u << (bits - 1);
    for (uint32_t Q = M; Q > 1u; Q >>= 1) {
        const uint32_t P = Q - 1u;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (X[i] & Q)
                X[0] ^= P;
            else {
                const uint32_t t = (X[0] ^ X[i]) & P;
                X[0] ^= t;
                X[i] ^= t;
            }
        }
    }
    X[1] ^= X[0];
    X[2] ^= X[1];
    uint32_t t = 0;
    for (uint32_t Q = M; Q > 1u; Q >>= 1)
        if (X[2] & Q) t ^= Q - 1u;
    X[0] ^= t, X[1] ^= t, X[2] ^= t;
    return (spread10(X[0]) << 2) | (spread10(X[1]) << 1) | spread10(X[2]);
}
__global__ void morton_keys_kernel(const double *__restrict__ aos, uint32_t n, const double *__restrict__ bb,
                                   uint32_t *__restrict__ keys, uint32_t *__restrict__ vals, int hilbert_bits = 0)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double ox = bb[0], oy = bb[1], oz = bb[2], inv = bb[3], top = bb[4];
    const double fx = fmin(fmax((aos[3 * (size_t)i] - ox) * inv, 0.0), top);
    const double fy = fmin(fmax((aos[3 * (size_t)i + 1] - oy) * inv, 0.0), top);
    const double fz = fmin(fmax((aos[3 * (size_t)i + 2] - oz) * inv, 0.0), top);
    keys[i] = hilbert_bits > 0 ? hilbert3((uint32_t)fx, (uint32_t)fy, (uint32_t)fz, hilbert_bits)
                               : (spread10((uint32_t)fx) | (spread10((uint32_t)fy) << 1) | (spread10((uint32_t)fz) << 2));
    vals[i] = i;
}
__global__ void gather_soa_kernel(const double *__restrict__ aos, const uint32_t *__restrict__ perm, uint32_t n,
                                  double *__restrict__ x, double *__restrict__ y, double *__restrict__ z)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const size_t s = perm[j];
    x[j] = aos[3 * s];
    y[j] = aos[3 * s + 1];
    z[j] = aos[3 * s + 2];
}
// stand-alone Pairs::getTransform entry: model points of the pairs as an MPoint array (pair i <-> position i)
__global__ void pack_points_kernel(const double *__restrict__ aos, uint32_t n, MPoint *__restrict__ out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    MPoint m;
    m.x = aos[3 * (size_t)i];
    m.y = aos[3 * (size_t)i + 1];
    m.z = aos[3 * (size_t)i + 2];
    m.idx = i;
    out[i] = m;
}
__global__ void iota_kernel(uint32_t *__restrict__ v, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = i;
}

} // namespace icpb

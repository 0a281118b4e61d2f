// grid_view.h -- the model index that replaces libkdtree++ on the B200 path,
// and the exact nearest-neighbour search over it.
//
// Replaces: FindPairsKDTree::findPairs -> kdtree.find_nearest(node, d_box)
//           (reference icp/icp.hpp:242-252).
//
// Layout (all in HBM, built once per model by grid_build.cuh):
//   pts[n]          model points sorted by level-0 cell key (x fastest), 32 B
//                   each: fp64 x,y,z in the global frame + insertion index
//   cell_start[c]   exclusive prefix of points per level-0 cell, ncell+1 entries
//                   (so x-adjacent cells form one contiguous run of pts)
//   mask[l][c]      l >= 1: one byte per level-l cell, bit (dx|dy<<1|dz<<2) set
//                   when that child cell of level l-1 holds at least one point.
//                   Level l cells have edge h*2^l; the top level is one cell.
//
// Search: exact Euclidean NN inside the inclusive ball d_box.  The uniform grid
// is searched as "rings" of cells around the query, but rings are expanded
// hierarchically: the start level is the one whose cell edge covers the current
// upper bound U (previous iteration's pair, or d_box), so the ball touches at
// most 2x2x2 cells of that level; those cells are then descended depth-first,
// nearest child first, skipping empty space through the occupancy masks and
// pruning every cell whose box is farther than the best distance so far.
// Pruning is conservative (eps slack, ">" not ">=") so the answer is the
// lexicographic minimum over (d^2, insertion index) of ALL model points: the
// same point a kd-tree returns, with equidistant ties resolved to the lowest
// insertion index.  Distances are evaluated exactly like libkdtree++ does:
// sqrt((dx*dx + dy*dy) + dz*dz) in fp64 without FMA contraction.
//
// The functions are __host__ __device__ so tests/ can run the very same search
// code on the CPU (tests/harness) against the oracle before it meets a GPU.
#pragma once
#include <stdint.h>
#include <math.h>

#ifdef __CUDACC__
#define ICPB_HD __host__ __device__ __forceinline__
#else
#define ICPB_HD inline
#endif

#define ICPB_MAX_LEVELS 16
#define ICPB_NONE 0xFFFFFFFFu
#define ICPB_SEARCH_STACK 112
#ifndef ICPB_COUNT
#define ICPB_COUNT(what) // instrumentation hook (tests/harness defines it to count traversal work)
#endif

struct
#ifdef __CUDACC__
    __align__(32)
#else
    __attribute__((aligned(32)))
#endif
        MPoint {
    double x, y, z;
    unsigned long long idx; // insertion index (order of addToModel visits)
};

// fp32 copy of a model point in cell units relative to the grid origin: the candidate pre-filter reads these
// (16 B, coalesced float4 loads); only candidates that can still win are re-evaluated exactly from pts[].
struct
#ifdef __CUDACC__
    __align__(16)
#else
    __attribute__((aligned(16)))
#endif
        FPoint {
    float x, y, z, w;
};

struct GridView {
    double o[3];  // origin (min corner of the model bounding box)
    double h;     // level-0 cell edge
    double inv_h; // 1/h
    double eps;   // geometric slack for pruning (1e-9 * h)
    int dim[ICPB_MAX_LEVELS][3];
    int L; // top level: dim[L] == {1,1,1}
    const uint32_t *cell_start;
    const uint8_t *mask[ICPB_MAX_LEVELS];
    const MPoint *pts;
    const FPoint *ptsf; // same order as pts
    uint32_t n;
};

ICPB_HD double icpb_inf() { return __builtin_huge_val(); }

// level-0 cell coordinate of a model point along one axis
ICPB_HD int icpb_cell_coord(double p, double o, double inv_h, int dim)
{
    double t = floor((p - o) * inv_h);
    if (!(t > 0.0)) t = 0.0;
    if (t > (double)(dim - 1)) t = (double)(dim - 1);
    return (int)t;
}

// squared distance from q to the box of cell (cx,cy,cz) at `level`, shrunk by eps
ICPB_HD double icpb_cell_mindist2(const GridView &g, int level, int cx, int cy, int cz, double qx, double qy,
                                  double qz)
{
    const double hl = g.h * (double)(1u << level);
    double lo, gap, d2 = 0.0;
    lo = g.o[0] + (double)cx * hl;
    gap = fmax(fmax(lo - qx, qx - (lo + hl)) - g.eps, 0.0);
    d2 += gap * gap;
    lo = g.o[1] + (double)cy * hl;
    gap = fmax(fmax(lo - qy, qy - (lo + hl)) - g.eps, 0.0);
    d2 += gap * gap;
    lo = g.o[2] + (double)cz * hl;
    gap = fmax(fmax(lo - qz, qz - (lo + hl)) - g.eps, 0.0);
    d2 += gap * gap;
    return d2;
}

// The same lower bound in fp32 "cell units" (u = (q - o)/h), used by the traversal: cell corners are small
// integers (exact in fp32), the query coordinate carries a conversion error that `slack` (>= that error plus
// the cell-assignment slack, in cells) absorbs, and the sum is scaled down to stay below the true value.
// Conservative by construction: never larger than the true squared distance (in cells^2) to any point filed
// under the cell, so pruning with it cannot lose a candidate.
ICPB_HD float icpb_cell_mindist2f(int level, int cx, int cy, int cz, float ux, float uy, float uz, float slack)
{
    const float s = (float)(1u << level);
    float lo, gap, d2;
    lo = (float)cx * s;
    gap = fmaxf(fmaxf(lo - ux, ux - (lo + s)) - slack, 0.0f);
    d2 = gap * gap;
    lo = (float)cy * s;
    gap = fmaxf(fmaxf(lo - uy, uy - (lo + s)) - slack, 0.0f);
    d2 += gap * gap;
    lo = (float)cz * s;
    gap = fmaxf(fmaxf(lo - uz, uz - (lo + s)) - slack, 0.0f);
    d2 += gap * gap;
    return d2 * 0.999999f;
}
// upper bound of (bound * inv_h^2) in fp32
ICPB_HD float icpb_bound_cells(double bound, double inv_h2)
{
    const double b = bound * inv_h2;
    if (!(b < 3.0e38)) return (float)icpb_inf();
    return (float)b * 1.000001f + 1e-30f;
}

ICPB_HD void icpb_nn_eval(const MPoint &m, uint32_t pos, double qx, double qy, double qz, double &bd2,
                          unsigned long long &bidx, uint32_t &bpos)
{
    const double dx = qx - m.x, dy = qy - m.y, dz = qz - m.z;
    const double d2 = (dx * dx + dy * dy) + dz * dz;
    if (d2 < bd2 || (d2 == bd2 && m.idx < bidx)) {
        bd2 = d2;
        bidx = m.idx;
        bpos = pos;
    }
}

// squared fp32 radius (cell units) inside which a candidate must be re-evaluated exactly: the fp32 distance of a
// point that can still beat (or tie) the best is at most sqrt(bound) + (conversion errors of query and point)
ICPB_HD float icpb_cand_radius2(float boundf, float slack)
{
    const float r = sqrtf(boundf) + 2.0f * slack;
    return r * r * 1.00001f;
}

ICPB_HD void icpb_scan_cell(const GridView &g, int cx, int cy, int cz, double qx, double qy, double qz, float ux, float uy,
                            float uz, float slack, double dbox2s, double inv_h2, double &bd2, unsigned long long &bidx,
                            uint32_t &bpos, float &boundf, float &candf)
{
    const uint32_t c = ((uint32_t)cz * (uint32_t)g.dim[0][1] + (uint32_t)cy) * (uint32_t)g.dim[0][0] + (uint32_t)cx;
    const uint32_t s = g.cell_start[c], e = g.cell_start[c + 1];
    ICPB_COUNT(cells);
    // four candidates per step: the loads are issued together (memory-level parallelism), the tail repeats the last one
    for (uint32_t p = s; p < e; p += 4) {
        const uint32_t last = e - 1;
        const uint32_t p1 = p + 1 < e ? p + 1 : last, p2 = p + 2 < e ? p + 2 : last, p3 = p + 3 < e ? p + 3 : last;
        const FPoint f0 = g.ptsf[p], f1 = g.ptsf[p1], f2 = g.ptsf[p2], f3 = g.ptsf[p3];
        float dx, dy, dz;
        dx = ux - f0.x, dy = uy - f0.y, dz = uz - f0.z;
        const float d0 = dx * dx + dy * dy + dz * dz;
        dx = ux - f1.x, dy = uy - f1.y, dz = uz - f1.z;
        const float d1 = dx * dx + dy * dy + dz * dz;
        dx = ux - f2.x, dy = uy - f2.y, dz = uz - f2.z;
        const float d2 = dx * dx + dy * dy + dz * dz;
        dx = ux - f3.x, dy = uy - f3.y, dz = uz - f3.z;
        const float d3 = dx * dx + dy * dy + dz * dz;
        if (fminf(fminf(d0, d1), fminf(d2, d3)) > candf) continue; // none can be nearer than the best so far (nor tie)
#ifdef __CUDACC__
#pragma unroll
#endif
        for (int k = 0; k < 4; ++k) {
            const float dk = k == 0 ? d0 : (k == 1 ? d1 : (k == 2 ? d2 : d3));
            const uint32_t pk = k == 0 ? p : (k == 1 ? p1 : (k == 2 ? p2 : p3));
            if (dk > candf || (k > 0 && pk != p + (uint32_t)k)) continue; // filtered, or a repeated tail slot
            ICPB_COUNT(exact);
            const double before = bd2;
            icpb_nn_eval(g.pts[pk], pk, qx, qy, qz, bd2, bidx, bpos);
            if (bd2 < before) {
                boundf = icpb_bound_cells((dbox2s < bd2) ? dbox2s : bd2, inv_h2);
                candf = icpb_cand_radius2(boundf, slack);
            }
        }
    }
}

// Exact NN of q.  dbox2s = d_box^2 widened by a few ulps (+inf when unbounded);
// warm_pos = position in pts[] of a candidate to seed the bound (ICPB_NONE for none).
// Returns the position in pts[] (ICPB_NONE if the model is empty or nothing lies
// within the bound) and the exact squared distance.  *visited counts cells scanned.
ICPB_HD void icpb_nn_search(const GridView &g, double qx, double qy, double qz, double dbox2s, uint32_t warm_pos,
                            uint32_t &out_pos, double &out_d2, int &start_level)
{
    const double INF = icpb_inf();
    double bd2 = INF;
    unsigned long long bidx = ~0ull;
    uint32_t bpos = ICPB_NONE;
    start_level = 0;
    out_pos = ICPB_NONE;
    out_d2 = INF;
    if (g.n == 0) return;
    if (warm_pos < g.n) icpb_nn_eval(g.pts[warm_pos], warm_pos, qx, qy, qz, bd2, bidx, bpos);

    double bound = (dbox2s < bd2) ? dbox2s : bd2;
    // start level: smallest l with h*2^l >= 2U, so the ball spans <= 2 cells per axis
    int l = 0;
    double U = INF;
    if (bound < INF) {
        U = sqrt(bound) * (1.0 + 1e-15) + g.eps;
        const double ratio = 2.0 * U * g.inv_h;
        while (l < g.L && (double)(1u << l) < ratio) ++l;
    } else {
        l = g.L;
    }
    start_level = l;

    // traversal state in fp32 cell units
    const double inv_h2 = g.inv_h * g.inv_h;
    const double uxd = (qx - g.o[0]) * g.inv_h, uyd = (qy - g.o[1]) * g.inv_h, uzd = (qz - g.o[2]) * g.inv_h;
    const float ux = (float)uxd, uy = (float)uyd, uz = (float)uzd;
    const float slack = 1e-6f * (fmaxf(fmaxf(fabsf(ux), fabsf(uy)), fabsf(uz)) + 4096.0f);
    float boundf = icpb_bound_cells(bound, inv_h2);
    float candf = icpb_cand_radius2(boundf, slack);

    unsigned long long stack[ICPB_SEARCH_STACK];
    int sp = 0;
    {
        const double inv_hl = g.inv_h / (double)(1u << l);
        int lo[3], hi[3];
        const double q[3] = {qx, qy, qz};
        for (int a = 0; a < 3; ++a) {
            const double dm1 = (double)(g.dim[l][a] - 1);
            double flo = floor((q[a] - U - g.o[a]) * inv_hl);
            double fhi = floor((q[a] + U - g.o[a]) * inv_hl);
            if (!(flo > 0.0)) flo = 0.0; // also catches -inf / NaN
            if (!(fhi < dm1)) fhi = dm1;
            if (flo > dm1 || fhi < 0.0) { // ball misses the grid on this axis
                out_pos = bpos;
                out_d2 = bd2;
                return;
            }
            lo[a] = (int)flo;
            hi[a] = (int)fhi;
        }
        float best_md2 = (float)INF;
        int best_at = -1;
        for (int z = lo[2]; z <= hi[2]; ++z)
            for (int y = lo[1]; y <= hi[1]; ++y)
                for (int x = lo[0]; x <= hi[0]; ++x) {
                    const float md2 = icpb_cell_mindist2f(l, x, y, z, ux, uy, uz, slack);
                    if (md2 > boundf) continue;
                    if (sp >= ICPB_SEARCH_STACK - 64) continue; // cannot happen (<= 27 roots)
                    if (md2 < best_md2) {
                        best_md2 = md2;
                        best_at = sp;
                    }
                    stack[sp++] = ((unsigned long long)l << 60) | ((unsigned long long)z << 40) |
                                  ((unsigned long long)y << 20) | (unsigned long long)x;
                }
        if (best_at >= 0) { // nearest root on top
            const unsigned long long t = stack[best_at];
            stack[best_at] = stack[sp - 1];
            stack[sp - 1] = t;
        }
    }

    // Two-phase ("while-while") traversal: every lane first walks the pyramid until it holds a cell group to scan
    // (a level-1 node, or a level-0 root cell), then scans it -- lanes of a warp stay in the same code section
    // instead of interleaving node expansion with candidate evaluation.
    const unsigned ORD = 0x76534210u; // child visiting order as XOR codes: 0,1,2,4,3,5,6,7
    for (;;) {
        int have = 0, lx = 0, ly = 0, lz = 0;
        unsigned lmask = 0;
        while (sp > 0) {
            const unsigned long long e = stack[--sp];
            ICPB_COUNT(pops);
            const int lev = (int)(e >> 60);
            const int cx = (int)(e & 0xFFFFFu), cy = (int)((e >> 20) & 0xFFFFFu), cz = (int)((e >> 40) & 0xFFFFFu);
            if (icpb_cell_mindist2f(lev, cx, cy, cz, ux, uy, uz, slack) > boundf) continue;
            if (lev == 0) {
                have = 1;
                lx = cx, ly = cy, lz = cz;
                break;
            }
            const size_t ci = ((size_t)cz * g.dim[lev][1] + cy) * g.dim[lev][0] + cx;
            const unsigned m = g.mask[lev][ci];
            ICPB_COUNT(nodes);
            if (!m) continue;
            if (lev == 1) {
                have = 2;
                lmask = m;
                lx = cx, ly = cy, lz = cz;
                break;
            }
            const float hc = (float)(1u << (lev - 1));
            const unsigned near = (ux >= (float)(2 * cx + 1) * hc ? 1u : 0u) | (uy >= (float)(2 * cy + 1) * hc ? 2u : 0u) |
                                  (uz >= (float)(2 * cz + 1) * hc ? 4u : 0u);
            for (int k = 7; k >= 0; --k) {
                const unsigned ch = near ^ ((ORD >> (4 * k)) & 7u);
                if (!((m >> ch) & 1u)) continue;
                const int x = 2 * cx + (int)(ch & 1u), y = 2 * cy + (int)((ch >> 1) & 1u),
                          z = 2 * cz + (int)((ch >> 2) & 1u);
                if (icpb_cell_mindist2f(lev - 1, x, y, z, ux, uy, uz, slack) > boundf) continue;
                if (sp < ICPB_SEARCH_STACK)
                    stack[sp++] = ((unsigned long long)(lev - 1) << 60) | ((unsigned long long)z << 40) |
                                  ((unsigned long long)y << 20) | (unsigned long long)x;
            }
        }
        if (!have) break;
        if (have == 1) {
            icpb_scan_cell(g, lx, ly, lz, qx, qy, qz, ux, uy, uz, slack, dbox2s, inv_h2, bd2, bidx, bpos, boundf, candf);
        } else {
            const unsigned near = (ux >= (float)(2 * lx + 1) ? 1u : 0u) | (uy >= (float)(2 * ly + 1) ? 2u : 0u) |
                                  (uz >= (float)(2 * lz + 1) ? 4u : 0u);
            for (int k = 0; k < 8; ++k) {
                const unsigned ch = near ^ ((ORD >> (4 * k)) & 7u);
                if (!((lmask >> ch) & 1u)) continue;
                const int x = 2 * lx + (int)(ch & 1u), y = 2 * ly + (int)((ch >> 1) & 1u),
                          z = 2 * lz + (int)((ch >> 2) & 1u);
                if (icpb_cell_mindist2f(0, x, y, z, ux, uy, uz, slack) > boundf) continue;
                icpb_scan_cell(g, x, y, z, qx, qy, qz, ux, uy, uz, slack, dbox2s, inv_h2, bd2, bidx, bpos, boundf, candf);
            }
        }
    }
    out_pos = bpos;
    out_d2 = bd2;
}

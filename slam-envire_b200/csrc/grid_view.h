// grid_view.h -- the model index that replaces libkdtree++ on the B200 path,
// and the exact nearest-neighbour search over it.
//
// Replaces: FindPairsKDTree::findPairs -> kdtree.find_nearest(node, d_box)
//           (reference icp/icp.hpp:242-252).
//
// Layout (all in HBM, built once per model by the K1 kernels):
//   pts[n]          model points sorted by level-0 cell key (x fastest), 32 B
//                   each: fp64 x,y,z in the global frame + insertion index
//   ptsf[n]         the same points as float4 in cell units relative to the grid origin (candidate pre-filter);
//                   both arrays are followed by ICPB_PAD sentinel entries (far away, insertion index ~0) so that
//                   the scanners can always read four candidates
//   cell_start[c]   exclusive prefix of points per level-0 cell, ncell+1 entries
//                   (so x-adjacent cells form one contiguous run of pts)
//   cell_box[c]     tight bounding box of the points of level-0 cell c, six nibbles like bits 8-31 of a node word
//                   (sixteenths of h); read by the pyramid walk before it scans a cell
//   node[l][c]      l >= 1: one 32-bit word per level-l cell.  Bits 0-7: child occupancy, bit (dx|dy<<1|dz<<2)
//                   set when that child cell of level l-1 holds at least one point.  Bits 8-31: the TIGHT
//                   bounding box of the points below the node, six nibbles lo x,y,z / hi x,y,z in sixteenths
//                   of the node's edge (lo rounded down, hi up).  Scenes are surfaces: a node's points fill a
//                   thin slab of its cube, and a query several cells off the surface prunes on the slab where
//                   the cube alone would make it descend.  Level l cells have edge h*2^l; the top level is one
//                   cell.
//
// Search: exact Euclidean NN inside the inclusive ball d_box, one thread per query, no stack in memory.
//   * The upper bound U is the previous iteration's pair (warm start) or d_box.  Everything that steers the walk
//     is fp32 in cell units with a slack that keeps every bound provably <= the true distance; fp64 is spent only
//     on the source transform, on the candidates that can still win, and on the final sqrt.
//   * 2U <= one cell edge (the converged regime, most iterations): the ball touches at most 2x2x2 level-0
//     cells; x-adjacent cells are one contiguous run of pts, so at most four runs are scanned, own column first.
//   * otherwise: the start level is the one whose cell edge covers 2U (<= 2x2x2 roots); each root is walked
//     depth-first, nearest child first, with the path kept as one byte of pending children per level in two
//     registers ("trail"), skipping empty space through the occupancy bits and pruning on the tight boxes.
// Pruning is conservative (slack, ">" not ">=") so the answer is the lexicographic minimum over
// (d^2, insertion index) of ALL model points: the same point a kd-tree returns, with equidistant ties resolved
// to the lowest insertion index.  Distances are evaluated exactly like libkdtree++ does:
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
#define ICPB_PAD 3 // sentinel entries behind pts[n-1] / ptsf[n-1]: the scanners read four candidates at a time
#ifndef ICPB_COUNT
#define ICPB_COUNT(what) // instrumentation hook (tests/harness defines it to count traversal work)
#endif
// fp32 bounds may use FMA (the library is built with -fmad=false for fp64 parity; the fp32 side only has to be
// conservative, not reproducible)
#ifdef __CUDA_ARCH__
#define ICPB_FMAF(a, b, c) __fmaf_rn((a), (b), (c))
#else
#define ICPB_FMAF(a, b, c) ((a) * (b) + (c))
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
    double eps;   // geometric slack of the cell assignment (1e-9 * h)
    int dim[ICPB_MAX_LEVELS][3];
    int L; // top level: dim[L] == {1,1,1}
    const uint32_t *cell_start;
    const uint32_t *cell_box; // level-0 tight boxes: six nibbles like a node word (bits 8-31), sixteenths of h
    const uint32_t *node[ICPB_MAX_LEVELS]; // see the layout comment
    const MPoint *pts;
    const FPoint *ptsf; // same order as pts
    uint32_t n;
    int block_levels; // searches that start at a level <= this scan their cell block flat (icpb_nn_block); -1: never
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

// ---- node words -----------------------------------------------------------------------------------------------
ICPB_HD uint32_t icpb_node_word(unsigned mask, const int lo[3], const int hi[3])
{
    return (uint32_t)mask | ((uint32_t)lo[0] << 8) | ((uint32_t)lo[1] << 12) | ((uint32_t)lo[2] << 16) |
           ((uint32_t)hi[0] << 20) | ((uint32_t)hi[1] << 24) | ((uint32_t)hi[2] << 28);
}
// nibble of a cell-unit coordinate u inside the level-`level` cube starting at cube_lo (sixteenths of the edge)
ICPB_HD int icpb_node_nibble(double u, double cube_lo, int level)
{
    double t = floor((u - cube_lo) * (16.0 / (double)(1u << level)));
    if (!(t > 0.0)) t = 0.0;
    if (t > 15.0) t = 15.0;
    return (int)t;
}
// the nibbles of a child's box seen from its parent (child code bit a selects the upper half on axis a)
ICPB_HD int icpb_nibble_up(uint32_t child_word, int shift, int upper_half) { return 8 * upper_half + (int)((child_word >> shift) & 15u) / 2; }

// ---- fp32 bounds in cell units --------------------------------------------------------------------------------
// Squared gap between the query coordinate u and the interval [lo, hi] along one axis.  u is the fp32 copy of the
// exact cell-unit coordinate, off by at most `slack` (which also absorbs the cell-assignment slack and the fp32
// rounding of the subtraction): the result never exceeds the true squared gap.
ICPB_HD float icpb_gap2f(float lo, float hi, float u, float slack)
{
    const float gap = fmaxf(fmaxf(lo - u, u - hi) - slack, 0.0f);
    return gap * gap;
}
// lower bound of the squared distance to the cube of cell (cx,cy,cz) at `level`: per axis the gap is
// |u - centre| - half edge (centre and half edge are exact in fp32), shrunk by slack like icpb_gap2f
ICPB_HD float icpb_cell_mindist2f(int level, int cx, int cy, int cz, float ux, float uy, float uz, float slack)
{
    const float s = (float)(1u << level), half = 0.5f * s, hs = half + slack;
    float gap, d2;
    gap = fmaxf(fabsf(ux - ICPB_FMAF((float)cx, s, half)) - hs, 0.0f);
    d2 = gap * gap;
    gap = fmaxf(fabsf(uy - ICPB_FMAF((float)cy, s, half)) - hs, 0.0f);
    d2 = ICPB_FMAF(gap, gap, d2);
    gap = fmaxf(fabsf(uz - ICPB_FMAF((float)cz, s, half)) - hs, 0.0f);
    d2 = ICPB_FMAF(gap, gap, d2);
    return d2 * 0.999999f;
}
// lower bound of the squared distance to the tight box stored in node word w of cell (cx,cy,cz) at `level` >= 1
ICPB_HD float icpb_node_mindist2f(uint32_t w, int level, int cx, int cy, int cz, float ux, float uy, float uz, float slack)
{
    const float s = (float)(1u << level), q = s * 0.0625f; // exact: small integers times powers of two
    float lo, d2;
    lo = (float)cx * s;
    d2 = icpb_gap2f(ICPB_FMAF((float)((w >> 8) & 15u), q, lo), ICPB_FMAF((float)(((w >> 20) & 15u) + 1u), q, lo), ux, slack);
    lo = (float)cy * s;
    d2 += icpb_gap2f(ICPB_FMAF((float)((w >> 12) & 15u), q, lo), ICPB_FMAF((float)(((w >> 24) & 15u) + 1u), q, lo), uy, slack);
    lo = (float)cz * s;
    d2 += icpb_gap2f(ICPB_FMAF((float)((w >> 16) & 15u), q, lo), ICPB_FMAF((float)(((w >> 28) & 15u) + 1u), q, lo), uz, slack);
    return d2 * 0.999999f;
}
// upper bound of (bound * inv_h^2) in fp32
ICPB_HD float icpb_bound_cells(double bound, double inv_h2)
{
    const double b = bound * inv_h2;
    if (!(b < 3.0e38)) return (float)icpb_inf();
    return (float)b * 1.000001f + 1e-30f;
}
// squared fp32 radius (cell units) inside which a candidate must be re-evaluated exactly: the fp32 distance of a
// point that can still beat (or tie) the best is at most sqrt(bound) + (conversion errors of query and point)
ICPB_HD float icpb_cand_radius2(float boundf, float slack)
{
    const float r = sqrtf(boundf) + 2.0f * slack;
    return r * r * 1.00001f;
}

// ---- search state of one query ----------------------------------------------------------------------------------
struct NNState {
    double qx, qy, qz;  // query, global frame
    double dbox2s;      // d_box^2 widened by a few ulps (+inf when unbounded)
    double inv_h2;
    double bd2;         // best exact squared distance so far
    double sd2;         // smallest exact squared distance of an evaluated point that is NOT the best (runner-up)
    double rs2;         // the search covers at least this squared radius (0: only what the answer needs), see below
    unsigned long long bidx;
    uint32_t bpos;
    float ux, uy, uz;   // query in cell units (fp32)
    float slack;
    float boundf;       // PRUNING radius^2 in cell units, rounded up: max(min(d_box^2, bd2), rs2)
    float candf;        // fp32 pre-filter radius^2 for candidates that can still WIN: from min(d_box^2, bd2)
    float mf;           // smallest fp32 squared distance (cell units) of a scanned candidate that was not evaluated
};

// Search skipping (K3): besides the winner a search reports a lower bound `lb` on the distance from the query to
// every model point OTHER than the winner.  Everything the search prunes lies outside the pruning radius, every
// candidate it scans is either evaluated exactly (runner-up sd2) or has an fp32 distance above candf (tracked in
// mf), hence  lb^2 = min(sd2, lower bound from mf, pruning radius^2).  When the query moves by delta the bound
// drops to lb - delta; as long as the old winner stays strictly inside it, it provably is still the exact nearest
// neighbour (no tie possible) and the next search is skipped (icpb_nn_skip).  rs2 widens the pruning radius beyond
// what the answer needs so that the bound is useful; it never changes the answer.
ICPB_HD void icpb_nn_refresh_bound(NNState &s)
{
    const double need = (s.dbox2s < s.bd2) ? s.dbox2s : s.bd2;
    s.candf = icpb_cand_radius2(icpb_bound_cells(need, s.inv_h2), s.slack);
    s.boundf = icpb_bound_cells(need < s.rs2 ? s.rs2 : need, s.inv_h2);
}

ICPB_HD void icpb_nn_eval(const MPoint &m, uint32_t pos, NNState &s)
{
    const double dx = s.qx - m.x, dy = s.qy - m.y, dz = s.qz - m.z;
    const double d2 = (dx * dx + dy * dy) + dz * dz;
    ICPB_COUNT(exact);
    if (d2 < s.bd2 || (d2 == s.bd2 && m.idx < s.bidx)) {
        const bool nearer = d2 < s.bd2;
        if (s.bd2 < s.sd2) s.sd2 = s.bd2; // the old best becomes a runner-up
        s.bd2 = d2;
        s.bidx = m.idx;
        s.bpos = pos;
        if (nearer) icpb_nn_refresh_bound(s);
    } else if (pos != s.bpos && d2 < s.sd2) {
        s.sd2 = d2;
    }
}

ICPB_HD float icpb_dist2f(const FPoint &f, const NNState &s)
{
    const float dx = s.ux - f.x, dy = s.uy - f.y, dz = s.uz - f.z;
    return ICPB_FMAF(dz, dz, ICPB_FMAF(dy, dy, dx * dx));
}
// Scan up to four contiguous runs [s_k, e_k) of pts (empty runs have e_k <= s_k) in ONE loop: every trip tests four
// candidates of the lane's current run (the loads are issued together, the tail repeats the last point) and moves on
// to the lane's next run when this one is exhausted -- the lanes of a warp stay busy until their own candidates run
// out instead of waiting for each other run by run.  The fp32 pre-filter only MARKS the candidates that can still
// win (one bit per point, 32 points at a time); the marked ones are then evaluated exactly back to back.  The point
// that already is the best (the warm start, usually) is not evaluated again.
ICPB_HD void icpb_flush_marks(const GridView &g, uint32_t base, uint32_t &marks, NNState &s)
{
    while (marks) {
#ifdef __CUDA_ARCH__
        const uint32_t pos = base + (uint32_t)(__ffs((int)marks) - 1);
#else
        const uint32_t pos = base + (uint32_t)__builtin_ctz(marks);
#endif
        marks &= marks - 1u;
        if (pos != s.bpos) icpb_nn_eval(g.pts[pos], pos, s);
    }
}
template <bool TRACK>
ICPB_HD void icpb_scan_runs(const GridView &g, uint32_t s0, uint32_t e0, uint32_t s1, uint32_t e1, uint32_t s2,
                            uint32_t e2, uint32_t s3, uint32_t e3, NNState &s)
{
    uint32_t p = s0, e = e0, base = s0, marks = 0;
    int r = 0;
    for (;;) {
        if (p >= e) { // this run is done: evaluate what it marked, then the lane's next non-empty run
            icpb_flush_marks(g, base, marks, s);
            do {
                ++r;
                if (r > 3) return;
                p = r == 1 ? s1 : (r == 2 ? s2 : s3);
                e = r == 1 ? e1 : (r == 2 ? e2 : e3);
            } while (p >= e);
            base = p;
        }
        ICPB_COUNT(cells);
        // four candidates per trip, no tail handling: past the end of the run lie the next cell's points (real model
        // points: testing them is harmless) and pts/ptsf end in ICPB_PAD far-away sentinels
        const FPoint f0 = g.ptsf[p], f1 = g.ptsf[p + 1], f2 = g.ptsf[p + 2], f3 = g.ptsf[p + 3];
        const float d0 = icpb_dist2f(f0, s), d1 = icpb_dist2f(f1, s), d2 = icpb_dist2f(f2, s), d3 = icpb_dist2f(f3, s);
        const float cf = s.candf;
        const uint32_t sh = p - base;
        marks |= ((d0 <= cf ? 1u : 0u) | (d1 <= cf ? 2u : 0u) | (d2 <= cf ? 4u : 0u) | (d3 <= cf ? 8u : 0u)) << sh;
        // candidates that are not marked are never evaluated: their fp32 distance bounds the runner-up
        if (TRACK)
            s.mf = fminf(s.mf, fminf(fminf(d0 <= cf ? 3.0e38f : d0, d1 <= cf ? 3.0e38f : d1),
                                     fminf(d2 <= cf ? 3.0e38f : d2, d3 <= cf ? 3.0e38f : d3)));
        p += 4;
        // 32 marks are full -- or there is no bound yet (first candidates of a cold search): evaluate now
        if ((sh == 28u || !(cf < 3.0e38f)) && p < e) {
            icpb_flush_marks(g, base, marks, s);
            base = p;
        }
    }
}

// The walk below scans one level-0 cell at a time, where only a few lanes of a warp are at a cell together: there the
// candidates that pass the pre-filter are evaluated on the spot (the bound tightens at once and prunes the walk).
// Scan the contiguous run [p, e) of pts: four candidates per step (the loads are issued together, the tail repeats
// the last point), exact evaluation only for those inside the pre-filter radius
ICPB_HD void icpb_scan_run(const GridView &g, uint32_t p, uint32_t e, NNState &s)
{
    ICPB_COUNT(cells);
    for (; p < e; p += 4) {
        const FPoint f0 = g.ptsf[p], f1 = g.ptsf[p + 1], f2 = g.ptsf[p + 2], f3 = g.ptsf[p + 3]; // see icpb_scan_runs
        const float d0 = icpb_dist2f(f0, s), d1 = icpb_dist2f(f1, s), d2 = icpb_dist2f(f2, s), d3 = icpb_dist2f(f3, s);
        if (fminf(fminf(d0, d1), fminf(d2, d3)) > s.candf) continue; // none can be nearer than the best so far (nor tie)
        if (d0 <= s.candf) icpb_nn_eval(g.pts[p], p, s);
        if (d1 <= s.candf) icpb_nn_eval(g.pts[p + 1], p + 1, s);
        if (d2 <= s.candf) icpb_nn_eval(g.pts[p + 2], p + 2, s);
        if (d3 <= s.candf) icpb_nn_eval(g.pts[p + 3], p + 3, s);
    }
}

// child visiting order as XOR codes relative to the child octant the query is in: 0,1,2,4,3,5,6,7
#define ICPB_ORD 0x76534210u

// the child octant of node (cx,cy,cz) at `lev` that holds the query (clamped to the node)
ICPB_HD unsigned icpb_near_code(const NNState &s, int lev, int cx, int cy, int cz)
{
    const float hc = (float)(1u << (lev - 1));
    return (s.ux >= (float)(2 * cx + 1) * hc ? 1u : 0u) | (s.uy >= (float)(2 * cy + 1) * hc ? 2u : 0u) |
           (s.uz >= (float)(2 * cz + 1) * hc ? 4u : 0u);
}
// occupancy bits of node word w permuted into visiting order (bit k <-> child near ^ ORD[k])
ICPB_HD unsigned icpb_pending(uint32_t w, unsigned near)
{
    unsigned pend = 0;
#ifdef __CUDACC__
#pragma unroll
#endif
    for (int k = 0; k < 8; ++k) pend |= ((w >> (near ^ ((ICPB_ORD >> (4 * k)) & 7u))) & 1u) << k;
    return pend;
}

// Depth-first walk below node (cx,cy,cz) of level `top` >= 1 whose word is w (already known to be non-empty and
// within the bound).  trail byte b holds the pending children of the path's node at level b+1 in visiting order
// (bit k <-> child near ^ ORD[k]).
ICPB_HD void icpb_walk(const GridView &g, int top, int cx, int cy, int cz, uint32_t w, NNState &s)
{
    // The current node's pending children (visiting order) and its `near` octant live in registers; the trail is
    // written when the walk descends and read when it comes back up, not at every child.
    unsigned long long tr0 = 0;
    unsigned int tr1 = 0;
    int lev = top;
    unsigned near = icpb_near_code(s, lev, cx, cy, cz);
    unsigned pend = icpb_pending(w, near);
    ICPB_COUNT(expands);
    for (;;) {
        if (pend == 0) { // this node is finished: back to its parent
            if (lev == top) return;
            ++lev;
            cx >>= 1, cy >>= 1, cz >>= 1;
            const int b = lev - 1;
            pend = b < 8 ? (unsigned)(tr0 >> (8 * b)) & 255u : (tr1 >> (8 * (b - 8))) & 255u;
            near = icpb_near_code(s, lev, cx, cy, cz);
            continue;
        }
#ifdef __CUDA_ARCH__
        const int k = __ffs((int)pend) - 1;
#else
        const int k = __builtin_ctz(pend);
#endif
        pend &= pend - 1u;
        const unsigned ch = near ^ ((ICPB_ORD >> (4 * k)) & 7u);
        const int x = 2 * cx + (int)(ch & 1u), y = 2 * cy + (int)((ch >> 1) & 1u), z = 2 * cz + (int)((ch >> 2) & 1u);
        ICPB_COUNT(pops);
        if (icpb_cell_mindist2f(lev - 1, x, y, z, s.ux, s.uy, s.uz, s.slack) > s.boundf) continue;
        if (lev == 1) {
            const uint32_t c = ((uint32_t)z * (uint32_t)g.dim[0][1] + (uint32_t)y) * (uint32_t)g.dim[0][0] + (uint32_t)x;
            const uint32_t cb = g.cell_box[c], cs = g.cell_start[c], ce = g.cell_start[c + 1];
            if (icpb_node_mindist2f(cb, 0, x, y, z, s.ux, s.uy, s.uz, s.slack) > s.boundf) continue; // the cell's tight box
            icpb_scan_run(g, cs, ce, s);
            continue;
        }
        const size_t ci = ((size_t)z * g.dim[lev - 1][1] + y) * g.dim[lev - 1][0] + x;
        const uint32_t cw = g.node[lev - 1][ci];
        ICPB_COUNT(nodes);
        if (icpb_node_mindist2f(cw, lev - 1, x, y, z, s.ux, s.uy, s.uz, s.slack) > s.boundf) continue; // tight box
        const int b = lev - 1; // descend: park this node's remaining children on the trail
        if (b < 8)
            tr0 = (tr0 & ~(255ull << (8 * b))) | ((unsigned long long)pend << (8 * b));
        else
            tr1 = (tr1 & ~(255u << (8 * (b - 8)))) | (pend << (8 * (b - 8)));
        --lev;
        cx = x, cy = y, cz = z;
        near = icpb_near_code(s, lev, cx, cy, cz);
        pend = icpb_pending(cw, near);
        ICPB_COUNT(expands);
    }
}

// ---- the search, in the pieces the kernels compose -----------------------------------------------------------------
// What the bound of one query implies for the grid: the level whose cells cover the ball with <= 2 per axis and those
// cells (`cn` holds the query or is the closest, `cf` is the other one, == cn when the ball stays inside one cell).
struct NNPlan {
    int l;      // start level; -1: nothing to search (empty model, or the ball misses the grid)
    int cn[3], cf[3];
};

// state of a fresh query: cell-unit copy of the point, nothing found yet, the bound is d_box
ICPB_HD void icpb_nn_init(const GridView &g, double qx, double qy, double qz, double dbox2s, NNState &s, double rs2 = 0.0)
{
    s.qx = qx, s.qy = qy, s.qz = qz;
    s.dbox2s = dbox2s;
    s.inv_h2 = g.inv_h * g.inv_h;
    s.bd2 = icpb_inf();
    s.sd2 = icpb_inf();
    s.rs2 = rs2;
    s.mf = 3.0e38f;
    s.bidx = ~0ull;
    s.bpos = ICPB_NONE;
    s.ux = (float)((qx - g.o[0]) * g.inv_h);
    s.uy = (float)((qy - g.o[1]) * g.inv_h);
    s.uz = (float)((qz - g.o[2]) * g.inv_h);
    s.slack = 1e-6f * (fmaxf(fmaxf(fabsf(s.ux), fabsf(s.uy)), fabsf(s.uz)) + 4096.0f);
    icpb_nn_refresh_bound(s);
}

// start level: smallest l with 2^l >= 2*(ru + slack), so the ball spans <= 2 cells per axis even after the fp32
// rounding of the range arithmetic in icpb_nn_plan (which the extra slack dominates).  Never above level
// ICPB_WALK_TOP: the walk's trail holds one byte per level in 12 bytes of registers, and a grid of at most 8192 cells
// per axis (13 levels) has at most 2 nodes per axis there anyway -- all of them are roots then.
#define ICPB_WALK_TOP 12
ICPB_HD int icpb_nn_start_level(const GridView &g, const NNState &s)
{
    const float ru = sqrtf(s.boundf) * 1.000001f + s.slack;
    const int top = g.L < ICPB_WALK_TOP ? g.L : ICPB_WALK_TOP;
    int l = 0;
    if (ru < 1.0e30f) {
        while (l < top && (float)(1u << l) < 2.0f * (ru + s.slack)) ++l;
    } else {
        l = top;
    }
    return l;
}

// from the current bound to the cells to search
ICPB_HD void icpb_nn_plan(const GridView &g, const NNState &s, NNPlan &pl)
{
    pl.l = -1;
    for (int a = 0; a < 3; ++a) pl.cn[a] = pl.cf[a] = 0;
    if (g.n == 0) return;
    const float u[3] = {s.ux, s.uy, s.uz};
    // radius of the ball in cell units, rounded up: covers the fp32 copy's distance from the exact query coordinate
    const float ru = sqrtf(s.boundf) * 1.000001f + s.slack;
    const int l = icpb_nn_start_level(g, s);
    const float inv_s = 1.0f / (float)(1u << l);
    for (int a = 0; a < 3; ++a) {
        const float dm1 = (float)(g.dim[l][a] - 1);
        float flo = floorf((u[a] - ru) * inv_s), fhi = floorf((u[a] + ru) * inv_s);
        if (!(flo > 0.0f)) flo = 0.0f; // also -inf / NaN
        if (!(fhi < dm1)) fhi = dm1;
        if (flo > fhi) return; // the ball misses the grid on this axis
        float fc = floorf(u[a] * inv_s);
        fc = fminf(fmaxf(fc, flo), fhi);
        pl.cn[a] = (int)fc;
        pl.cf[a] = (fc == flo) ? (int)fhi : (int)flo;
    }
    pl.l = l;
}

// pl.l == 0: at most 2x2x2 level-0 cells; x-adjacent cells are contiguous in pts[]: <= 4 runs, the query's column first
template <bool TRACK>
ICPB_HD void icpb_nn_fast(const GridView &g, NNState &s, const NNPlan &pl)
{
    const int *cn = pl.cn, *cf = pl.cf;
    const uint32_t nx0 = (uint32_t)g.dim[0][0], ny0 = (uint32_t)g.dim[0][1];
    const int x0 = cn[0] < cf[0] ? cn[0] : cf[0], x1 = cn[0] < cf[0] ? cf[0] : cn[0];
    const float gx = fminf(icpb_gap2f((float)x0, (float)(x0 + 1), s.ux, s.slack), icpb_gap2f((float)x1, (float)(x1 + 1), s.ux, s.slack));
    uint32_t rs[4], re[4];
#ifdef __CUDACC__
#pragma unroll
#endif
    for (int k = 0; k < 4; ++k) {
        const int y = (k & 1) ? cf[1] : cn[1], z = (k & 2) ? cf[2] : cn[2];
        rs[k] = re[k] = 0;
        if (((k & 1) && cf[1] == cn[1]) || ((k & 2) && cf[2] == cn[2])) continue;
        const float md = (gx + icpb_gap2f((float)y, (float)(y + 1), s.uy, s.slack) + icpb_gap2f((float)z, (float)(z + 1), s.uz, s.slack)) * 0.999999f;
        ICPB_COUNT(pops);
        if (md > s.boundf) continue;
        const uint32_t c = ((uint32_t)z * ny0 + (uint32_t)y) * nx0;
        rs[k] = g.cell_start[c + (uint32_t)x0];
        re[k] = g.cell_start[c + (uint32_t)x1 + 1u];
    }
    icpb_scan_runs<TRACK>(g, rs[0], re[0], rs[1], re[1], rs[2], re[2], rs[3], re[3], s);
}

// pl.l >= 1: walk the pyramid below the <= 2x2x2 root nodes of level l, nearest root first
ICPB_HD void icpb_nn_walk_roots(const GridView &g, NNState &s, const NNPlan &pl)
{
    const int l = pl.l;
    const int *cn = pl.cn, *cf = pl.cf;
    for (int k = 0; k < 8; ++k) {
        const unsigned code = (ICPB_ORD >> (4 * k)) & 7u;
        if (((code & 1u) && cf[0] == cn[0]) || ((code & 2u) && cf[1] == cn[1]) || ((code & 4u) && cf[2] == cn[2])) continue;
        const int x = (code & 1u) ? cf[0] : cn[0], y = (code & 2u) ? cf[1] : cn[1], z = (code & 4u) ? cf[2] : cn[2];
        ICPB_COUNT(pops);
        if (icpb_cell_mindist2f(l, x, y, z, s.ux, s.uy, s.uz, s.slack) > s.boundf) continue;
        const size_t ci = ((size_t)z * g.dim[l][1] + y) * g.dim[l][0] + x;
        const uint32_t w = g.node[l][ci];
        ICPB_COUNT(nodes);
        if (!(w & 255u)) continue;
        if (icpb_node_mindist2f(w, l, x, y, z, s.ux, s.uy, s.uz, s.slack) > s.boundf) continue;
        icpb_walk(g, l, x, y, z, w, s);
    }
}

// After a search: lower bound (global units, rounded down) on the distance from the query to every model point
// other than the reported best (to every model point when nothing was found).  See the comment at NNState.
ICPB_HD float icpb_nn_lower_bound(const GridView &g, const NNState &s)
{
    const double need = (s.dbox2s < s.bd2) ? s.dbox2s : s.bd2;
    double lb2 = need < s.rs2 ? s.rs2 : need; // everything pruned lies outside the pruning radius
    if (s.sd2 < lb2) lb2 = s.sd2;
    float lb = !(lb2 < 1.0e37) ? 3.0e18f : sqrtf((float)lb2) * 0.999999f;
    // scanned, never evaluated: a candidate's true cell-unit distance is >= sqrt(its fp32 d^2) - 2*slack (see
    // icpb_cand_radius2); its fp32 d^2 was tracked in mf, or -- without tracking, and in the pyramid walk -- is only
    // known to exceed the pre-filter radius, which never is below what the final answer needs
    const float f2 = s.rs2 > 0.0 ? s.mf : icpb_bound_cells(need, s.inv_h2);
    if (f2 < 3.0e38f) lb = fminf(lb, (sqrtf(f2) * 0.99999f - 2.0f * s.slack) * (float)g.h * 0.999999f);
    return lb > 0.0f ? lb : 0.0f;
}

// ---- flat block scan (start level <= 1: the pruning ball spans at most 3 level-0 cells per axis) -------------------
// The pyramid walk pays for its pruning with a divergent depth-first traversal.  When the ball is at most two cell
// edges wide it touches at most 3x3x3 level-0 cells, and those are enumerated directly instead: per (y,z) row the
// cells that survive the cube test, are not empty and whose TIGHT box (cell_box) the ball reaches; x-adjacent cells
// are contiguous in pts[], so a row contributes one run.  Scenes are surfaces: a query some way off the surface
// prunes most of the block on the tight boxes, like the walk does, but with independent loads and no traversal.
// The runs are then scanned in one flattened loop with deferred exact evaluation (see icpb_scan_runs).
#define ICPB_BLOCK_ROWS 9
struct NNRun {
    uint32_t s, e;
};

ICPB_HD bool icpb_nn_block_ranges(const GridView &g, const NNState &s, int lo[3], int hi[3])
{
    if (g.n == 0) return false;
    const float ru = sqrtf(s.boundf) * 1.000001f + s.slack;
    const float u[3] = {s.ux, s.uy, s.uz};
    for (int a = 0; a < 3; ++a) {
        const float dm1 = (float)(g.dim[0][a] - 1);
        float flo = floorf(u[a] - ru), fhi = floorf(u[a] + ru);
        if (!(flo > 0.0f)) flo = 0.0f;
        if (!(fhi < dm1)) fhi = dm1;
        if (flo > fhi) return false; // the ball misses the grid on this axis
        lo[a] = (int)flo;
        hi[a] = (int)fhi;
    }
    return true;
}

template <bool TRACK>
ICPB_HD void icpb_nn_block(const GridView &g, NNState &s, NNRun *runs, int stride)
{
    int lo[3], hi[3];
    if (!icpb_nn_block_ranges(g, s, lo, hi)) return;
    const uint32_t nx0 = (uint32_t)g.dim[0][0], ny0 = (uint32_t)g.dim[0][1];
    const float bf = s.boundf;
    int nr = 0;
    for (int z = lo[2]; z <= hi[2]; ++z) {
        const float gz = icpb_gap2f((float)z, (float)(z + 1), s.uz, s.slack);
        for (int y = lo[1]; y <= hi[1]; ++y) {
            const float gyz = gz + icpb_gap2f((float)y, (float)(y + 1), s.uy, s.slack);
            ICPB_COUNT(pops);
            if (gyz * 0.999999f > bf) continue; // the whole row is out of reach
            const uint32_t row = ((uint32_t)z * ny0 + (uint32_t)y) * nx0;
            uint32_t rs = 0, re = 0;
            for (int x = lo[0]; x <= hi[0]; ++x) {
                if ((gyz + icpb_gap2f((float)x, (float)(x + 1), s.ux, s.slack)) * 0.999999f > bf) continue;
                const uint32_t c = row + (uint32_t)x;
                const uint32_t cs = g.cell_start[c], ce = g.cell_start[c + 1];
                if (cs == ce) continue;
                ICPB_COUNT(nodes);
                if (icpb_node_mindist2f(g.cell_box[c], 0, x, y, z, s.ux, s.uy, s.uz, s.slack) > bf) continue; // tight box
                if (re == 0) rs = cs;
                re = ce;
            }
            if (re > rs && nr < ICPB_BLOCK_ROWS) {
                runs[nr * stride].s = rs;
                runs[nr * stride].e = re;
                ++nr;
            }
        }
    }
    // one flattened loop over the runs: four candidates per trip, marks (icpb_scan_runs), exact evaluation at the end
    uint32_t p = 0, e = 0, base = 0, marks = 0;
    int r = -1, nm = 0;
    for (;;) {
        if (p >= e) {
            // the marks of a run's last chunk wait in a slot the scan is through with until every run is scanned: the exact
            // evaluations then happen where the lanes of the warp have come together again (cfg2: 16.0 vs 16.7 ms of
            // search per align against evaluating at the end of every run)
            if (marks) { // slots 0..r are free: this run's range has been read
                runs[nm * stride].s = base;
                runs[nm * stride].e = marks;
                marks = 0;
                ++nm;
            }
            ++r;
            if (r >= nr) break;
            p = runs[r * stride].s;
            e = runs[r * stride].e;
            base = p;
        }
        ICPB_COUNT(cells);
        const FPoint f0 = g.ptsf[p], f1 = g.ptsf[p + 1], f2 = g.ptsf[p + 2], f3 = g.ptsf[p + 3];
        const float d0 = icpb_dist2f(f0, s), d1 = icpb_dist2f(f1, s), d2 = icpb_dist2f(f2, s), d3 = icpb_dist2f(f3, s);
        const float cf = s.candf;
        const uint32_t sh = p - base;
        marks |= ((d0 <= cf ? 1u : 0u) | (d1 <= cf ? 2u : 0u) | (d2 <= cf ? 4u : 0u) | (d3 <= cf ? 8u : 0u)) << sh;
        if (TRACK)
            s.mf = fminf(s.mf, fminf(fminf(d0 <= cf ? 3.0e38f : d0, d1 <= cf ? 3.0e38f : d1),
                                     fminf(d2 <= cf ? 3.0e38f : d2, d3 <= cf ? 3.0e38f : d3)));
        p += 4;
        if ((sh == 28u || !(cf < 3.0e38f)) && p < e) {
            icpb_flush_marks(g, base, marks, s);
            base = p;
        }
    }
    for (r = 0; r < nm; ++r) { // lanes that marked anything mostly did so in one run: their evaluations coincide at r == 0
        uint32_t mk = runs[r * stride].e;
        icpb_flush_marks(g, runs[r * stride].s, mk, s);
    }
}

// The search proper for a prepared state (icpb_nn_init + optional warm start): flat block scan, fast path or
// pyramid walk.  runs: scratch for ICPB_BLOCK_ROWS entries, `stride` apart (shared memory on the device).
ICPB_HD int icpb_nn_run(const GridView &g, NNState &s, NNRun *runs, int stride)
{
    if (g.n == 0) return 0;
    const int l0 = icpb_nn_start_level(g, s);
    const bool track = s.rs2 > 0.0; // the candidates' fp32 distances are tracked only when the bound can be useful
    if (l0 <= g.block_levels) {
        if (track)
            icpb_nn_block<true>(g, s, runs, stride);
        else
            icpb_nn_block<false>(g, s, runs, stride);
        return l0;
    }
    NNPlan pl;
    icpb_nn_plan(g, s, pl);
    if (pl.l == 0) {
        if (track)
            icpb_nn_fast<true>(g, s, pl);
        else
            icpb_nn_fast<false>(g, s, pl);
    } else if (pl.l > 0) {
        icpb_nn_walk_roots(g, s, pl);
    }
    return pl.l > 0 ? pl.l : 0;
}

// Exact NN of q by one thread on its own.  dbox2s = d_box^2 widened by a few ulps (+inf when unbounded);
// warm_pos = position in pts[] of a candidate to seed the bound (ICPB_NONE for none); rs2 = squared radius the
// search is made to cover at least (0: only what the answer needs) so that *out_lb is useful.
// Returns the position in pts[] (ICPB_NONE if the model is empty or nothing lies
// within the bound) and the exact squared distance; start_level = level the search started at.
ICPB_HD void icpb_nn_search(const GridView &g, double qx, double qy, double qz, double dbox2s, uint32_t warm_pos,
                            uint32_t &out_pos, double &out_d2, int &start_level, double rs2 = 0.0, float *out_lb = nullptr)
{
    NNState s;
    NNRun runs[ICPB_BLOCK_ROWS];
    icpb_nn_init(g, qx, qy, qz, dbox2s, s, rs2);
    if (warm_pos < g.n) icpb_nn_eval(g.pts[warm_pos], warm_pos, s);
    start_level = icpb_nn_run(g, s, runs, 1);
    out_pos = s.bpos;
    out_d2 = s.bd2;
    if (out_lb) *out_lb = icpb_nn_lower_bound(g, s);
}

// ---- search skipping ------------------------------------------------------------------------------------------------
struct NNTune {
    float reach_cells; // a search whose needed radius is at most this many cell edges (the converged regime) ...
    float floor_cells; // ... is made to cover this many cell edges more, so that its bound lets later iterations skip
    int skip;          // 0: every query is searched every iteration
};

// Upper bound (fp32, rounded up) of |dM * v|: how far the query moved when its transform changed by
// dM = M_new - M_old (3x4 row-major, already rounded to float).  The fp32 evaluation error is covered per axis.
ICPB_HD float icpb_motion_bound(const float *dM, float vx, float vy, float vz)
{
    float n2 = 0.0f;
#ifdef __CUDACC__
#pragma unroll
#endif
    for (int r = 0; r < 3; ++r) {
        const float a = dM[4 * r] * vx, b = dM[4 * r + 1] * vy, c = dM[4 * r + 2] * vz, t = dM[4 * r + 3];
        const float v = fabsf((a + b) + (c + t)) + 1.0e-6f * ((fabsf(a) + fabsf(b)) + (fabsf(c) + fabsf(t)));
        n2 = ICPB_FMAF(v, v, n2);
    }
    return sqrtf(n2) * 1.00001f + 1.0e-30f;
}

// Can the search of this query be skipped?  wm/has_w: the winner of its last search, lb_prev: the bound of that search
// (already reduced by earlier motions), delta: upper bound of the motion since then.  Returns 1 (the old winner is
// still the exact nearest neighbour, squared distance d2w) or 2 (provably no pair within d_box; d2w = inf) when it
// can; always returns the reduced bound in lb_out and -- for the warm start of the search that follows otherwise --
// the exact squared distance to the old winner in d2w (inf without one).
ICPB_HD int icpb_nn_skip_pt(double qx, double qy, double qz, double d_box, bool has_w, const MPoint &wm, float lb_prev,
                            float delta, double &d2w, float &lb_out)
{
    float lbn = (lb_prev - delta) * 0.999999f;
    if (!(lbn > 0.0f)) lbn = 0.0f;
    lb_out = lbn;
    d2w = icpb_inf();
    if (has_w) {
        const double dx = qx - wm.x, dy = qy - wm.y, dz = qz - wm.z;
        d2w = (dx * dx + dy * dy) + dz * dz;
    }
    const double lb2 = (double)lbn * (double)lbn; // exact product of two floats
    if (d2w < lb2) return 1;                      // every other point is farther than the old winner: still the exact NN
    const double dbox2s = d_box * d_box * (1.0 + 1e-12);
    if (lb2 > dbox2s && d2w > dbox2s) { // nothing can lie within d_box
        d2w = icpb_inf();
        return 2;
    }
    return 0;
}
ICPB_HD int icpb_nn_skip(const GridView &g, double qx, double qy, double qz, double d_box, uint32_t wpos, float lb_prev,
                         float delta, double &d2w, float &lb_out)
{
    MPoint wm;
    wm.x = wm.y = wm.z = 0.0;
    wm.idx = 0;
    if (wpos < g.n) wm = g.pts[wpos];
    return icpb_nn_skip_pt(qx, qy, qz, d_box, wpos < g.n, wm, lb_prev, delta, d2w, lb_out);
}

// Lazy search: may the search of this query be DEFERRED?  Only the n_po nearest pairs survive the trim
// (Pairs::trim, icp/icp.cpp:23-41); of the others nothing but their existence is ever used (the pair count).  A query
// whose nearest neighbour provably is farther than tau_guard -- min(distance to its old winner, bound on all other
// points) > tau_guard -- while its old winner lies within d_box has a pair that will not be kept, whichever point it
// is with: no search, and the lower bound stands in for its distance (it only has to stay above the new tau, which
// icpb_iteration_finish verifies).  These are the queries with the widest balls, i.e. the most expensive searches.
ICPB_HD int icpb_nn_defer(double d2w, float lbn, double d_box, double tau_guard, double &standin)
{
    if (!(d2w < icpb_inf())) return 0;
    const double dw = sqrt(d2w);
    if (!(dw <= d_box)) return 0; // no sure pair: the nearest neighbour has to be found (or ruled out by the skip test)
    standin = dw < (double)lbn ? dw : (double)lbn;
    return standin > tau_guard ? 1 : 0;
}

// Radius a search that has to happen anyway is made to cover, so that its bound lets later iterations skip (0: only
// what the answer needs).  need2 = min(d_box^2, warm-start distance^2) is what the answer requires.  Measured on cfg2:
// while the source still is a point spacing or more off the model, runner-ups are about as near as winners and a wider
// search only costs; the reach limits the widening to the converged regime, where it is nearly free.
ICPB_HD float icpb_cover_radius(const GridView &g, double need2, float delta, const NNTune &t)
{
    if (!t.skip || !(need2 < 1.0e30)) return 0.0f; // cold search: no bound to speak of
    const float h = (float)g.h;
    const float U = sqrtf((float)need2);
    if (U > t.reach_cells * h) return 0.0f;
    if (delta > 0.25f * t.floor_cells * h) return 0.0f; // still moving fast: the extra reach would be used up in < 4 iterations
    const float R = fminf(U + t.floor_cells * h, 0.48f * h); // stays inside a 2x2x2 cell block, slack included
    return R > U ? R : 0.0f;
}

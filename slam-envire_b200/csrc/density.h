// density.h -- PointcloudAdapter's density stepping in closed form (K0).
//
// Replaces: the visiting order of PointcloudAdapter::next (reference icp/icp.hpp:126-138):
//     idx = size_t(index); index += 1.0 / density;   while idx < n
// The reference accumulates `index` by repeated fp64 addition, so the visited indices are NOT floor(k / density):
// they carry the rounding of every addition.  To gather the visited vertices on the device (all k in parallel) the
// host restates the accumulation in closed form:
//   * while `index` stays inside one binade [2^e, 2^(e+1)) it is a multiple of u = 2^(e-52), and fl(index + step)
//     = index + inc with ONE constant inc (step rounded to the grid of u) -- except that, when step lies exactly
//     half-way between two grid points, round-half-even makes the very first sum of the binade depend on the parity
//     of the value it starts from; every later value is even and the increment is constant from then on;
//   * so the sequence is piecewise linear in exact integer arithmetic: per binade two explicit steps (done with real
//     fp64 additions, which also handle the crossing into the binade) and one run  index_j = X + j * inc.
// A handful of segments (<= 3 per binade) describes the whole order; the device evaluates
//     idx(k) = (base_u + (k - k0) * inc_u) >> shift
// per visited vertex.  tests/test_host_logic.py compares it with the reference loop index by index.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __CUDACC__
#define ICPB_DHD __host__ __device__ __forceinline__
#else
#define ICPB_DHD inline
#endif

struct DensitySeg {
    unsigned long long k0;     // first visit of the segment
    unsigned long long base_u; // index value of visit k0 in units of 2^-shift
    unsigned long long inc_u;  // increment per visit in the same units (0: a single visit)
    int shift;                 // idx = value >> shift (>= 64: idx = 0)
    int pad;
};
#define ICPB_DENSITY_MAX_SEGS 512

// vertex visited at step k (seg: the segment with the largest k0 <= k)
ICPB_DHD unsigned long long icpb_density_index(const DensitySeg &seg, unsigned long long k)
{
    if (seg.shift >= 64) return 0ull;
    return (seg.base_u + (k - seg.k0) * seg.inc_u) >> seg.shift;
}
ICPB_DHD int icpb_density_find(const DensitySeg *segs, int nseg, unsigned long long k)
{
    int lo = 0, hi = nseg; // last segment with k0 <= k
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (segs[mid].k0 <= k) lo = mid;
        else hi = mid;
    }
    return lo;
}

// Host: the segments for n vertices at `density`.  Returns the number of segments (0: density == 1 is the identity --
// not handled here; -1: more than max_segs segments, e.g. density >> 1: gather on the host instead) and the number of
// visited vertices in *visited.
inline int icpb_density_segments(size_t n, double density, DensitySeg *segs, int max_segs, unsigned long long *visited)
{
    const double step = 1.0 / density;
    double x = 0.0;
    unsigned long long k = 0;
    int ns = 0;
    auto units = [](double v, int e) -> unsigned long long { // v / 2^(e-52), exact for multiples of that unit
        return (unsigned long long)ldexp(v, 52 - e);
    };
    auto binade = [](double v) -> int { // e with v in [2^e, 2^(e+1)), v > 0
        int E;
        frexp(v, &E);
        return E - 1;
    };
    auto single = [&](double v) -> bool { // one visit at value v
        if (ns >= max_segs) return false;
        DensitySeg s;
        s.k0 = k;
        s.inc_u = 0;
        s.pad = 0;
        if (v < 1.0) {
            s.base_u = 0;
            s.shift = 0;
        } else {
            const int e = binade(v);
            s.base_u = units(v, e);
            s.shift = 52 - e;
        }
        segs[ns++] = s;
        return true;
    };
    for (;;) {
        // two explicit visits: the sums that enter a binade (and settle the parity there) are real fp64 additions
        for (int rep = 0; rep < 2; ++rep) {
            if (!((size_t)x < n)) goto done;
            if (!single(x)) return -1;
            x += step;
            ++k;
        }
        if (!((size_t)x < n)) goto done;
        if (!(x >= 1.0)) continue; // below 1 every visit is vertex 0; no run needed (density > 1)
        const int e = binade(x);
        if (e > 52) return -1;
        const double y = x + step;
        if (!(y > x) || binade(y) != e) continue; // the next sum leaves the binade already
        const unsigned long long X = units(x, e), inc = units(y, e) - X, top = 1ull << 53;
        if (inc == 0) continue;
        const int sh = 52 - e;
        unsigned long long J = (top - 1ull - X) / inc; // visits j = 0..J stay inside the binade
        // ... and below n: (X + j * inc) >> sh < n
        const unsigned __int128 lim = (unsigned __int128)n << sh;
        if ((unsigned __int128)X + (unsigned __int128)J * inc >= lim) {
            if ((unsigned __int128)X >= lim) goto done;
            J = (unsigned long long)((lim - 1 - X) / inc);
        }
        if (ns >= max_segs) return -1;
        DensitySeg s;
        s.k0 = k;
        s.base_u = X;
        s.inc_u = inc;
        s.shift = sh;
        s.pad = 0;
        segs[ns++] = s;
        x = ldexp((double)(X + J * inc), e - 52); // the last value of the run (exact: a multiple of the unit below 2^53 units)
        k += J + 1;
        x += step;
    }
done:
    *visited = k;
    return ns;
}

// primitives.cuh -- hand-written device primitives used by the grid build and
// by K7: exclusive scan (u32) and a stable LSD radix sort (u32 or u64 keys,
// optional u32 payload).  sm_100a; no CUB/Thrust on the product path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace icpb {

// ---------------------------------------------------------------------------
// exclusive scan over uint32, arbitrary length, in-place allowed
// ---------------------------------------------------------------------------
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *warp_sums, uint32_t &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < nwarps ? warp_sums[lane] : 0u;
        uint32_t winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < nwarps) warp_sums[lane] = winc - w; // exclusive warp offsets
        if (lane == 31) warp_sums[32] = winc;          // block total
    }
    __syncthreads();
    total = warp_sums[32];
    return warp_sums[warp] + inc - v;
}

// each block scans one tile; tile totals go to sums[] (when not null)
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_kernel(const uint32_t *in, uint32_t *out, size_t n,
                                                                 uint32_t *sums)
{
    __shared__ uint32_t warp_sums[33];
    const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t local = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0u;
        local += v[k];
    }
    uint32_t total;
    uint32_t off = block_exclusive_scan(local, warp_sums, total);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) out[base + k] = off;
        off += v[k];
    }
    if (sums && threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_add_kernel(uint32_t *out, size_t n, const uint32_t *sums)
{
    const uint32_t add = sums[blockIdx.x];
    const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) out[base + k] += add;
}

// scratch must hold scan_scratch_elems(n) uint32
inline size_t scan_scratch_elems(size_t n)
{
    size_t total = 0;
    while (n > (size_t)SCAN_TILE) {
        n = (n + SCAN_TILE - 1) / SCAN_TILE;
        total += n + 32;
    }
    return total + 32;
}

inline void exclusive_scan_u32(const uint32_t *in, uint32_t *out, size_t n, uint32_t *scratch, cudaStream_t s,
                               unsigned long long *launches)
{
    if (n == 0) return;
    const size_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (nb == 1) {
        scan_tile_kernel<<<1, SCAN_THREADS, 0, s>>>(in, out, n, nullptr);
        if (launches) ++*launches;
        return;
    }
    scan_tile_kernel<<<(unsigned)nb, SCAN_THREADS, 0, s>>>(in, out, n, scratch);
    if (launches) ++*launches;
    exclusive_scan_u32(scratch, scratch, nb, scratch + nb + 32, s, launches);
    scan_add_kernel<<<(unsigned)nb, SCAN_THREADS, 0, s>>>(out, n, scratch);
    if (launches) ++*launches;
}

// ---------------------------------------------------------------------------
// stable LSD radix sort, 8 bits per pass
// ---------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_STEPS = 16;                  // 32-element steps per warp
constexpr int RS_WARP_TILE = 32 * RS_STEPS;   // contiguous elements owned by one warp
constexpr int RS_TILE = RS_WARP_TILE * RS_WARPS;

template <typename K>
__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const K *keys, size_t n, int shift, uint32_t *block_hist,
                                                             uint32_t nb)
{
    __shared__ uint32_t hist[256];
    hist[threadIdx.x] = 0;
    __syncthreads();
    const size_t base = (size_t)blockIdx.x * RS_TILE;
    for (int k = threadIdx.x; k < RS_TILE; k += RS_THREADS) {
        const size_t i = base + k;
        if (i < n) atomicAdd(&hist[(unsigned)(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    block_hist[(size_t)threadIdx.x * nb + blockIdx.x] = hist[threadIdx.x]; // bin-major: one scan gives all offsets
}

template <typename K, bool HAS_VAL>
__global__ void __launch_bounds__(RS_THREADS)
    rs_scatter_kernel(const K *keys_in, const uint32_t *vals_in, K *keys_out, uint32_t *vals_out, size_t n, int shift,
                      const uint32_t *offsets, uint32_t nb)
{
    __shared__ uint32_t whist[RS_WARPS][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = threadIdx.x; k < RS_WARPS * 256; k += RS_THREADS) (&whist[0][0])[k] = 0;
    __syncthreads();
    const size_t wbase = (size_t)blockIdx.x * RS_TILE + (size_t)warp * RS_WARP_TILE;
    const unsigned lt_mask = (1u << lane) - 1u;
    // phase A: per-warp digit counts of the warp's contiguous sub-tile
    for (int st = 0; st < RS_STEPS; ++st) {
        const size_t i = wbase + (size_t)st * 32 + lane;
        const bool valid = i < n;
        const unsigned act = __ballot_sync(0xFFFFFFFFu, valid);
        if (valid) {
            const unsigned d = (unsigned)(keys_in[i] >> shift) & 255u;
            const unsigned peers = __match_any_sync(act, d);
            if ((peers & lt_mask) == 0) whist[warp][d] += __popc(peers);
        }
        __syncwarp();
    }
    __syncthreads();
    // phase B: digit-wise running offsets: global (bin, block) offset, then warps in order
    {
        const int d = threadIdx.x;
        uint32_t run = offsets[(size_t)d * nb + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            const uint32_t c = whist[w][d];
            whist[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
    // phase C: stable scatter
    for (int st = 0; st < RS_STEPS; ++st) {
        const size_t i = wbase + (size_t)st * 32 + lane;
        const bool valid = i < n;
        const unsigned act = __ballot_sync(0xFFFFFFFFu, valid);
        if (valid) {
            const K key = keys_in[i];
            const unsigned d = (unsigned)(key >> shift) & 255u;
            const unsigned peers = __match_any_sync(act, d);
            const uint32_t pos = whist[warp][d] + __popc(peers & lt_mask);
            keys_out[pos] = key;
            if (HAS_VAL) vals_out[pos] = vals_in[i];
            __syncwarp(act);
            if ((peers & lt_mask) == 0) whist[warp][d] += __popc(peers);
        }
        __syncwarp();
    }
}

inline size_t rs_num_blocks(size_t n) { return (n + RS_TILE - 1) / RS_TILE; }
// scratch (uint32 elements) needed by radix_sort for n keys
inline size_t rs_scratch_elems(size_t n)
{
    const size_t nb = rs_num_blocks(n);
    return 256 * nb + 32 + scan_scratch_elems(256 * nb);
}

// Sorts `bits` low bits of the keys (rounded up to whole bytes).  Ping-pongs
// between (k0,v0) and (k1,v1); returns 0 if the result is in (k0,v0), 1 if in (k1,v1).
template <typename K, bool HAS_VAL>
inline int radix_sort(K *k0, uint32_t *v0, K *k1, uint32_t *v1, size_t n, int bits, uint32_t *scratch,
                      cudaStream_t s, unsigned long long *launches)
{
    if (n == 0) return 0;
    const uint32_t nb = (uint32_t)rs_num_blocks(n);
    uint32_t *block_hist = scratch;
    uint32_t *scan_scratch = scratch + 256 * (size_t)nb + 32;
    int cur = 0;
    for (int shift = 0; shift < bits; shift += 8) {
        K *ki = cur ? k1 : k0, *ko = cur ? k0 : k1;
        uint32_t *vi = cur ? v1 : v0, *vo = cur ? v0 : v1;
        rs_hist_kernel<K><<<nb, RS_THREADS, 0, s>>>(ki, n, shift, block_hist, nb);
        exclusive_scan_u32(block_hist, block_hist, 256 * (size_t)nb, scan_scratch, s, launches);
        rs_scatter_kernel<K, HAS_VAL><<<nb, RS_THREADS, 0, s>>>(ki, vi, ko, vo, n, shift, block_hist, nb);
        if (launches) *launches += 2;
        cur ^= 1;
    }
    return cur;
}

} // namespace icpb

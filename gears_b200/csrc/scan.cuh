// scan.cuh — device-wide exclusive scan of u32 counters (contact-chain offsets).
// reduce-per-tile -> scan of tile sums (one CTA) -> scan-per-tile with carry.  R 8 B + W 4 B per element.
// Measured and rejected (round 2, 58.9 M counters on a B200): a single-pass scan with decoupled look-back (R 4 + W 4,
// one launch; status words with an epoch so nothing is cleared between calls) in four shapes -- 256 or 512 threads,
// 32 to 256 status words per look-back round trip, one tile per CTA or persistent CTAs with the next tile's loads in
// flight during the wait -- took 0.185 to 0.24 ms against 0.146 ms for these three kernels (2 M bodies: 0.031 to 0.043
// against 0.030): a tile of 16 KB lasts ~5 ns at HBM speed while a round trip to the far L2 partition costs ~1 us,
// and the CTAs that wait hold the registers the loads of the next ones need.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "radix_sort.cuh"

namespace gp {

constexpr int SC_THREADS = 256;
constexpr int SC_ITEMS = 16;
constexpr int SC_TILE = SC_THREADS * SC_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t* total, uint32_t* smem33) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) smem33[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < (SC_THREADS / 32) ? smem33[lane] : 0;
        uint32_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        if (lane < (SC_THREADS / 32)) smem33[lane] = wi - w;
        if (lane == 31) smem33[32] = wi;
    }
    __syncthreads();
    const uint32_t r = smem33[warp] + inc - v;
    if (total) *total = smem33[32];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SC_THREADS) sc_reduce(const uint32_t* __restrict__ in, uint32_t n,
                                                        uint32_t* __restrict__ tile_sums,
                                                        const uint32_t* __restrict__ gate) {
    __shared__ uint32_t sm[33];
    if (gate && *gate == 0) return;
    const uint64_t base = (uint64_t)blockIdx.x * SC_TILE;
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SC_ITEMS; ++i) {
        const uint64_t p = base + (uint64_t)i * SC_THREADS + threadIdx.x;
        if (p < n) s += in[p];
    }
    uint32_t total;
    block_exclusive_scan_256(s, &total, sm);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// out[i] = sum_{j<i} in[j]; the last tile also writes out[n] = grand total.
// `in` and `out` may be the same array (every thread reads its 16 elements before it writes them).
// Warp-striped access: item i of lane l sits at warp*512 + i*32 + l, so every load/store instruction of a warp
// covers one contiguous 128 B line; the scan order is restored through a 32x16 transpose in shared memory.
__global__ void __launch_bounds__(SC_THREADS) sc_scan(const uint32_t* in, uint32_t n,
                                                      const uint32_t* __restrict__ tile_offsets, uint32_t* out,
                                                      const uint32_t* __restrict__ gate) {
    __shared__ uint32_t sm[33];
    __shared__ uint32_t tr[SC_THREADS / 32][32 * SC_ITEMS + 32];  // +1 word per 16: conflict-free transposes
    if (gate && *gate == 0) return;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t tile = gridDim.x - 1u - blockIdx.x;  // last tiles first: those are the ones sc_reduce left in L2
    const uint64_t wbase = (uint64_t)tile * SC_TILE + (uint64_t)warp * (32 * SC_ITEMS);
    uint32_t* t = tr[warp];
#pragma unroll
    for (int i = 0; i < SC_ITEMS; ++i) {
        const uint64_t p = wbase + (uint64_t)i * 32 + lane;
        const uint32_t e = i * 32 + lane;  // element index within the warp's 512
        t[e + (e >> 4)] = (p < n) ? in[p] : 0u;
    }
    __syncwarp();
    // lane l owns the 16 consecutive elements [16 l, 16 l + 16)
    uint32_t v[SC_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SC_ITEMS; ++i) {
        const uint32_t e = lane * SC_ITEMS + i;
        v[i] = t[e + (e >> 4)];
        s += v[i];
    }
    uint32_t run = block_exclusive_scan_256(s, nullptr, sm) + tile_offsets[tile];
#pragma unroll
    for (int i = 0; i < SC_ITEMS; ++i) {
        const uint32_t e = lane * SC_ITEMS + i;
        t[e + (e >> 4)] = run;
        run += v[i];
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < SC_ITEMS; ++i) {
        const uint64_t p = wbase + (uint64_t)i * 32 + lane;
        const uint32_t e = i * 32 + lane;
        if (p < n) out[p] = t[e + (e >> 4)];
    }
    // the grand total goes to out[n]: the thread that owns element n-1 knows it
    const uint64_t last = (uint64_t)n - 1;
    // (elements at or beyond n were read as 0, so this thread's running sum after its 16 elements is the total)
    if (n && last >= wbase && last < wbase + 32 * SC_ITEMS && (uint32_t)((last - wbase) >> 4) == lane) out[n] = run;
}

// tile_sums: ceil(n/SC_TILE) u32 of scratch.  out has n+1 entries; out == in (in place) is allowed.
// gate (device pointer, may be null): all three kernels return immediately when *gate == 0.
static inline void exclusive_scan_u32(const uint32_t* in, uint32_t n, uint32_t* out, uint32_t* tile_sums,
                                      cudaStream_t st, uint64_t* launches, const uint32_t* gate = nullptr) {
    if (n == 0) {
        cudaMemsetAsync(out, 0, sizeof(uint32_t), st);
        return;
    }
    const uint32_t tiles = (n + SC_TILE - 1) / SC_TILE;
    sc_reduce<<<tiles, SC_THREADS, 0, st>>>(in, n, tile_sums, gate);
    rs_scan<<<1, RS_SCAN_THREADS, 0, st>>>(tile_sums, tiles, gate);
    sc_scan<<<tiles, SC_THREADS, 0, st>>>(in, n, tile_sums, out, gate);
    if (launches) *launches += 3;
}

}  // namespace gp

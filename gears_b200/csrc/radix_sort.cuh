// radix_sort.cuh — stable LSD radix sort of (key, u32 value) pairs, hand-written for sm_100a.
//
// Used for (a) the broad phase's body keys (uniform-grid cell id -> body index, north_star K2) and
// (b) ordering reported pair sets by (rank_a, rank_b).  Three kernels per 8-bit digit:
//   upsweep   : per-CTA digit histogram over the CTA's contiguous run of tiles   (reads 4 B/key)
//   scan      : exclusive scan of the [digit][cta] histogram matrix (one CTA)
//   downsweep : re-read tile, rank keys with warp match-any multisplit, stage the tile in shared
//               memory in digit order, write coalesced runs                      (R 8 B + W 8 B per key)
// Grid = a multiple of the SM count (persistent CTAs, each owning a contiguous chunk of tiles) so the
// histogram matrix stays small and every SM streams.  HBM-bound; no tensor cores (nothing to contract).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gp {

constexpr int RS_RADIX_BITS = 8;
constexpr int RS_RADIX = 1 << RS_RADIX_BITS;
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 4096 keys per tile
constexpr int RS_SCAN_THREADS = 1024;

template <typename KeyT>
__device__ __forceinline__ uint32_t rs_digit(KeyT k, int shift) {
    return (uint32_t)(k >> shift) & (RS_RADIX - 1);
}

struct RsPlan {
    uint32_t n;
    uint32_t ntiles;
    uint32_t grid;           // CTAs
    uint32_t tiles_per_cta;  // contiguous tiles owned by each CTA
};

static inline RsPlan rs_plan(uint32_t n, int sm_count) {
    RsPlan p;
    p.n = n;
    p.ntiles = (n + RS_TILE - 1) / RS_TILE;
    uint32_t max_grid = (uint32_t)sm_count * 4u;
    p.grid = p.ntiles < max_grid ? (p.ntiles ? p.ntiles : 1u) : max_grid;
    p.tiles_per_cta = (p.ntiles + p.grid - 1) / p.grid;
    if (p.tiles_per_cta == 0) p.tiles_per_cta = 1;
    return p;
}

template <typename KeyT>
__global__ void __launch_bounds__(RS_THREADS) rs_upsweep(const KeyT* __restrict__ keys, uint32_t n_max,
                                                         const uint32_t* __restrict__ n_dev, int shift,
                                                         uint32_t tiles_per_cta, uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[RS_WARPS][RS_RADIX];
    const uint32_t n = n_dev ? min(*n_dev, n_max) : n_max;  // the element count may live on the device
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * RS_RADIX; i += RS_THREADS) (&sh[0][0])[i] = 0;
    __syncthreads();
    const uint64_t begin = (uint64_t)blockIdx.x * tiles_per_cta * RS_TILE;
    uint64_t end = begin + (uint64_t)tiles_per_cta * RS_TILE;
    if (end > n) end = n;
    if (begin > n) end = begin;
    // per-warp private histograms: shared atomics never collide across warps
    for (uint64_t i = begin + threadIdx.x; i < end; i += RS_THREADS) {
        atomicAdd(&sh[warp][rs_digit(keys[i], shift)], 1u);
    }
    __syncthreads();
    for (int d = threadIdx.x; d < RS_RADIX; d += RS_THREADS) {
        uint32_t s = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) s += sh[w][d];
        hist[(size_t)d * gridDim.x + blockIdx.x] = s;
    }
}

// exclusive scan of m counters in place, one CTA
__global__ void __launch_bounds__(RS_SCAN_THREADS) rs_scan(uint32_t* __restrict__ hist, uint32_t m,
                                                           const uint32_t* __restrict__ gate = nullptr) {
    __shared__ uint32_t warp_sums[32];
    if (gate && *gate == 0) return;
    const uint32_t per = (m + RS_SCAN_THREADS - 1) / RS_SCAN_THREADS;
    const uint32_t b = threadIdx.x * per;
    const uint32_t e = (b + per < m) ? b + per : m;
    uint32_t s = 0;
    for (uint32_t i = b; i < e; ++i) s += hist[i];
    // block exclusive scan of s
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t v = warp_sums[lane];
        uint32_t vi = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, vi, o);
            if (lane >= o) vi += t;
        }
        warp_sums[lane] = vi - v;  // exclusive
    }
    __syncthreads();
    uint32_t run = warp_sums[warp] + inc - s;
    for (uint32_t i = b; i < e; ++i) {
        uint32_t v = hist[i];
        hist[i] = run;
        run += v;
    }
}

template <typename KeyT, bool IOTA>
__global__ void __launch_bounds__(RS_THREADS)
    rs_downsweep(const KeyT* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, KeyT* __restrict__ keys_out,
                 uint32_t* __restrict__ vals_out, uint32_t n_max, const uint32_t* __restrict__ n_dev, int shift,
                 uint32_t tiles_per_cta, const uint32_t* __restrict__ hist) {
    const uint32_t n = n_dev ? min(*n_dev, n_max) : n_max;
    extern __shared__ __align__(16) unsigned char rs_smem[];
    KeyT* skeys = reinterpret_cast<KeyT*>(rs_smem);
    uint32_t* svals = reinterpret_cast<uint32_t*>(rs_smem + sizeof(KeyT) * RS_TILE);
    __shared__ uint32_t cnt[RS_WARPS][RS_RADIX];
    __shared__ uint32_t dstart[RS_RADIX];
    __shared__ uint32_t gdelta[RS_RADIX];
    __shared__ uint32_t wsum[RS_WARPS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t gbase = 0;  // thread d holds the running global offset of digit d for this CTA
    if (tid < RS_RADIX) gbase = hist[(size_t)tid * gridDim.x + blockIdx.x];

    const uint32_t tile0 = blockIdx.x * tiles_per_cta;
    const uint32_t ntiles = (n + RS_TILE - 1) / RS_TILE;
    for (uint32_t t = tile0; t < tile0 + tiles_per_cta && t < ntiles; ++t) {
        const uint64_t base = (uint64_t)t * RS_TILE;
        const uint32_t valid = (n - base < (uint64_t)RS_TILE) ? (uint32_t)(n - base) : (uint32_t)RS_TILE;
        KeyT k[RS_ITEMS];
        uint32_t v[RS_ITEMS];
        uint32_t rnk[RS_ITEMS];
        for (int i = tid; i < RS_WARPS * RS_RADIX; i += RS_THREADS) (&cnt[0][0])[i] = 0;
        // warp-striped load: item i of lane l sits at warp*32*ITEMS + i*32 + l (coalesced, order-preserving)
#pragma unroll
        for (int i = 0; i < RS_ITEMS; ++i) {
            const uint32_t p = warp * (32 * RS_ITEMS) + i * 32 + lane;
            if (p < valid) {
                k[i] = keys_in[base + p];
                v[i] = IOTA ? (uint32_t)(base + p) : vals_in[base + p];
            } else {
                k[i] = ~(KeyT)0;  // pads sort to the very end of the tile
                v[i] = 0;
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < RS_ITEMS; ++i) {
            const uint32_t d = rs_digit(k[i], shift);
            const uint32_t peers = __match_any_sync(0xffffffffu, d);
            const int leader = __ffs(peers) - 1;
            uint32_t old = 0;
            if (lane == leader) {
                old = cnt[warp][d];
                cnt[warp][d] = old + __popc(peers);
            }
            old = __shfl_sync(0xffffffffu, old, leader);
            rnk[i] = old + __popc(peers & lt_mask);
            __syncwarp();
        }
        __syncthreads();
        // per digit: exclusive prefix over warps, then exclusive scan over digits
        uint32_t total = 0;
        if (tid < RS_RADIX) {
#pragma unroll
            for (int w = 0; w < RS_WARPS; ++w) {
                uint32_t c = cnt[w][tid];
                cnt[w][tid] = total;
                total += c;
            }
        }
        uint32_t inc = total;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t tt = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += tt;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        if (tid < RS_RADIX) {
            uint32_t off = 0;
#pragma unroll
            for (int w = 0; w < RS_WARPS; ++w)
                if (w < warp) off += wsum[w];
            const uint32_t ds = off + inc - total;
            dstart[tid] = ds;
            gdelta[tid] = gbase - ds;  // global index = gdelta[d] + position in tile
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < RS_ITEMS; ++i) {
            const uint32_t d = rs_digit(k[i], shift);
            const uint32_t p = dstart[d] + cnt[warp][d] + rnk[i];
            skeys[p] = k[i];
            svals[p] = v[i];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < RS_ITEMS; ++i) {
            const uint32_t p = i * RS_THREADS + tid;
            if (p < valid) {
                const KeyT kk = skeys[p];
                const uint32_t g = gdelta[rs_digit(kk, shift)] + p;
                keys_out[g] = kk;
                vals_out[g] = svals[p];
            }
        }
        if (tid < RS_RADIX) {
            // pads were counted in the last digit; they are never written, so do not advance past them
            uint32_t adv = total;
            if (tid == RS_RADIX - 1 && valid < RS_TILE) adv -= (RS_TILE - valid);
            gbase += adv;
        }
        __syncthreads();
    }
}

template <typename KeyT>
static inline size_t rs_smem_bytes() {
    return (sizeof(KeyT) + sizeof(uint32_t)) * (size_t)RS_TILE;
}

// Temp storage: hist = RS_RADIX * grid u32.  Sorts `key_bits` low bits of the first min(n, *n_dev) elements
// (n = host upper bound that sizes the grid; n_dev = optional device-resident count).  Ping-pongs between (k0,v0)
// and (k1,v1); returns 0 if the result is in (k0,v0), 1 if in (k1,v1).  If iota_first, the values of
// the first pass are the element indices (v0 is not read).  *launches is incremented per kernel.
template <typename KeyT>
static inline int rs_sort(KeyT* k0, uint32_t* v0, KeyT* k1, uint32_t* v1, uint32_t n, int key_bits, bool iota_first,
                          uint32_t* hist, int sm_count, cudaStream_t st, uint64_t* launches,
                          const uint32_t* n_dev = nullptr) {
    if (n == 0) return 0;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(rs_downsweep<KeyT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)rs_smem_bytes<KeyT>());
        cudaFuncSetAttribute(rs_downsweep<KeyT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)rs_smem_bytes<KeyT>());
        attr_set = true;
    }
    const RsPlan p = rs_plan(n, sm_count);
    int cur = 0;
    if (key_bits < 1) key_bits = 1;
    for (int shift = 0; shift < key_bits; shift += RS_RADIX_BITS) {
        const KeyT* ki = cur ? k1 : k0;
        const uint32_t* vi = cur ? v1 : v0;
        KeyT* ko = cur ? k0 : k1;
        uint32_t* vo = cur ? v0 : v1;
        rs_upsweep<KeyT><<<p.grid, RS_THREADS, 0, st>>>(ki, n, n_dev, shift, p.tiles_per_cta, hist);
        rs_scan<<<1, RS_SCAN_THREADS, 0, st>>>(hist, RS_RADIX * p.grid);
        if (iota_first && shift == 0)
            rs_downsweep<KeyT, true>
                <<<p.grid, RS_THREADS, rs_smem_bytes<KeyT>(), st>>>(ki, vi, ko, vo, n, n_dev, shift, p.tiles_per_cta, hist);
        else
            rs_downsweep<KeyT, false>
                <<<p.grid, RS_THREADS, rs_smem_bytes<KeyT>(), st>>>(ki, vi, ko, vo, n, n_dev, shift, p.tiles_per_cta, hist);
        if (launches) *launches += 3;
        cur ^= 1;
    }
    return cur;
}

}  // namespace gp

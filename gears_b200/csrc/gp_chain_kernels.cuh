// gp_chain_kernels.cuh — K5 contact chains in the reference's pair order.
// Part of the physics step; see gp_kernels.cuh for the map of kernels to reference lines.
#pragma once
#include "gp_common.cuh"

namespace gp {

// ---------------------------------------------------------------------------------------------
// K5: contact chains.  offs = exclusive scan of chain_cnt (host launches scan.cuh).  k_fill_chains
// drops, for every pair and dynamic endpoint x, the entry (partner rank << 32 | side << 31 | pair id) into x's
// segment; k_sort_chains orders each segment by partner rank = the reference's visiting order for x,
// links every pair to its successor on x's side (pair record: next_a / next_b) and counts, per pair, on how
// many chains it is not yet at the head (pending).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    k_fill_chains(const uint4* __restrict__ pairs, uint32_t pair_cap, const uint32_t* __restrict__ offs,
                  uint32_t* __restrict__ chain_cnt, unsigned long long* __restrict__ adj, const Ctl* ctl,
                  const uint32_t* __restrict__ gate) {
    if (gate && *gate == 0) return;
    const uint32_t np = min(ctl->n_pairs, pair_cap);
    for (uint32_t pid = blockIdx.x * blockDim.x + threadIdx.x; pid < np; pid += gridDim.x * blockDim.x) {
        const uint4 r = pairs[2 * (size_t)pid];
        const uint32_t ra = r.z & ~TOPBIT, rb = r.w & ~TOPBIT;
        if (!(r.x & TOPBIT)) {
            const uint32_t a = r.x;
            const uint32_t s = offs[a] + atomicSub(&chain_cnt[a], 1u) - 1u;
            adj[s] = ((unsigned long long)rb << 32) | pid;  // side 0: this body is the pair's a
        }
        if (!(r.y & TOPBIT)) {
            const uint32_t b = r.y;
            const uint32_t s = offs[b] + atomicSub(&chain_cnt[b], 1u) - 1u;
            adj[s] = ((unsigned long long)ra << 32) | TOPBIT | pid;  // side 1
        }
    }
}

constexpr uint32_t CHAIN_INLINE = 48;  // longer chains go to the block-wide sorter

// entries [first, d) of a sorted chain, stride apart: entry i-1 gets entry i as its successor on this body's side
__device__ __forceinline__ void link_chain(const unsigned long long* __restrict__ seg, uint32_t first, uint32_t d,
                                           uint32_t stride, uint4* __restrict__ pairs) {
    for (uint32_t i = first + 1; i < d; i += stride) {
        const uint32_t prev = (uint32_t)seg[i - 1], cur = (uint32_t)seg[i] & ~TOPBIT;
        pair_link(pairs, prev & ~TOPBIT)[prev >> 31] = cur;
        atomicAdd(&pair_link(pairs, cur)[2], 1u);
    }
}

__global__ void __launch_bounds__(256)
    k_sort_chains(const uint32_t* __restrict__ offs, unsigned long long* __restrict__ adj,
                  uint4* __restrict__ pairs, uint32_t* __restrict__ heavy, uint32_t heavy_cap, Ctl* ctl,
                  const uint32_t* __restrict__ gate) {
    if (gate && *gate == 0) return;
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= ctl->n) return;
    const uint32_t b = offs[x], e = offs[x + 1];
    const uint32_t d = e - b;
    if (d == 0) return;
    if (d > CHAIN_INLINE) {
        const uint32_t s = atomicAdd(&ctl->n_heavy, 1u);
        if (s < heavy_cap)
            heavy[s] = x;
        else
            atomicOr(&ctl->step_flags, SF_HEAVY_OVERFLOW);
        return;
    }
    // insertion sort in place (segments are short and L1-resident)
    for (uint32_t i = b + 1; i < e; ++i) {
        const unsigned long long v = adj[i];
        uint32_t j = i;
        while (j > b && adj[j - 1] > v) {
            adj[j] = adj[j - 1];
            --j;
        }
        adj[j] = v;
    }
    link_chain(adj + b, 0u, d, 1u, pairs);
}

// A chain longer than CHAIN_INLINE is sorted by ONE WARP: bitonic network in its all-ascending form (the first
// sub-step of each stage mirrors the partner: l = i ^ (k-1), the rest use l = i ^ j), so every compare-exchange moves
// the smaller key to the lower index and the virtual +inf padding at [d, m) never has to move.  Up to WSORT_SMEM
// entries the chain is staged in the warp's shared-memory buffer; longer chains are sorted in place in global memory
// (through L2: the lanes exchange data there).  Then the warp links the chain (successor pointers + pending counts).
constexpr uint32_t WSORT_SMEM = 128;
__device__ __forceinline__ void warp_sort_and_link(unsigned long long* seg, uint32_t d, unsigned long long* wbuf,
                                                   uint4* __restrict__ pairs) {
    const uint32_t lane = threadIdx.x & 31;
    uint32_t m = 1;
    while (m < d) m <<= 1;
    if (d <= WSORT_SMEM) {
        for (uint32_t i = lane; i < d; i += 32) wbuf[i] = __ldcg(&seg[i]);
        __syncwarp();
        for (uint32_t k = 2; k <= m; k <<= 1)
            for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                const uint32_t mask = (j == (k >> 1)) ? (k - 1u) : j;
                for (uint32_t i = lane; i < m; i += 32) {
                    const uint32_t l = i ^ mask;
                    if (l > i && l < d) {
                        const unsigned long long vi = wbuf[i], vl = wbuf[l];
                        if (vi > vl) wbuf[i] = vl, wbuf[l] = vi;
                    }
                }
                __syncwarp();
            }
        for (uint32_t i = lane; i < d; i += 32) __stcg(&seg[i], wbuf[i]);
        for (uint32_t i = lane + 1; i < d; i += 32) {
            const uint32_t prev = (uint32_t)wbuf[i - 1], cur = (uint32_t)wbuf[i] & ~TOPBIT;
            pair_link(pairs, prev & ~TOPBIT)[prev >> 31] = cur;
            atomicAdd(&pair_link(pairs, cur)[2], 1u);
        }
        __syncwarp();
        return;
    }
    for (uint32_t k = 2; k <= m; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            const uint32_t mask = (j == (k >> 1)) ? (k - 1u) : j;
            for (uint32_t i = lane; i < m; i += 32) {
                const uint32_t l = i ^ mask;
                if (l > i && l < d) {
                    const unsigned long long vi = __ldcg(&seg[i]), vl = __ldcg(&seg[l]);
                    if (vi > vl) __stcg(&seg[i], vl), __stcg(&seg[l], vi);
                }
            }
            __syncwarp();
        }
    for (uint32_t i = lane + 1; i < d; i += 32) {
        const uint32_t prev = (uint32_t)__ldcg(&seg[i - 1]), cur = (uint32_t)__ldcg(&seg[i]) & ~TOPBIT;
        pair_link(pairs, prev & ~TOPBIT)[prev >> 31] = cur;
        atomicAdd(&pair_link(pairs, cur)[2], 1u);
    }
    __syncwarp();
}

// one warp per heavy body (a dense pile makes EVERY chain heavy: a million of them)
__global__ void __launch_bounds__(256)
    k_sort_heavy(const uint32_t* __restrict__ offs, unsigned long long* __restrict__ adj, uint4* __restrict__ pairs,
                 const uint32_t* __restrict__ heavy, uint32_t heavy_cap, const Ctl* ctl,
                 const uint32_t* __restrict__ gate) {
    __shared__ unsigned long long wbuf[8][WSORT_SMEM];
    if (gate && *gate == 0) return;
    const uint32_t nh = min(ctl->n_heavy, heavy_cap);
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t hi = warp; hi < nh; hi += nwarps) {
        const uint32_t x = heavy[hi];
        const uint32_t b = offs[x], d = offs[x + 1] - b;
        warp_sort_and_link(adj + b, d, wbuf[threadIdx.x >> 5], pairs);
    }
}


}  // namespace gp

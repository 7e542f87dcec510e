// gp_pairfind_kernels.cuh — K4 pair find: warp-autonomous tiles, 256-bit AABB record loads, per-warp staged bursts.
// Part of the physics step; see gp_kernels.cuh for the map of kernels to reference lines.
#pragma once
#include "gp_common.cuh"

namespace gp {

// ---------------------------------------------------------------------------------------------
// K4: broad phase + exact snapshot test.  One thread per sorted body p; candidates are the bodies
// at later sorted positions in the 13 forward neighbour cells + the rest of the own cell (5
// contiguous row ranges through LB) + the big list.  Each unordered pair is found exactly once.
// A pair is a CANDIDATE if the AABBs inflated by the margin m overlap (superset of every pair that
// can intersect at its turn while no AABB moves further than m, verified in K6); it carries the
// S_snapshot flag if the un-inflated AABBs intersect (physics.rs:159-164, strict).
// Pairs are appended with warp-aggregated atomics (one atomicAdd per warp-wide ballot).
// Pair record: x = a | static<<31, y = b | static<<31, z = rank_a | snapshot<<31, w = rank_b (| hit<<31 in K6).
// ---------------------------------------------------------------------------------------------
// AABB record of the step (sbox, 32 B per slot): lo = (min.xyz, +-m) where m = fl(mu * extent) is the body's
// OUTER margin and the sign bit is the static flag; hi = (max.xyz, rank).  K3 writes it, so every kernel of the
// step (K4, k_requery, k_verify_watch) sees the same margin bits.
__device__ __forceinline__ float box_margin(const float4 lo) { return fabsf(lo.w); }
__device__ __forceinline__ uint32_t box_static(const float4 lo) { return __float_as_uint(lo.w) & TOPBIT; }

__device__ __forceinline__ bool inflated_overlap(const float4 alo, const float4 ahi, const float4 blo, const float4 bhi,
                                                 float m2) {
    return (alo.x - bhi.x < m2) && (blo.x - ahi.x < m2) && (alo.y - bhi.y < m2) && (blo.y - ahi.y < m2) &&
           (alo.z - bhi.z < m2) && (blo.z - ahi.z < m2);
}
// the largest of the six separations: every comparison above is `separation < m2` (equivalent for non-NaN boxes;
// a NaN box may only ADD a candidate, and candidates are re-tested exactly in the sweep)
__device__ __forceinline__ float max_separation(const float4 alo, const float4 ahi, const float4 blo, const float4 bhi) {
    return fmaxf(fmaxf(fmaxf(alo.x - bhi.x, blo.x - ahi.x), fmaxf(alo.y - bhi.y, blo.y - ahi.y)),
                 fmaxf(alo.z - bhi.z, blo.z - ahi.z));
}

// K4.  A WARP owns a tile of 32 consecutive slots (= bodies in cell order):
//   1. each lane loads its AABB record with ONE 256-bit load (LDG.E.ENL2.256, new on sm_100; kept in registers and
//      in the warp's shared-memory tile), its cell key and the 10 cell-table entries that bound its 5 candidate runs
//      (the rest of its own cell + the 13 forward neighbour cells are 5 contiguous runs of slots): every unordered
//      pair is found exactly once;
//   2. one warp scan of the per-lane test counts gives every lane its offset in the tile's flattened test list;
//      each lane writes its (candidate slot, tile-local body) entries there -- a loop of ~4 short iterations;
//   3. the warp walks the list, two tests per lane per iteration, candidate records fetched back to back;
//   4. hits are staged per warp in shared memory and leave in bursts with ONE global atomic per flush.
// No CTA barrier, no binary search.  Lists longer than WP_LIST are processed in windows; a lane with thousands of
// candidates (a degenerate cell) switches the tile to a body-at-a-time walk.
// Measured on uniform_16M (profiles/r02): 0.69 ms against 1.07 ms for the round-1 kernel, which flattened the tests of
// a 256-body tile over the CTA behind a TMA bulk copy (two block scans, a range compaction, a binary search per four
// tests and four CTA barriers per tile); the 128-bit-load variant of this kernel took 0.73 ms.  Both were deleted.
constexpr int WP_RANGES = 5;     // candidate slot ranges per body: 5 cell-row runs
constexpr int WP_THREADS = 128;
constexpr int WP_WARPS = WP_THREADS / 32;
constexpr int WP_LIST = 384;    // flattened tests per window
#ifndef GP_WP_PER_LANE
#define GP_WP_PER_LANE 2
#endif
constexpr int WP_PER_LANE = GP_WP_PER_LANE;  // tests per lane and iteration (their candidate records are in flight together)
constexpr int WP_STAGE = 32 + 32 * WP_PER_LANE;  // staged records per warp: flushed above 32, one iteration adds 32 per test
constexpr uint32_t WP_DENSE = 4096;  // a lane with more candidates than this: body-at-a-time walk

struct WarpPairSmem {
    float4 abox[64];            // the tile's 32 AABB records (lo, hi)
    uint4 stage[WP_STAGE];      // sweep-pair records waiting for the next flush
    uint2 wstage[WP_STAGE];     // watch-list entries
    uint32_t q[WP_LIST];        // candidate slot of test t of the window
    uint8_t j[WP_LIST];         // tile-local body of test t
    uint32_t stage_cnt, wstage_cnt, pad[2];
};

// one AABB record (lo, hi): 32 B, 32 B-aligned, read-only during K4
__device__ __forceinline__ void ld_box(const float4* __restrict__ sbox, uint32_t q, float4& lo, float4& hi) {
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
             : "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w), "=f"(hi.x), "=f"(hi.y), "=f"(hi.z), "=f"(hi.w)
             : "l"(sbox + 2 * (size_t)q));
}

__global__ void __launch_bounds__(WP_THREADS)
    k_find_pairs(const float4* __restrict__ sbox, const uint32_t* __restrict__ LB, uint4* __restrict__ pairs,
                      uint2* __restrict__ watch, uint32_t* __restrict__ chain_cnt, uint32_t pair_cap, Ctl* ctl) {
    __shared__ __align__(16) WarpPairSmem smem[WP_WARPS];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpPairSmem& sm = smem[warp];
    const uint32_t n = ctl->n, n_small = ctl->n_small;
    const GridParams g = ctl->grid;
    const float in_ratio = ctl->margin > 0.0f ? ctl->margin_in / ctl->margin : 1.0f;
    const uint32_t row = (uint32_t)g.dx, slab = (uint32_t)g.dx * (uint32_t)g.dy;
    if (lane == 0) sm.stage_cnt = 0, sm.wstage_cnt = 0;
    __syncwarp();
    uint32_t nsnap = 0;

    auto flush = [&]() {  // warp-synchronous
        __syncwarp();
        const uint32_t c = sm.stage_cnt;
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&ctl->n_pairs, c);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (lane == 0 && base + c > pair_cap) atomicOr(&ctl->step_flags, SF_PAIR_OVERFLOW);
        for (uint32_t i = lane; i < c; i += 32) {
            const uint32_t slot = base + i;
            if (slot < pair_cap) {
                const uint4 rec = sm.stage[i];
                pairs[2 * (size_t)slot] = rec;
                pairs[2 * (size_t)slot + 1] = link_init();
                if (!(rec.x & TOPBIT)) atomicAdd(&chain_cnt[rec.x], 1u);
                if (!(rec.y & TOPBIT)) atomicAdd(&chain_cnt[rec.y & ~TOPBIT], 1u);
                nsnap += rec.z >> 31;
            }
        }
        __syncwarp();
        if (lane == 0) sm.stage_cnt = 0;
        __syncwarp();
    };
    auto flush_watch = [&]() {  // warp-synchronous
        __syncwarp();
        const uint32_t c = sm.wstage_cnt;
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&ctl->n_watch, c);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (lane == 0 && base + c > pair_cap) atomicOr(&ctl->step_flags, SF_PAIR_OVERFLOW);
        for (uint32_t i = lane; i < c; i += 32)
            if (base + i < pair_cap) watch[base + i] = sm.wstage[i];
        __syncwarp();
        if (lane == 0) sm.wstage_cnt = 0;
        __syncwarp();
    };
    auto flush_if_needed = [&]() {  // warp-synchronous, uniform
        __syncwarp();
        if (sm.stage_cnt > 32u) flush();
        if (sm.wstage_cnt > 32u) flush_watch();
    };
    // one candidate: a = tile body (slot pa), b = slot q.  Same predicate and record as k_find_pairs.
    auto test = [&](const float4 alo, const float4 ahi, uint32_t pa, const float4 blo, const float4 bhi, uint32_t q) {
        const float sep = max_separation(alo, ahi, blo, bhi);
        const float m2 = box_margin(alo) + box_margin(blo);
        if (!(sep < m2)) return;
        const uint32_t sa = box_static(alo), sb = box_static(blo);
        if (sa & sb) return;  // static-static: no-op in the reference (physics.rs:225)
        const uint32_t ra = __float_as_uint(ahi.w), rb = __float_as_uint(bhi.w);
        const bool a_first = ra < rb;  // pair order = the reference's iteration order
        const uint32_t x = a_first ? (pa | sa) : (q | sb), y = a_first ? (q | sb) : (pa | sa);
        if (sep < m2 * in_ratio) {
            const bool snap = alo.x < bhi.x && ahi.x > blo.x && alo.y < bhi.y && ahi.y > blo.y && alo.z < bhi.z &&
                              ahi.z > blo.z;  // physics.rs:159-164 on the snapshot (strict): the S_snapshot flag
            sm.stage[atomicAdd(&sm.stage_cnt, 1u)] =
                make_uint4(x, y, (a_first ? ra : rb) | (snap ? TOPBIT : 0u), a_first ? rb : ra);
        } else {
            sm.wstage[atomicAdd(&sm.wstage_cnt, 1u)] = make_uint2(x, y);
        }
    };

    const uint32_t ntiles = (n + 31u) / 32u;
    const uint32_t gw = blockIdx.x * WP_WARPS + warp, nw = gridDim.x * WP_WARPS;
    for (uint32_t tile = gw; tile < ntiles; tile += nw) {
        const uint32_t p0 = tile * 32u, p = p0 + lane;
        const bool live = p < n;
        float4 alo = make_float4(0.f, 0.f, 0.f, 0.f), ahi = alo;
        uint32_t q0[WP_RANGES], len[WP_RANGES];
#pragma unroll
        for (int r = 0; r < WP_RANGES; ++r) q0[r] = 0, len[r] = 0;
        if (live) {
            ld_box(sbox, p, alo, ahi);
            if (p < n_small) {
                // same expression as body_key() in K1: bit-identical cell
                const uint32_t k = cell_key(g, (alo.x + ahi.x) * 0.5f, (alo.y + ahi.y) * 0.5f, (alo.z + ahi.z) * 0.5f);
                const uint32_t e0 = LB[k + 2];
                const uint32_t b1 = LB[k + row - 1], e1 = LB[k + row + 2];
                const uint32_t b2 = LB[k + slab - row - 1], e2 = LB[k + slab - row + 2];
                const uint32_t b3 = LB[k + slab - 1], e3 = LB[k + slab + 2];
                const uint32_t b4 = LB[k + slab + row - 1], e4 = LB[k + slab + row + 2];
                q0[0] = p + 1, len[0] = e0 > p + 1 ? e0 - (p + 1) : 0u;
                q0[1] = b1, len[1] = e1 - b1;
                q0[2] = b2, len[2] = e2 - b2;
                q0[3] = b3, len[3] = e3 - b3;
                q0[4] = b4, len[4] = e4 - b4;
            }
        }
        sm.abox[2 * lane] = alo, sm.abox[2 * lane + 1] = ahi;
        // big list (a floor, walls): every body tests the big bodies behind it, directly.  Warp-uniform loop.
        for (uint32_t q = n_small; q < n; ++q) {
            if (live && q > p) test(alo, ahi, p, sbox[2 * (size_t)q], sbox[2 * (size_t)q + 1], q);
            flush_if_needed();
        }
        uint32_t cnt = 0;
        bool huge = false;
#pragma unroll
        for (int r = 0; r < WP_RANGES; ++r) huge |= len[r] > WP_DENSE, cnt += len[r];
        __syncwarp();  // abox is complete
        if (__any_sync(0xffffffffu, huge)) {
            // a degenerate cell (thousands of bodies): the 32 lanes share the ranges of one body at a time
#pragma unroll
            for (int r = 0; r < WP_RANGES; ++r)
                for (int L = 0; L < 32; ++L) {
                    const uint32_t bq0 = __shfl_sync(0xffffffffu, q0[r], L), bln = __shfl_sync(0xffffffffu, len[r], L);
                    if (bln == 0) continue;  // uniform across the warp
                    const float4 blo_a = sm.abox[2 * L], bhi_a = sm.abox[2 * L + 1];
                    for (uint32_t o = 0; o < bln; o += 32) {
                        const uint32_t q = bq0 + o + lane;
                        if (o + lane < bln) test(blo_a, bhi_a, p0 + L, sbox[2 * (size_t)q], sbox[2 * (size_t)q + 1], q);
                        flush_if_needed();
                    }
                }
            __syncwarp();  // abox is rewritten by the next tile
            continue;
        }
        // offsets of the lanes' tests in the tile's flattened list (cnt <= 5 * WP_DENSE per lane: no overflow)
        uint32_t inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if ((int)lane >= o) inc += t;
        }
        const uint32_t off = inc - cnt, Tw = __shfl_sync(0xffffffffu, inc, 31);
        for (uint32_t base = 0; base < Tw; base += WP_LIST) {
            const uint32_t wend = min(base + (uint32_t)WP_LIST, Tw);
            {   // my entries that fall into the window [base, wend)
                uint32_t pre = off;
#pragma unroll
                for (int r = 0; r < WP_RANGES; ++r) {
                    const uint32_t lo = max(pre, base), hi = min(pre + len[r], wend);
                    for (uint32_t t = lo; t < hi; ++t) {
                        sm.q[t - base] = q0[r] + (t - pre);
                        sm.j[t - base] = (uint8_t)lane;
                    }
                    pre += len[r];
                }
            }
            __syncwarp();
            const uint32_t nt = wend - base;
            for (uint32_t t0 = 0; t0 < nt; t0 += 32 * WP_PER_LANE) {
                // (entry 0 exists: nt > 0) all loads of the iteration first, then the tests
                uint32_t qs[WP_PER_LANE], js[WP_PER_LANE];
                bool vs[WP_PER_LANE];
                float4 bl[WP_PER_LANE], bh[WP_PER_LANE];
#pragma unroll
                for (int u = 0; u < WP_PER_LANE; ++u) {
                    const uint32_t t = t0 + 32 * u + lane;
                    vs[u] = t < nt;
                    qs[u] = sm.q[vs[u] ? t : 0u], js[u] = sm.j[vs[u] ? t : 0u];
                }
#pragma unroll
                for (int u = 0; u < WP_PER_LANE; ++u) ld_box(sbox, qs[u], bl[u], bh[u]);
#pragma unroll
                for (int u = 0; u < WP_PER_LANE; ++u)
                    if (vs[u]) test(sm.abox[2 * js[u]], sm.abox[2 * js[u] + 1], p0 + js[u], bl[u], bh[u], qs[u]);
                flush_if_needed();
            }
            __syncwarp();  // the list is rewritten by the next window
        }
        __syncwarp();  // abox is rewritten by the next tile
    }
    if (sm.stage_cnt) flush();
    if (sm.wstage_cnt) flush_watch();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nsnap += __shfl_xor_sync(0xffffffffu, nsnap, o);
    if (lane == 0 && nsnap) atomicAdd(&ctl->n_snapshot, nsnap);
}

}  // namespace gp

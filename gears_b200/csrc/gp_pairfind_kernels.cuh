// gp_pairfind_kernels.cuh — K4 pair find: TMA-staged tiles, flattened tests, per-warp staged bursts.
// Part of the physics step; see gp_kernels.cuh for the map of kernels to reference lines.
#pragma once
#include "gp_common.cuh"
#include "scan.cuh"

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

constexpr int FP_THREADS = 256;   // bodies per tile = threads per CTA
constexpr int FP_RANGES = 5;      // candidate slot ranges per body in the flattened walk: 5 cell-row runs
constexpr int FP_BATCH = 4;       // tests per thread per batch
constexpr int FP_WARPS = FP_THREADS / 32;
constexpr int FP_WSTAGE = 160;    // staged pair records per warp (flushed above 32: a batch adds <= 128)
constexpr int FP_NR = FP_THREADS * FP_RANGES;

// ---- TMA bulk copy (cp.async.bulk, 1-D) of a tile of AABBs into shared memory, completion on an mbarrier ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// bounded wait: a kernel of this library never spins forever (returns false after ~2^26 polls)
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return true;
    }
    return false;
}

struct FindPairsSmem {
    float4 abox[2 * FP_THREADS];  // AABB records of the tile's bodies, staged by TMA
    uint32_t rstart[FP_NR];       // NON-EMPTY candidate ranges of the tile, compacted: first candidate slot,
    uint32_t roff[FP_NR + 1];     //   exclusive scan of their lengths (test t belongs to range id, roff[id] <= t),
    uint8_t rbody[FP_NR];         //   and the tile-local body the range belongs to
    uint4 stage[FP_WARPS][FP_WSTAGE];  // per-warp sweep-pair records waiting for the next flush
    uint2 wstage[FP_WARPS][FP_WSTAGE]; // per-warp watch-list entries
    uint32_t scan[33];
    uint32_t stage_cnt[FP_WARPS], wstage_cnt[FP_WARPS];
    uint64_t bar;
};

// K4.  Persistent CTAs; a tile is 256 consecutive slots (= bodies in cell order).  Per tile:
//  1. one thread issues a TMA bulk copy of the tile's 8 KB of AABB records into shared memory; meanwhile every
//     thread fetches the 10 cell-table entries of its body (independent loads) and derives its 6 candidate ranges:
//     the rest of its own cell + the 13 forward neighbour cells are 5 contiguous runs of slots, + the big list;
//  2. two block scans compact the non-empty ranges and flatten the tile's tests into [0, T);
//  3. the tests are dealt out evenly, FP_BATCH consecutive tests per thread per batch (one binary search per
//     batch, then a walk): every lane is busy whatever the cell populations are; the candidates' records of a
//     batch are fetched back to back before any is tested;
//  4. hits are staged per warp in shared memory and leave in bursts with ONE global atomic per flush; the warps
//     of a CTA do not synchronise while they test.
// A candidate (separation < m_a + m_b) within the inner margin becomes a sweep pair
//   (a | static<<31, b | static<<31, rank_a | snapshot<<31, rank_b), a = earlier in iteration order,
// the others go to the watch list.  S_snapshot flag: the exact test physics.rs:159-164 on the snapshot.
__global__ void __launch_bounds__(FP_THREADS)
    k_find_pairs(const float4* __restrict__ sbox, const uint32_t* __restrict__ LB, uint4* __restrict__ pairs,
                 uint2* __restrict__ watch, uint32_t* __restrict__ chain_cnt, uint32_t pair_cap, Ctl* ctl) {
    extern __shared__ __align__(128) unsigned char fp_smem_raw[];
    FindPairsSmem& sm = *reinterpret_cast<FindPairsSmem*>(fp_smem_raw);
    const uint32_t tid = threadIdx.x;
    const uint32_t n = ctl->n;
    const GridParams g = ctl->grid;
    const float in_ratio = ctl->margin > 0.0f ? ctl->margin_in / ctl->margin : 1.0f;
    const uint32_t n_small = ctl->n_small;
    const uint32_t row = (uint32_t)g.dx, slab = (uint32_t)g.dx * (uint32_t)g.dy;
    const uint32_t lane = tid & 31, warp = tid >> 5;
    if (tid == 0) mbar_init(&sm.bar, 1);
    if (lane == 0) sm.stage_cnt[warp] = 0, sm.wstage_cnt[warp] = 0;
    __syncthreads();
    uint32_t parity = 0, nsnap = 0;
    const uint32_t ntiles = (n + FP_THREADS - 1) / FP_THREADS;
    uint4* const wstage = sm.stage[warp];
    uint32_t* const wcnt = &sm.stage_cnt[warp];
    uint2* const wwatch = sm.wstage[warp];
    uint32_t* const wwcnt = &sm.wstage_cnt[warp];

    auto flush = [&]() {  // warp-synchronous
        __syncwarp();
        const uint32_t c = *wcnt;
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&ctl->n_pairs, c);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (lane == 0 && base + c > pair_cap) atomicOr(&ctl->step_flags, SF_PAIR_OVERFLOW);
        for (uint32_t i = lane; i < c; i += 32) {
            const uint32_t slot = base + i;
            if (slot < pair_cap) {
                const uint4 rec = wstage[i];
                pairs[2 * (size_t)slot] = rec;
                pairs[2 * (size_t)slot + 1] = link_init();
                if (!(rec.x & TOPBIT)) atomicAdd(&chain_cnt[rec.x], 1u);
                if (!(rec.y & TOPBIT)) atomicAdd(&chain_cnt[rec.y & ~TOPBIT], 1u);
                nsnap += rec.z >> 31;
            }
        }
        __syncwarp();
        if (lane == 0) *wcnt = 0;
        __syncwarp();
    };
    auto flush_watch = [&]() {  // warp-synchronous
        __syncwarp();
        const uint32_t c = *wwcnt;
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&ctl->n_watch, c);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (lane == 0 && base + c > pair_cap) atomicOr(&ctl->step_flags, SF_PAIR_OVERFLOW);
        for (uint32_t i = lane; i < c; i += 32)
            if (base + i < pair_cap) watch[base + i] = wwatch[i];
        __syncwarp();
        if (lane == 0) *wwcnt = 0;
        __syncwarp();
    };
    // one candidate: a = tile body (record from shared memory, slot pa), b = slot q
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
            // physics.rs:159-164 on the snapshot (strict): the S_snapshot flag
            const bool snap = alo.x < bhi.x && ahi.x > blo.x && alo.y < bhi.y && ahi.y > blo.y && alo.z < bhi.z &&
                              ahi.z > blo.z;
            wstage[atomicAdd(wcnt, 1u)] =
                make_uint4(x, y, (a_first ? ra : rb) | (snap ? TOPBIT : 0u), a_first ? rb : ra);
        } else {
            wwatch[atomicAdd(wwcnt, 1u)] = make_uint2(x, y);
        }
    };

    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint32_t p0 = tile * FP_THREADS;
        const uint32_t valid = min((uint32_t)FP_THREADS, n - p0);
        if (tid == 0) tma_load_1d(sm.abox, sbox + 2 * (size_t)p0, valid * 32u, &sm.bar);
        const uint32_t p = p0 + tid;
        const bool live = tid < valid;
        uint32_t q0[FP_RANGES], len[FP_RANGES];
#pragma unroll
        for (int r = 0; r < FP_RANGES; ++r) q0[r] = 0, len[r] = 0;
        if (live) {
            if (p < n_small) {
                // the AABB is needed for the cell: read it straight from global (the TMA copy of the same lines is
                // in flight; one L2 fill serves both).  Same expression as body_key() in K1: bit-identical cell.
                const float4 alo = sbox[2 * (size_t)p], ahi = sbox[2 * (size_t)p + 1];
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
        // big list (a floor, walls: at most max_big bodies): every body tests the big bodies behind it, directly --
        // one or two tests per body do not deserve the range bookkeeping below.  Warp-uniform loop.
        {
            float4 mlo = make_float4(0, 0, 0, 0), mhi = mlo;
            if (live) mlo = sbox[2 * (size_t)p], mhi = sbox[2 * (size_t)p + 1];
            for (uint32_t q = n_small; q < n; ++q) {
                if (live && q > p) test(mlo, mhi, p, sbox[2 * (size_t)q], sbox[2 * (size_t)q + 1], q);
                __syncwarp();
                if (*wcnt > 32u) flush();
                if (*wwcnt > 32u) flush_watch();
            }
        }
        // compact the non-empty ranges (thread-major order) and scan their lengths
        uint32_t nne = 0;
        unsigned long long sum64 = 0;
#pragma unroll
        for (int r = 0; r < FP_RANGES; ++r) nne += len[r] ? 1u : 0u, sum64 += len[r];
        const bool dense = __syncthreads_or(sum64 > 0x400000ull);  // a degenerate tile (thousands of bodies per cell)
        if (dense) {
            // the flattened index would overflow: every thread walks its own ranges
            const bool tma_ok = mbar_wait(&sm.bar, parity);
            parity ^= 1u;
            if (!tma_ok && tid == 0) atomicOr(&ctl->step_flags, SF_INTERNAL);
            // warp-cooperative: the 32 lanes share the ranges of one body at a time
#pragma unroll
            for (int r = 0; r < FP_RANGES; ++r)
                for (int L = 0; L < 32; ++L) {
                    const uint32_t bq0 = __shfl_sync(0xffffffffu, q0[r], L), bln = __shfl_sync(0xffffffffu, len[r], L);
                    if (bln == 0) continue;  // uniform across the warp
                    const uint32_t j = warp * 32 + L;
                    const float4 alo = sm.abox[2 * j], ahi = sm.abox[2 * j + 1];
                    for (uint32_t o = 0; o < bln; o += 32) {
                        const uint32_t q = bq0 + o + lane;
                        if (o + lane < bln) test(alo, ahi, p0 + j, sbox[2 * (size_t)q], sbox[2 * (size_t)q + 1], q);
                        __syncwarp();
                        if (*wcnt > 32u) flush();
                        if (*wwcnt > 32u) flush_watch();
                    }
                }
            __syncthreads();
            continue;
        }
        uint32_t Rn, T;
        const uint32_t rbase = block_exclusive_scan_256(nne, &Rn, sm.scan);
        uint32_t run = block_exclusive_scan_256((uint32_t)sum64, &T, sm.scan);
        {
            uint32_t w = rbase;
#pragma unroll
            for (int r = 0; r < FP_RANGES; ++r)
                if (len[r]) {
                    sm.rstart[w] = q0[r];
                    sm.roff[w] = run;
                    sm.rbody[w] = (uint8_t)tid;
                    run += len[r];
                    ++w;
                }
            if (tid == FP_THREADS - 1) sm.roff[Rn] = run;  // = T
        }
        const bool tma_ok = mbar_wait(&sm.bar, parity);
        parity ^= 1u;
        if (!tma_ok && tid == 0) atomicOr(&ctl->step_flags, SF_INTERNAL);
        __syncthreads();

        for (uint32_t tb = warp * 32 * FP_BATCH; tb < T; tb += FP_THREADS * FP_BATCH) {
            const uint32_t t0 = tb + lane * FP_BATCH;
            if (t0 < T) {
                // range that holds test t0: largest id with roff[id] <= t0 (ranges are non-empty: roff is strictly
                // increasing)
                uint32_t lo_i = 0, hi_i = Rn;
                while (hi_i - lo_i > 1) {
                    const uint32_t mid = (lo_i + hi_i) >> 1;
                    if (sm.roff[mid] <= t0) lo_i = mid; else hi_i = mid;
                }
                uint32_t id = lo_i;
                uint32_t end_t = sm.roff[id + 1];
                uint32_t q = sm.rstart[id] + (t0 - sm.roff[id]);
                const uint32_t nt = min((uint32_t)FP_BATCH, T - t0);
                // where the tests of this batch live, then all their loads, then the tests
                uint32_t qs[FP_BATCH], js[FP_BATCH];
#pragma unroll
                for (int u = 0; u < FP_BATCH; ++u) {
                    if ((uint32_t)u < nt) {
                        if (t0 + u >= end_t) {
                            ++id;
                            end_t = sm.roff[id + 1];
                            q = sm.rstart[id];
                        }
                        qs[u] = q++;
                        js[u] = sm.rbody[id];
                    } else {
                        qs[u] = qs[0], js[u] = js[0];
                    }
                }
                float4 bl[FP_BATCH], bh[FP_BATCH];
#pragma unroll
                for (int u = 0; u < FP_BATCH; ++u) bl[u] = sbox[2 * (size_t)qs[u]], bh[u] = sbox[2 * (size_t)qs[u] + 1];
#pragma unroll
                for (int u = 0; u < FP_BATCH; ++u)
                    if ((uint32_t)u < nt)
                        test(sm.abox[2 * js[u]], sm.abox[2 * js[u] + 1], p0 + js[u], bl[u], bh[u], qs[u]);
            }
            __syncwarp();
            if (*wcnt > 32u) flush();  // the next batch adds at most 32 * FP_BATCH = 128 records
            if (*wwcnt > 32u) flush_watch();
        }
        __syncthreads();  // rstart/roff/abox are rewritten by the next tile
    }
    if (*wcnt) flush();
    if (*wwcnt) flush_watch();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nsnap += __shfl_xor_sync(0xffffffffu, nsnap, o);
    if ((tid & 31) == 0 && nsnap) atomicAdd(&ctl->n_snapshot, nsnap);
}


}  // namespace gp

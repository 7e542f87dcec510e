// gp_sweep_kernels.cuh — K6 ordered contact sweep (check_and_resolve_collision).
// Part of the physics step; see gp_kernels.cuh for the map of kernels to reference lines.
#pragma once
#include "gp_common.cuh"
#include "gp_chain_kernels.cuh"

namespace gp {

// ---------------------------------------------------------------------------------------------
// K6: the ordered contact sweep as ONE persistent cooperative kernel: a CTA per SM slot, a grid-wide
// barrier per dependency level.  A frontier holds the pairs that are at the head of both chains.
// Per pair: exact intersects with CURRENT positions (physics.rs:482-489), resolve (physics.rs:176-299),
// then pop both chains; a pair whose pending count drops to 0 joins the next frontier.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ldcg4(const float4* p) { return __ldcg(p); }

struct BodyRW {
    V3 pos, vel;
    float mass, rest;
    V3 omin, omax;
    bool is_static;
};

// physics.rs:176-299; a = earlier in iteration order.  Returns true if positions were corrected.
__device__ __forceinline__ bool resolve_pair(BodyRW& a, BodyRW& b, V3 amin, V3 amax, V3 bmin, V3 bmax) {
    const float overlap_x = fminf(amax.x, bmax.x) - fmaxf(amin.x, bmin.x);
    const float overlap_y = fminf(amax.y, bmax.y) - fmaxf(amin.y, bmin.y);
    const float overlap_z = fminf(amax.z, bmax.z) - fmaxf(amin.z, bmin.z);
    const V3 a_center = (amin + amax) * 0.5f;
    const V3 b_center = (bmin + bmax) * 0.5f;
    const V3 cd = b_center - a_center;
    float min_overlap;
    V3 n;
    if (overlap_x <= overlap_y && overlap_x <= overlap_z) {
        min_overlap = overlap_x;
        n = mk(cd.x > 0.0f ? 1.0f : -1.0f, 0.0f, 0.0f);
    } else if (overlap_y <= overlap_z) {
        min_overlap = overlap_y;
        n = mk(0.0f, cd.y > 0.0f ? 1.0f : -1.0f, 0.0f);
    } else {
        min_overlap = overlap_z;
        n = mk(0.0f, 0.0f, cd.z > 0.0f ? 1.0f : -1.0f);
    }
    const float inv_a = a.is_static ? 0.0f : 1.0f / a.mass;
    const float inv_b = b.is_static ? 0.0f : 1.0f / b.mass;
    const float tim = inv_a + inv_b;
    if (tim <= 0.0f) return false;
    const V3 rv = b.vel - a.vel;
    const float vn = dot3(rv, n);
    if (vn > 0.0f) return false;
    bool moved = false;
    if (min_overlap > POSITION_CORRECTION_SLOP) {
        const float cm = ((min_overlap - POSITION_CORRECTION_SLOP) * POSITION_CORRECTION_FACTOR) / tim;
        const V3 corr = n * cm;
        if (!a.is_static) a.pos = a.pos - corr * inv_a;
        if (!b.is_static) b.pos = b.pos + corr * inv_b;
        moved = true;
    }
    const float e = fabsf(vn) < VELOCITY_THRESHOLD ? 0.0f : (a.rest + b.rest) * 0.5f;
    const float j = ((-(1.0f + e)) * vn) / tim;
    const V3 imp = n * j;
    if (!a.is_static) a.vel = a.vel - imp * inv_a;
    if (!b.is_static) b.vel = b.vel + imp * inv_b;
    const V3 tv = rv - n * vn;
    const float ts = len3(tv);
    if (ts > 0.001f) {
        const V3 t = tv / ts;
        const float f = fminf(FRICTION_COEFFICIENT * fabsf(j), ts * tim);
        if (!a.is_static) a.vel = a.vel - (t * f) * inv_a;
        if (!b.is_static) b.vel = b.vel + (t * f) * inv_b;
    }
    if (!a.is_static && len3(a.vel) < VELOCITY_THRESHOLD * 0.5f) a.vel = mk(0.f, 0.f, 0.f);
    if (!b.is_static && len3(b.vel) < VELOCITY_THRESHOLD * 0.5f) b.vel = mk(0.f, 0.f, 0.f);
    return moved;
}

// Frontier appends are staged per CTA in shared memory and leave with ONE global atomic per flush: the
// next-frontier counter is a single address, and same-address atomics serialise in L2.
constexpr uint32_t FB_CAP = 2048;
struct FrontierStage {
    uint32_t buf[FB_CAP];
    uint32_t cnt, base;
};
__device__ __forceinline__ void stage_push(FrontierStage& fs, uint32_t q) { fs.buf[atomicAdd(&fs.cnt, 1u)] = q; }
// called by ALL threads of the CTA (contains barriers); room = pushes the next iteration may add
__device__ __forceinline__ void stage_flush(FrontierStage& fs, uint32_t* __restrict__ frontier, uint32_t* counter,
                                            uint32_t room) {
    __syncthreads();
    const uint32_t c = fs.cnt;
    __syncthreads();  // everybody has read the count before anybody pushes again
    if (c + room <= FB_CAP) return;
    if (threadIdx.x == 0) fs.base = atomicAdd(counter, c);
    __syncthreads();
    const uint32_t base = fs.base;
    for (uint32_t i = threadIdx.x; i < c; i += blockDim.x) frontier[base + i] = fs.buf[i];
    __syncthreads();
    if (threadIdx.x == 0) fs.cnt = 0;
    __syncthreads();
}

// Displacement tracking.  dmax[x] = largest per-axis displacement body x had at any time in the current sweep (its
// swept AABB lies inside snapshot (+) dmax); reach[x] = the same over the earlier sweeps of this step (committed by
// k_restore).  moved_bits marks bodies with dmax or reach != 0 (k_verify_watch looks there first).
// A body whose AABB leaves its snapshot inflated by mu*extent is an ESCAPEE (esc_bits, escapees[]): the candidate
// set built from the snapshot may miss pairs it reaches; k_requery looks for them.
// Only one thread touches a dynamic body per round, so plain (L2) accesses suffice for dmax.
__device__ __forceinline__ void track_move(const BodyRW& x, uint32_t ix, float4 lo_snap, float mu_verify,
                                           float* __restrict__ dmax, uint32_t* __restrict__ moved_bits,
                                           uint32_t* __restrict__ esc_bits, uint32_t* __restrict__ escapees,
                                           uint32_t esc_cap, Ctl* ctl, float& maxrel, float& maxabs) {
    const V3 lo_now = x.omin + x.pos;
    const float d = fmaxf(fabsf(lo_now.x - lo_snap.x), fmaxf(fabsf(lo_now.y - lo_snap.y), fabsf(lo_now.z - lo_snap.z)));
    if (d == 0.0f) return;
    const float e = fmaxf(x.omax.x - x.omin.x, fmaxf(x.omax.y - x.omin.y, x.omax.z - x.omin.z));
    maxrel = fmaxf(maxrel, d / e);
    const float old = __ldcg(&dmax[ix]);
    const uint32_t bit = 1u << (ix & 31u);
    if (!(d <= old)) __stcg(&dmax[ix], d);
    if (old == 0.0f) atomicOr(&moved_bits[ix >> 5], bit);
    if (!(d <= mu_verify * e)) {
        const uint32_t prev = atomicOr(&esc_bits[ix >> 5], bit);
        if (!(prev & bit)) {  // first escape of this body in this step
            const uint32_t s = atomicAdd(&ctl->n_escapees, 1u);
            if (s < esc_cap)
                escapees[s] = ix;
            else
                atomicOr(&ctl->step_flags, SF_ESCAPEE_OVERFLOW);
        }
        maxabs = fmaxf(maxabs, d);
    }
}

constexpr int COOP_THREADS = 768;  // cooperative kernels: ONE large CTA per SM (a grid barrier costs per CTA)
__global__ void __launch_bounds__(COOP_THREADS, 1)
    k_sweep(const Bodies B, const float4* __restrict__ box, uint4* __restrict__ pairs, uint32_t pair_cap,
            uint32_t* __restrict__ frontier0, uint32_t* __restrict__ frontier1, Ctl* ctl, float* __restrict__ dmax,
            uint32_t* __restrict__ moved_bits, uint32_t* __restrict__ esc_bits, uint32_t* __restrict__ escapees,
            uint32_t esc_cap, const uint32_t* __restrict__ gate, const uint32_t* __restrict__ list, uint32_t skip) {
    // skip (EXPERIMENTAL, GP_SWEEP_SKIP=1, first sweep only, not yet run on hardware): a pair without the S_snapshot
    // flag whose two bodies have not been moved by an earlier pair of this sweep is tested on exactly the positions K4
    // tested (same operands, same comparisons) -- it misses again, and its bodies need not be fetched at all.  84 % of
    // the candidates of the uniform scene are such pairs.  "Moved" = the bit in moved_bits, which in this mode is set
    // for EVERY position write (the default sets it only when the box corner changed, which can differ by rounding).
    if (gate && *gate == 0) return;  // uniform across the grid: nobody reaches a barrier
    cg::grid_group grid = cg::this_grid();
    __shared__ FrontierStage fs;
    if (threadIdx.x == 0) fs.cnt = 0;
    __syncthreads();
    const uint32_t np = min(ctl->n_pairs, pair_cap);
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    // round 0 frontier: pairs at the head of all their chains.  First sweep: every pair is a candidate for it;
    // repair sweep (list != nullptr): only the pairs of the dirty components are swept again.
    const uint32_t nl = list ? ctl->n_dpairs : np;
    for (uint32_t p0 = blockIdx.x * blockDim.x; p0 < nl; p0 += gsz) {
        const uint32_t li = p0 + threadIdx.x;
        if (li < nl) {
            const uint32_t pid = list ? list[li] : li;
            if (__ldcg(&pair_link(pairs, pid)[2]) == 0) stage_push(fs, pid);
        }
        stage_flush(fs, frontier0, &ctl->frontier_n[0], blockDim.x);
    }
    stage_flush(fs, frontier0, &ctl->frontier_n[0], FB_CAP);
    grid.sync();
    const float mverify = 0.98f * ctl->margin;
    uint32_t rounds = 0, contacts = 0;
    float maxd = 0.0f, maxabs = 0.0f;
    while (true) {
        const uint32_t ci = rounds % 3u, co = (rounds + 1u) % 3u, cz = (rounds + 2u) % 3u;
        const uint32_t nf = __ldcg(&ctl->frontier_n[ci]);
        if (nf == 0) break;
        // counter cz was consumed in the previous round (all reads are behind that round's barrier)
        // and is next filled in the round after this one (behind this round's barrier)
        if (gtid == 0) ctl->frontier_n[cz] = 0;
        const uint32_t* fin = (rounds & 1u) ? frontier1 : frontier0;
        uint32_t* fout = (rounds & 1u) ? frontier0 : frontier1;
        for (uint32_t f0 = blockIdx.x * blockDim.x; f0 < nf; f0 += gsz) {
            const uint32_t f = f0 + threadIdx.x;
            if (f < nf) {
                const uint32_t pid = __ldcg(&fin[f]);
                const uint4 rec = __ldcg(&pairs[2 * (size_t)pid]);
                const uint4 link = __ldcg(&pairs[2 * (size_t)pid + 1]);  // same 32 B sector
                const uint32_t ia = rec.x & ~TOPBIT, ib = rec.y & ~TOPBIT;
                bool quick_miss = false;
                if (skip && !(rec.z & TOPBIT)) {
                    const uint32_t wa = __ldcg(&moved_bits[ia >> 5]), wb = __ldcg(&moved_bits[ib >> 5]);
                    quick_miss = !(((wa >> (ia & 31u)) | (wb >> (ib & 31u))) & 1u);
                }
                BodyRW a, b;
                a.is_static = (rec.x & TOPBIT) != 0;
                b.is_static = (rec.y & TOPBIT) != 0;
                if (!quick_miss) {
                const float4 pa = ldcg4(&B.P(ia)), pb = ldcg4(&B.P(ib));
                const float4 loa = ldcg4(&B.Lo(ia)), hia = ldcg4(&B.Hi(ia)), lob = ldcg4(&B.Lo(ib)), hib = ldcg4(&B.Hi(ib));
                a.pos = xyz(pa), a.mass = pa.w, a.omin = xyz(loa), a.omax = xyz(hia);
                b.pos = xyz(pb), b.mass = pb.w, b.omin = xyz(lob), b.omax = xyz(hib);
                const V3 amin = a.omin + a.pos, amax = a.omax + a.pos;
                const V3 bmin = b.omin + b.pos, bmax = b.omax + b.pos;
                // physics.rs:159-164 on the CURRENT positions
                const bool hit = amin.x < bmax.x && amax.x > bmin.x && amin.y < bmax.y && amax.y > bmin.y &&
                                 amin.z < bmax.z && amax.z > bmin.z;
                if (hit) {
                    ++contacts;
                    const float4 va = ldcg4(&B.V(ia)), vb = ldcg4(&B.V(ib));
                    a.vel = xyz(va), a.rest = va.w;
                    b.vel = xyz(vb), b.rest = vb.w;
                    const bool moved = resolve_pair(a, b, amin, amax, bmin, bmax);
                    if (!a.is_static) {
                        __stcg(&B.P(ia), make_float4(a.pos.x, a.pos.y, a.pos.z, a.mass));
                        __stcg(&B.V(ia), make_float4(a.vel.x, a.vel.y, a.vel.z, a.rest));
                    }
                    if (!b.is_static) {
                        __stcg(&B.P(ib), make_float4(b.pos.x, b.pos.y, b.pos.z, b.mass));
                        __stcg(&B.V(ib), make_float4(b.vel.x, b.vel.y, b.vel.z, b.rest));
                    }
                    if (moved) {
                        if (!a.is_static)
                            track_move(a, ia, box[2 * (size_t)ia], mverify, dmax, moved_bits, esc_bits, escapees, esc_cap,
                                       ctl, maxd, maxabs);
                        if (!b.is_static)
                            track_move(b, ib, box[2 * (size_t)ib], mverify, dmax, moved_bits, esc_bits, escapees, esc_cap,
                                       ctl, maxd, maxabs);
                        if (skip) {  // every position write counts (see above)
                            if (!a.is_static) atomicOr(&moved_bits[ia >> 5], 1u << (ia & 31u));
                            if (!b.is_static) atomicOr(&moved_bits[ib >> 5], 1u << (ib & 31u));
                        }
                    }
                    pairs[2 * (size_t)pid].w = rec.w | TOPBIT;  // S_sweep flag; k_restore: these bodies were written
                }
                }
                // wake the successors on both chains once they are at the head everywhere (static endpoints have
                // no chain: their link stays NONE)
                if (link.x != NONE && atomicSub(&pair_link(pairs, link.x)[2], 1u) == 1u) stage_push(fs, link.x);
                if (link.y != NONE && atomicSub(&pair_link(pairs, link.y)[2], 1u) == 1u) stage_push(fs, link.y);
            }
            stage_flush(fs, fout, &ctl->frontier_n[co], 2u * blockDim.x);
        }
        stage_flush(fs, fout, &ctl->frontier_n[co], FB_CAP);
        ++rounds;
        grid.sync();
    }
    // per-thread partials -> warp -> global
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        contacts += __shfl_xor_sync(0xffffffffu, contacts, o);
        maxd = fmaxf(maxd, __shfl_xor_sync(0xffffffffu, maxd, o));
        maxabs = fmaxf(maxabs, __shfl_xor_sync(0xffffffffu, maxabs, o));
    }
    if ((threadIdx.x & 31) == 0) {
        if (contacts) atomicAdd(&ctl->n_contacts, contacts);
        if (maxd > 0.0f) atomicMax(&ctl->max_disp_bits, __float_as_uint(maxd));
        if (maxabs > 0.0f) atomicMax(&ctl->max_abs_disp_bits, __float_as_uint(maxabs));
    }
    if (gtid == 0) {
        ctl->n_rounds = max(ctl->n_rounds, rounds);
        ctl->n_pairs_swept = np;
        if (!list) ctl->n_pairs_k4 = np;
    }
}


}  // namespace gp

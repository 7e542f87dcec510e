// gp_sweep_kernels.cuh — K6 ordered contact sweep (check_and_resolve_collision).
// Part of the physics step; see gp_kernels.cuh for the map of kernels to reference lines.
#pragma once
#include "gp_common.cuh"
#include "gp_chain_kernels.cuh"
#include "gp_pairfind_kernels.cuh"

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
// next-frontier counter is a single address, and same-address atomics serialise in L2.  A push never waits: when the
// stage is full it goes straight to the global list.  Escapees (bodies that left their margin) are staged the same way.
constexpr uint32_t FB_CAP = 12288;
constexpr uint32_t EB_CAP = 1024;
struct FrontierStage {
    uint32_t buf[FB_CAP];
    uint32_t ebuf[EB_CAP];
    uint32_t cnt, base, ecnt, ebase;
};
__device__ __forceinline__ void stage_push(FrontierStage& fs, uint32_t q, uint32_t* frontier, uint32_t* counter) {
    const uint32_t s = atomicAdd(&fs.cnt, 1u);
    if (s < FB_CAP)
        fs.buf[s] = q;
    else
        __stcg(&frontier[atomicAdd(counter, 1u)], q);
}
__device__ __forceinline__ void escapee_store(uint32_t* escapees, uint32_t esc_cap, Ctl* ctl, uint32_t slot, uint32_t ix) {
    if (slot < esc_cap)
        __stcg(&escapees[slot], ix);
    else
        atomicOr(&ctl->step_flags, SF_ESCAPEE_OVERFLOW);
}
// called by ALL threads of the CTA (contains barriers)
__device__ __forceinline__ void stage_flush(FrontierStage& fs, uint32_t* frontier, uint32_t* counter, uint32_t* escapees,
                                            uint32_t esc_cap, Ctl* ctl) {
    __syncthreads();
    const uint32_t c = min(fs.cnt, FB_CAP), ec = min(fs.ecnt, EB_CAP);
    if (threadIdx.x == 0 && c) fs.base = atomicAdd(counter, c);
    if (threadIdx.x == 32 && ec) fs.ebase = atomicAdd(&ctl->n_escapees, ec);
    __syncthreads();
    const uint32_t base = fs.base, ebase = fs.ebase;
    for (uint32_t i = threadIdx.x; i < c; i += blockDim.x) __stcg(&frontier[base + i], fs.buf[i]);
    for (uint32_t i = threadIdx.x; i < ec; i += blockDim.x) escapee_store(escapees, esc_cap, ctl, ebase + i, fs.ebuf[i]);
    __syncthreads();
    if (threadIdx.x == 0) fs.cnt = 0, fs.ecnt = 0;
    __syncthreads();
}

// Displacement tracking.  dmax[x] = largest per-axis displacement body x had at any time in the current sweep (its
// swept AABB lies inside snapshot (+) dmax); reach[x] = the same over the earlier sweeps of this step (committed by
// the repair).  moved_bits marks bodies with dmax or reach != 0 (the watch-list verification looks there first).
// A body whose AABB leaves its snapshot inflated by its margin is an ESCAPEE (esc_bits, escapees[]): the candidate
// set built from the snapshot may miss pairs it reaches; the requery phase looks for them.
// Only one thread touches a dynamic body per round, so plain (L2) accesses suffice for dmax.  `old` = dmax[ix] as the
// caller loaded it together with the body (one memory round trip for everything a contact needs).
constexpr float ESCAPE_FRAC = 0.98f;
__device__ __forceinline__ void track_move(const BodyRW& x, uint32_t ix, float4 lo_snap, float old, float* dmax,
                                           uint32_t* moved_bits, uint32_t* esc_bits, uint32_t* escapees,
                                           uint32_t esc_cap, Ctl* ctl, FrontierStage* fs, float& maxrel, float& maxabs) {
    const V3 lo_now = x.omin + x.pos;
    const float d = fmaxf(fabsf(lo_now.x - lo_snap.x), fmaxf(fabsf(lo_now.y - lo_snap.y), fabsf(lo_now.z - lo_snap.z)));
    if (d == 0.0f) return;
    const float e = fmaxf(x.omax.x - x.omin.x, fmaxf(x.omax.y - x.omin.y, x.omax.z - x.omin.z));
    maxrel = fmaxf(maxrel, d / e);
    const uint32_t bit = 1u << (ix & 31u);
    if (!(d <= old)) __stcg(&dmax[ix], d);
    if (old == 0.0f) atomicOr(&moved_bits[ix >> 5], bit);
    // escapee: moved further than 0.98 of the OUTER margin K3 stored for this body -- the very bits K4 compared
    // against (lo_snap.w), not a recomputation from the offsets, whose rounding differs at large coordinates
    if (!(d <= ESCAPE_FRAC * box_margin(lo_snap))) {
        const uint32_t prev = atomicOr(&esc_bits[ix >> 5], bit);
        if (!(prev & bit)) {  // first escape of this body in this step
            const uint32_t s = fs ? atomicAdd(&fs->ecnt, 1u) : EB_CAP;
            if (s < EB_CAP)
                fs->ebuf[s] = ix;
            else
                escapee_store(escapees, esc_cap, ctl, atomicAdd(&ctl->n_escapees, 1u), ix);
        }
        maxabs = fmaxf(maxabs, d);
    }
}

#ifndef GP_COOP_THREADS
#define GP_COOP_THREADS 768
#endif
constexpr int COOP_THREADS = GP_COOP_THREADS;  // the cooperative kernel: ONE large CTA per SM (a grid barrier costs per CTA)

}  // namespace gp

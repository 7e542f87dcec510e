// gp_slab_kernels.cuh — multi-GPU slabs: per-step exactness proof, owned-body download, set-up.
// Part of the physics step; see gp_kernels.cuh for the map of kernels to reference lines.
#pragma once
#include "gp_common.cuh"
#include "gp_pairfind_kernels.cuh"
#include "gp_repair_kernels.cuh"

namespace gp {

// ---------------------------------------------------------------------------------------------
// Slab verification (multi-GPU).  A rank knows its own bodies and, per the halo rule, every neighbour body whose
// inflated AABB comes within H of the shared face.  A body whose own reach (margin, or swept displacement if
// larger) pokes beyond that coverage may have a partner this rank has never seen: it is TAINTED.  The result of a
// body can only depend on bodies of its connected component of the candidate graph (sweep pairs + watch pairs), so
// taint spreads over components; every owned body that stays clean is bit-identical to the single-GPU result
// (multi-GPU exactness).  n_unverified counts the owned bodies that do not stay clean.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void taint_edge(uint32_t x, uint32_t y, uint32_t* __restrict__ taint_bits, uint32_t* flag) {
    if ((x | y) & TOPBIT) return;  // an edge through a static body carries nothing
    const bool tx = (__ldcg(&taint_bits[x >> 5]) >> (x & 31u)) & 1u, ty = (__ldcg(&taint_bits[y >> 5]) >> (y & 31u)) & 1u;
    if (tx == ty) return;
    const uint32_t z = tx ? y : x;
    atomicOr(&taint_bits[z >> 5], 1u << (z & 31u));
    *flag = 1u;
}

// A phase of k_contacts: clear, seed, propagate to the fixed point (a grid barrier per round), count.
constexpr uint32_t TAINT_MAX_ROUNDS = 64;
__device__ void taint_phase(const ContactArgs& a, cg::grid_group& grid) {
    const float4* sbox = a.sbox;
    const float *dmax = a.dmax, *reach = a.reach;
    const uint32_t* moved_bits = a.moved_bits;
    uint32_t* taint_bits = a.taint_bits;
    const uint4* pairs = a.pairs;
    const uint32_t pair_cap = a.pair_cap;
    const uint2* watch = a.watch;
    const float4* A = a.A1;
    Ctl* ctl = a.ctl;
    const uint32_t n = ctl->n;
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    const uint32_t words = (n + 31u) / 32u;
    (void)moved_bits;
    // seeds, part 1: bodies whose MARGIN pokes beyond what the neighbours mirror (K3 marked them while it had their
    // boxes in registers; the marks are consumed here)
    for (uint32_t wd = gtid; wd < words; wd += gsz) {
        const uint32_t t = a.seed_bits[wd];
        taint_bits[wd] = t;
        if (t) a.seed_bits[wd] = 0u;
    }
    if (gtid == 0) ctl->taint_flag[0] = 0u, ctl->taint_flag[1] = 0u, ctl->taint_flag[2] = 0u;
    grid.sync();
    // seeds, part 2: bodies that moved further than their margin (all of them are on the escapee list)
    {
        const uint32_t ne = min(__ldcg(&ctl->n_escapees), a.esc_cap);
        for (uint32_t e = gtid; e < ne; e += gsz) {
            const uint32_t i = __ldcg(&a.escapees[e]);
            const float4 lo = sbox[2 * (size_t)i], hi = sbox[2 * (size_t)i + 1];
            const float r = fmaxf(box_margin(lo), fmaxf(__ldcg(&reach[i]), __ldcg(&dmax[i]) * REACH_SLACK));
            const bool out_r = ctl->has_right && !(hi.x + r <= ctl->x_hi + ctl->halo);
            const bool out_l = ctl->has_left && !(lo.x - r >= ctl->x_lo - ctl->halo);
            if (out_l || out_r) atomicOr(&taint_bits[i >> 5], 1u << (i & 31u));
        }
    }
    grid.sync();
    const uint32_t np = min(__ldcg(&ctl->n_pairs), pair_cap), nw = min(ctl->n_watch, pair_cap);
    uint32_t it = 0;
    bool open = true;
    while (open && it < TAINT_MAX_ROUNDS) {
        // three rotating flags: round `it` raises flag it%3 and clears flag (it+1)%3, which was last READ at the end
        // of round it-2, i.e. behind the barrier of round it-1 (with two flags a fast thread would clear the flag a
        // slow thread is still about to read, and the grid would disagree about leaving the loop)
        uint32_t* flag = &ctl->taint_flag[it % 3u];
        if (gtid == 0) ctl->taint_flag[(it + 1u) % 3u] = 0u;
        for (uint32_t i = gtid; i < np; i += gsz) {
            const uint4 r = __ldcg(&pairs[2 * (size_t)i]);
            taint_edge(r.x, r.y, taint_bits, flag);
        }
        for (uint32_t i = gtid; i < nw; i += gsz) {
            const uint2 r = __ldcg(&watch[i]);
            if (r.x != NONE) taint_edge(r.x, r.y, taint_bits, flag);  // (activated entries are among the pairs)
        }
        grid.sync();
        open = __ldcg(flag) != 0u;
        ++it;
    }
    // owned dynamic bodies that are not provably independent of unseen bodies (the tainted ones are few: walk the
    // set bits, not the bodies)
    uint32_t c = 0;
    if (open) {  // the propagation did not reach its fixed point within the round limit: nothing is proven
        for (uint32_t i = gtid; i < n; i += gsz) {
            const uint32_t f = __float_as_uint(A[i].w);
            if (!(f & (F_GHOST | F_STATIC))) ++c;
        }
    } else {
        for (uint32_t wd = gtid; wd < words; wd += gsz) {
            uint32_t t = __ldcg(&taint_bits[wd]);
            while (t) {
                const uint32_t i = wd * 32u + (uint32_t)__ffs(t) - 1u;
                t &= t - 1u;
                if (i < n && !(__float_as_uint(A[i].w) & (F_GHOST | F_STATIC))) ++c;
            }
        }
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&ctl->n_unverified, c);
}

// Owned bodies (no halo copies; replicated big statics only on the rank that has no left neighbour) -> host layout,
// in SLOT ORDER (flags, exclusive scan, gather): the order is deterministic, so the host can hand the same rows back
// with gp_set_state_owned (row k goes to owned_slot[k]) as long as no step ran in between.
__global__ void __launch_bounds__(256)
    k_owned_flags(uint32_t cap, const Ctl* ctl, const float4* __restrict__ A, int is_first_rank, uint32_t* __restrict__ flags) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap) return;
    uint32_t own = 0;
    if (i < ctl->n_in) {  // between steps n_in = the live bodies of the last step
        const uint32_t f = __float_as_uint(A[i].w);
        own = !(f & F_GHOST) && (!(f & F_BIG) || is_first_rank);
    }
    flags[i] = own;
}
__global__ void __launch_bounds__(256)
    k_owned_gather(uint32_t cap, const Ctl* ctl, const Bodies B, const uint32_t* __restrict__ flags,
                   const uint32_t* __restrict__ offs, uint32_t* __restrict__ ids, float* __restrict__ pos3,
                   float* __restrict__ vel3, const uint32_t* __restrict__ taint_bits, uint8_t* __restrict__ unverified,
                   uint32_t* __restrict__ owned_slot) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap || !flags[i]) return;
    const uint32_t s = offs[i];
    const float4 p = B.P(i), v = B.V(i);
    ids[s] = __float_as_uint(B.Lo(i).w);
    pos3[3 * (size_t)s] = p.x, pos3[3 * (size_t)s + 1] = p.y, pos3[3 * (size_t)s + 2] = p.z;
    vel3[3 * (size_t)s] = v.x, vel3[3 * (size_t)s + 1] = v.y, vel3[3 * (size_t)s + 2] = v.z;
    if (unverified) unverified[s] = taint_bits ? (taint_bits[i >> 5] >> (i & 31u)) & 1u : 0u;  // of the LAST step
    owned_slot[s] = i;
}
// rows of the last gp_download_owned back into the bodies they came from
__global__ void __launch_bounds__(256)
    k_set_state_slots(uint32_t k, const uint32_t* __restrict__ owned_slot, const float* __restrict__ pos3,
                      const float* __restrict__ vel3, const float* __restrict__ acc3, Bodies B, float4* __restrict__ A) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= k) return;
    const uint32_t i = owned_slot[t];
    const size_t r = 3 * (size_t)t;
    if (pos3) {
        float4 p = B.P(i);
        p.x = pos3[r], p.y = pos3[r + 1], p.z = pos3[r + 2];
        B.P(i) = p;
    }
    if (vel3) {
        float4 v = B.V(i);
        v.x = vel3[r], v.y = vel3[r + 1], v.z = vel3[r + 2];
        B.V(i) = v;
    }
    if (acc3) {
        float4 a = A[i];
        a.x = acc3[r], a.y = acc3[r + 1], a.z = acc3[r + 2];
        A[i] = a;
    }
}

// gp_comm_init: body ids become the caller's global entity ids; count what the faces would send right now
__global__ void k_slab_prepare(uint32_t n, const Bodies B, const float4* __restrict__ A,
                               const uint32_t* __restrict__ entity, uint32_t* __restrict__ counts /*[3]: L, R, dyn big*/,
                               const Ctl* ctl) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (entity) B.Lo(i).w = __uint_as_float(entity[i]);  // at upload slot == caller index
    const uint32_t f = __float_as_uint(A[i].w);
    if (f & F_BIG) {
        if (!(f & F_STATIC)) atomicAdd(&counts[2], 1u);
        return;
    }
    const float4 p = B.P(i), lo = B.Lo(i), hi = B.Hi(i);
    const float wlo = lo.x + p.x, whi = hi.x + p.x;
    const float m = whi - wlo;  // generous: a margin of one extent
    if (ctl->has_left && wlo - m < ctl->x_lo + 2.0f * ctl->halo) atomicAdd(&counts[0], 1u);
    if (ctl->has_right && whi + m > ctl->x_hi - 2.0f * ctl->halo) atomicAdd(&counts[1], 1u);
}


}  // namespace gp

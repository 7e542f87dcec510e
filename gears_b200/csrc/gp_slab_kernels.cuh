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
// Lock-free union-find over body slots (link the larger root to the smaller one, path halving), as in connected-
// component labelling: ONE pass over the candidate edges replaces the round-by-round propagation (which cost a full
// scan of all pairs + a grid barrier per hop of the longest tainted chain: 79 us per step at 2.3 M bodies per rank).
__device__ __forceinline__ uint32_t uf_find(uint32_t* parent, uint32_t x) {
    uint32_t p = __ldcg(&parent[x]);
    while (p != x) {
        const uint32_t gp = __ldcg(&parent[p]);
        if (gp != p) parent[x] = gp;  // path halving (any ancestor is a valid parent: a benign race)
        x = p;
        p = gp;
    }
    return x;
}
__device__ __forceinline__ void uf_unite(uint32_t* parent, uint32_t x, uint32_t y) {
    for (;;) {
        x = uf_find(parent, x);
        y = uf_find(parent, y);
        if (x == y) return;
        if (x < y) {
            const uint32_t t = x;
            x = y;
            y = t;
        }
        if (atomicCAS(&parent[x], x, y) == x) return;  // x was still a root: now it hangs under the smaller root y
    }
}

// The components of the candidate graph: sweep pairs [p0, p1) and (optionally) the watch list, spread over `nthr`
// threads.  The edges do not depend on the sweep, so the CTAs that idle while CTA 0 runs the tail levels of the first
// sweep do this part of the proof there (pairs appended later by the verification are added by taint_phase).
// One edge at a time per thread: a variant with four edges in flight and a compare-and-swap fast path for root-root
// edges was measured SLOWER, 0.31 against 0.20 ms at 8.4 M bodies per rank.
__device__ void proof_unions(const ContactArgs& a, uint32_t p0, uint32_t p1, bool with_watch, uint32_t tid,
                             uint32_t nthr) {
    uint32_t* parent = a.uf_parent;
    for (uint32_t i = p0 + tid; i < p1; i += nthr) {
        const uint2 r = __ldcg(reinterpret_cast<const uint2*>(&a.pairs[2 * (size_t)i]));  // (x, y): never rewritten
        if (!((r.x | r.y) & TOPBIT)) uf_unite(parent, r.x, r.y);
    }
    if (!with_watch) return;
    const uint32_t nw = min(a.ctl->n_watch, a.pair_cap);
    for (uint32_t i = tid; i < nw; i += nthr) {
        const uint2 r = __ldcg(&a.watch[i]);
        // (an entry the verification activates later is then among the pairs as well: the same edge twice)
        if (r.x != NONE && !((r.x | r.y) & TOPBIT)) uf_unite(parent, r.x, r.y);
    }
}

// A phase of k_contacts.  A body's result can only depend on its connected component of the candidate graph
// (sweep pairs + watch pairs; an edge through a static body carries nothing), so: label the components, mark the
// components that contain a seed, and every owned dynamic body of a marked component is unverified.
// uf_parent = the forest (K3 left every body its own component); scratch, free once the repair loop is over:
// ra.dirty_bits = marked roots (left clean again).  pairs_done: the pairs [0, pairs_done) and the watch list were
// united during the first sweep's tail (0: nothing was).
__device__ void taint_phase(const ContactArgs& a, cg::grid_group& grid, uint32_t pairs_done) {
    const float4* sbox = a.sbox;
    const float *dmax = a.dmax, *reach = a.reach;
    uint32_t* taint_bits = a.taint_bits;
    uint32_t* seed_bits = a.seed_bits;
    uint32_t* parent = a.uf_parent;  // K3 left every body its own component
    uint32_t* rootmark = a.ra.dirty_bits;
    const float4* A = a.A1;
    Ctl* ctl = a.ctl;
    const uint32_t n = ctl->n;
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    const uint32_t words = (n + 31u) / 32u;
    const uint32_t np = min(__ldcg(&ctl->n_pairs), a.pair_cap);
    // ---- seeds, part 2: bodies that moved further than their margin (all of them are on the escapee list) join the
    //      seeds K3 marked from the margins (part 1); nothing reads the seeds before the barrier after the unions ----
    {
        const uint32_t ne = min(__ldcg(&ctl->n_escapees), a.esc_cap);
        for (uint32_t e = gtid; e < ne; e += gsz) {
            const uint32_t i = __ldcg(&a.escapees[e]);
            const float4 lo = sbox[2 * (size_t)i], hi = sbox[2 * (size_t)i + 1];
            const float r = fmaxf(box_margin(lo), fmaxf(__ldcg(&reach[i]), __ldcg(&dmax[i]) * REACH_SLACK));
            const bool out_r = ctl->has_right && !(hi.x + r <= ctl->x_hi + ctl->halo);
            const bool out_l = ctl->has_left && !(lo.x - r >= ctl->x_lo - ctl->halo);
            if (out_l || out_r) atomicOr(&seed_bits[i >> 5], 1u << (i & 31u));
        }
    }
    // ---- components: the edges the tail of the first sweep has not seen ----
    proof_unions(a, min(pairs_done, np), np, pairs_done == 0u, gtid, gsz);
    grid.sync();
    // ---- mark the components of the seeds ----
    for (uint32_t wd = gtid; wd < words; wd += gsz) {
        uint32_t t = __ldcg(&seed_bits[wd]);
        while (t) {
            const uint32_t i = wd * 32u + (uint32_t)__ffs(t) - 1u;
            t &= t - 1u;
            if (i < n) {
                const uint32_t r = uf_find(parent, i);
                atomicOr(&rootmark[r >> 5], 1u << (r & 31u));
            }
        }
    }
    grid.sync();
    // ---- a body is tainted iff its component is marked; count the owned dynamic ones ----
    uint32_t c = 0;
    // (a warp writes TW words of taint bits per iteration: the TW parent look-ups of a lane are in flight together,
    // then the TW mark look-ups -- one body at a time per lane left the pass latency-bound)
    constexpr int TW = 8;
    for (uint32_t i0 = (gtid & ~31u) * TW; i0 < words * 32u; i0 += gsz * TW) {
        uint32_t par[TW], root[TW], mark[TW];
#pragma unroll
        for (int u = 0; u < TW; ++u) {
            const uint32_t i = i0 + 32u * u + (threadIdx.x & 31u);
            par[u] = i < n ? __ldcg(&parent[i]) : NONE;
        }
#pragma unroll
        for (int u = 0; u < TW; ++u) {
            const uint32_t i = i0 + 32u * u + (threadIdx.x & 31u);
            root[u] = par[u] == NONE ? NONE : (par[u] == i ? i : uf_find(parent, par[u]));
        }
#pragma unroll
        for (int u = 0; u < TW; ++u) mark[u] = root[u] == NONE ? 0u : __ldcg(&rootmark[root[u] >> 5]);
#pragma unroll
        for (int u = 0; u < TW; ++u) {
            const uint32_t i = i0 + 32u * u + (threadIdx.x & 31u);
            const bool t = root[u] != NONE && ((mark[u] >> (root[u] & 31u)) & 1u);
            if (t && !(__float_as_uint(A[i].w) & (F_GHOST | F_STATIC))) ++c;
            const uint32_t word = __ballot_sync(0xffffffffu, t);
            if ((threadIdx.x & 31u) == 0u && i0 + 32u * u < words * 32u) taint_bits[(i0 >> 5) + u] = word;
        }
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&ctl->n_unverified, c);
}

// Behind the grid barrier that follows taint_phase (nothing reads the marks any more): leave the scratch clean --
// the marked roots and the seeds.  The last thing the kernel does, so no barrier follows.
__device__ void taint_cleanup(const ContactArgs& a) {
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    const uint32_t words = (a.ctl->n + 31u) / 32u;
    for (uint32_t wd = gtid; wd < words; wd += gsz) {
        if (__ldcg(&a.seed_bits[wd])) a.seed_bits[wd] = 0u;
        if (__ldcg(&a.ra.dirty_bits[wd])) a.ra.dirty_bits[wd] = 0u;
    }
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

// ---- re-balancing of the slab faces (gp_slab_rebalance): where are the bodies this rank owns? ----
constexpr uint32_t RB_BINS = 4096;
__device__ __forceinline__ bool rb_owned_centre(const Ctl* ctl, const Bodies B, const float4* A, uint32_t i, float& cx) {
    if (i >= ctl->n_in) return false;  // between steps n_in = the live bodies of the last step
    const uint32_t f = __float_as_uint(A[i].w);
    if (f & (F_GHOST | F_BIG)) return false;  // halo copies belong to the neighbour; big statics are replicated
    const float4 p = B.P(i), lo = B.Lo(i), hi = B.Hi(i);
    cx = ((lo.x + p.x) + (hi.x + p.x)) * 0.5f;  // the very expression of the ownership test in k_keys<true>
    return isfinite(cx);
}
// minmax[0] = min, [1] = max of the owned centres, as order-preserving uints (init: 0xffffffff, 0)
__global__ void __launch_bounds__(256)
    k_owned_minmax(uint32_t cap, const Ctl* ctl, const Bodies B, const float4* __restrict__ A, uint32_t* minmax) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float cx = 0.f;
    const bool own = i < cap && rb_owned_centre(ctl, B, A, i, cx);
    float lo = own ? cx : 3.0e38f, hi = own ? cx : -3.0e38f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        if (lo < 3.0e38f) atomicMin(&minmax[0], f2ord(lo));
        if (hi > -3.0e38f) atomicMax(&minmax[1], f2ord(hi));
    }
}
// histogram of the owned centres over RB_BINS equal bins of [x0, x0 + RB_BINS / inv_bin)
__global__ void __launch_bounds__(256)
    k_owned_xhist(uint32_t cap, const Ctl* ctl, const Bodies B, const float4* __restrict__ A, float x0, float inv_bin,
                  uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[RB_BINS];
    for (uint32_t b = threadIdx.x; b < RB_BINS; b += blockDim.x) sh[b] = 0;
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += gridDim.x * blockDim.x) {
        float cx;
        if (!rb_owned_centre(ctl, B, A, i, cx)) continue;
        const int b = min(max(__float2int_rd((cx - x0) * inv_bin), 0), (int)RB_BINS - 1);
        atomicAdd(&sh[b], 1u);
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < RB_BINS; b += blockDim.x)
        if (sh[b]) atomicAdd(&hist[b], sh[b]);
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

// gp_repair_kernels.cuh — verification of the speculative pair set, component-local repair, end of step.
// Part of the physics step; see gp_kernels.cuh for the map of kernels to reference lines.
#pragma once
#include "gp_common.cuh"
#include "gp_mirror_kernels.cuh"
#include "gp_pairfind_kernels.cuh"
#include "gp_chain_kernels.cuh"
#include "gp_sweep_kernels.cuh"

namespace gp {

// ---------------------------------------------------------------------------------------------
// Repair loop of the speculative broad phase.
//  k_requery : for every escapee x, find the bodies y with swept(x) ^ swept(y) != {} that are NOT
//              candidates yet, append those pairs, raise ctl->repair_needed.  One warp per escapee; the
//              grid rows its query box covers are contiguous runs of the cell-sorted AABB array.
//  k_repair_local : (gated) undo and re-chain only the connected components the new pairs touch (see below);
//              the repair sweep then covers those pairs only.
// When k_requery appends nothing the last sweep saw every pair that could intersect at its turn: the
// result equals the reference's full O(N^2) sweep.
// ---------------------------------------------------------------------------------------------
// Candidate predicate of the whole step: overlap(snap x, snap y, r_x + r_y), r_x = max(mu*extent(x), reach[x]).
// K4 built the set for reach = 0; a sweep in which x moved by dmax[x] > reach[x] extends it.  The pairs that
// pass with the grown reaches but not with the old ones are exactly the new ones (no duplicates).
constexpr float REACH_SLACK = 1.0001f;  // covers the rounding of the measured displacement

__global__ void __launch_bounds__(256)
    k_requery(const float4* __restrict__ box, const float4* __restrict__ sbox,
              const uint32_t* __restrict__ LB, const float* __restrict__ dmax, const float* __restrict__ reach,
              const uint32_t* __restrict__ esc_bits, const uint32_t* __restrict__ escapees, uint32_t esc_cap,
              uint4* __restrict__ pairs, uint32_t pair_cap, Ctl* ctl) {
    const uint32_t ne = min(ctl->n_escapees, esc_cap);
    if (ne == 0) return;
    const GridParams g = ctl->grid;
    const float mu = ctl->margin;
    const float ext_max = ctl->ext_max;
    const float partner_reach = fmaxf(mu * ext_max, __uint_as_float(ctl->max_abs_disp_bits) * REACH_SLACK);
    const uint32_t n_small = ctl->n_small, n = ctl->n;
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t ei = warp; ei < ne; ei += nwarps) {
        const uint32_t x = escapees[ei];
        const float rx_prev = reach[x];
        const float rx_now = dmax[x] * REACH_SLACK;
        if (!(rx_now > rx_prev)) continue;  // did not move further than in an earlier sweep
        const float4 alo = box[2 * (size_t)x], ahi = box[2 * (size_t)x + 1];
        const float ma = box_margin(alo);  // the very bits K4 used
        const float ra_old = fmaxf(ma, rx_prev), ra_new = fmaxf(ma, rx_now);
        const uint32_t ta = x | box_static(alo), ra = __float_as_uint(ahi.w);
        // partner centres lie within the query box grown by the partner's own reach and half extent
        const float grow = (ra_new + partner_reach) * 1.001f + 0.5f * ext_max * 1.001f;
        int c0[3], c1[3];
        {
            const float lo3[3] = {alo.x - grow, alo.y - grow, alo.z - grow};
            const float hi3[3] = {ahi.x + grow, ahi.y + grow, ahi.z + grow};
            const float o3[3] = {g.ox, g.oy, g.oz};
            const int d3[3] = {g.dx, g.dy, g.dz};
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                c0[k] = min(max(__float2int_rd((lo3[k] - o3[k]) * g.inv_cell) + 1, 1), d3[k] - 2);
                c1[k] = min(max(__float2int_rd((hi3[k] - o3[k]) * g.inv_cell) + 1, 1), d3[k] - 2);
            }
        }
        auto test = [&](uint32_t q) {
            const float4 blo = sbox[2 * (size_t)q], bhi = sbox[2 * (size_t)q + 1];
            const uint32_t y = q;
            const uint32_t tb = y | box_static(blo);
            if (y == x) return;
            if (ta & tb & TOPBIT) return;  // static-static (an escapee is dynamic; kept for symmetry)
            const float mb = box_margin(blo);
            const float ry_prev = reach[y], ry_now = dmax[y] * REACH_SLACK;
            const float rb_old = fmaxf(mb, ry_prev), rb_new = fmaxf(rb_old, ry_now);
            if (inflated_overlap(alo, ahi, blo, bhi, ra_old + rb_old)) return;   // found in an earlier round / by K4
            if (!inflated_overlap(alo, ahi, blo, bhi, ra_new + rb_new)) return;  // still out of reach
            // y is an escapee that grew too: its own query reports the pair (the lower slot wins)
            if (y < x && ry_now > ry_prev && ((esc_bits[y >> 5] >> (y & 31u)) & 1u)) return;
            const uint32_t rb = __float_as_uint(bhi.w);
            uint4 rec = (ra < rb) ? make_uint4(ta, tb, ra, rb) : make_uint4(tb, ta, rb, ra);
            cg::coalesced_group cgp = cg::coalesced_threads();
            uint32_t base = 0;
            if (cgp.thread_rank() == 0) base = atomicAdd(&ctl->n_pairs, cgp.size());
            base = cgp.shfl(base, 0);
            const uint32_t slot = base + cgp.thread_rank();
            if (slot < pair_cap) {
                pairs[2 * (size_t)slot] = rec;
                pairs[2 * (size_t)slot + 1] = link_init();
            } else {
                atomicOr(&ctl->step_flags, SF_PAIR_OVERFLOW);
            }
            ctl->repair_needed = 1u;
            atomicAdd(&ctl->n_requeried, 1u);
        };
        for (int z = c0[2]; z <= c1[2]; ++z)
            for (int y = c0[1]; y <= c1[1]; ++y) {
                const uint32_t rowk = (uint32_t)g.dx * ((uint32_t)y + (uint32_t)g.dy * (uint32_t)z);
                const uint32_t q0 = LB[rowk + (uint32_t)c0[0]], q1 = LB[rowk + (uint32_t)c1[0] + 1u];
                for (uint32_t q = q0 + lane; q < q1; q += 32) test(q);
            }
        for (uint32_t q = n_small + lane; q < n; q += 32) test(q);
    }
}

// ---------------------------------------------------------------------------------------------
// Component-local repair (ONE cooperative launch, gated on ctl->repair_needed).  The pairs that verification
// appended can only change the result of bodies in their connected components of the sweep-pair graph; every other
// component keeps the result of the sweep that already ran.  Phases (a grid barrier between them):
//   seed     dynamic endpoints of the new pairs are DIRTY
//   closure  breadth-first over the per-body adjacency of K5 (+ the few pairs appended since K4), level by level
//   restore  dirty bodies go back to their post-integration state; their pairs are collected (each once), lose their
//            hit flag and links; every escapee commits its reach
//   chains   degrees, segments (bump allocation in adj2), fill, sort by partner rank, successor links, pending
// The repair sweep then runs over the collected pairs only (k_sweep, list mode).  If more than a third of the live
// bodies turn dirty (a pile: one giant component) everything is declared dirty and the same code re-does it all.
// ---------------------------------------------------------------------------------------------
struct RepairArrays {
    uint32_t* dirty_bits;        // per slot
    uint32_t* dlist;             // dirty bodies
    uint32_t* dpairs;            // pairs of the dirty components
    uint32_t* offs2;             // per slot: start of the body's segment in adj2
    uint32_t* deg2;              // per slot: its length
    unsigned long long* adj2;    // chain entries of the dirty bodies
    uint32_t* dheavy;            // dirty bodies whose chain needs the block-wide sorter
    uint32_t dheavy_cap;
};

__device__ __forceinline__ bool mark_dirty(uint32_t x, const RepairArrays& ra, Ctl* ctl, uint32_t* level_counter) {
    const uint32_t bit = 1u << (x & 31u);
    if (atomicOr(&ra.dirty_bits[x >> 5], bit) & bit) return false;
    ra.dlist[atomicAdd(&ctl->n_dirty, 1u)] = x;
    atomicAdd(level_counter, 1u);
    return true;
}
__device__ __forceinline__ bool is_dirty(uint32_t x, const RepairArrays& ra) {
    return (__ldcg(&ra.dirty_bits[x >> 5]) >> (x & 31u)) & 1u;
}

__global__ void __launch_bounds__(COOP_THREADS, 1)
    k_repair_local(uint4* __restrict__ pairs, uint32_t pair_cap, const uint32_t* __restrict__ offs,
                   const unsigned long long* __restrict__ adj, uint32_t* __restrict__ chain_cnt,
                   const uint32_t* __restrict__ perm, const Bodies B0, const float4* __restrict__ A0, const Bodies B1,
                   float* __restrict__ dmax, float* __restrict__ reach, const uint32_t* __restrict__ escapees,
                   uint32_t esc_cap, RepairArrays ra, Ctl* ctl, float dt, float d16, float d18, float d35) {
    if (ctl->repair_needed == 0) return;  // uniform: nobody reaches a barrier
    cg::grid_group grid = cg::this_grid();
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    const uint32_t n = ctl->n;
    const uint32_t np0 = ctl->n_pairs_k4, np1 = ctl->n_pairs_swept, np2 = min(ctl->n_pairs, pair_cap);
    if (gtid == 0) {
        ctl->n_dirty = 0, ctl->n_dpairs = 0, ctl->dirty_bump = 0, ctl->n_dheavy = 0, ctl->dirty_all = 0;
        ctl->n_hits_cleared = 0;
        ctl->dirty_level[0] = ctl->dirty_level[1] = ctl->dirty_level[2] = 0;
    }
    grid.sync();
    // ---- seed (counted as level 0) ----
    for (uint32_t i = np1 + gtid; i < np2; i += gsz) {
        const uint4 r = pairs[2 * (size_t)i];
        if (!(r.x & TOPBIT)) mark_dirty(r.x, ra, ctl, &ctl->dirty_level[0]);
        if (!(r.y & TOPBIT)) mark_dirty(r.y & ~TOPBIT, ra, ctl, &ctl->dirty_level[0]);
    }
    grid.sync();
    // ---- closure, level by level: level k reads the bodies [lb, le) and appends level k+1 ----
    uint32_t lb = 0, le = __ldcg(&ctl->dirty_level[0]);
    bool all = false;
    for (uint32_t lvl = 0; le > lb; ++lvl) {
        // counter of level lvl+2 was last read at the end of level lvl-1 (behind that barrier)
        if (gtid == 0) ctl->dirty_level[(lvl + 2u) % 3u] = 0u;
        uint32_t* nxt = &ctl->dirty_level[(lvl + 1u) % 3u];
        for (uint32_t i = lb + gtid; i < le; i += gsz) {
            const uint32_t x = __ldcg(&ra.dlist[i]);
            for (uint32_t e = offs[x]; e < offs[x + 1]; ++e) {
                const uint32_t ent = (uint32_t)adj[e];
                const uint4 r = pairs[2 * (size_t)(ent & ~TOPBIT)];
                const uint32_t y = (ent & TOPBIT) ? r.x : r.y;  // the partner of x in this pair
                if (!(y & TOPBIT)) mark_dirty(y, ra, ctl, nxt);
            }
        }
        // pairs appended since K4 are not in the adjacency: scan them (a few hundred)
        for (uint32_t i = np0 + gtid; i < np2; i += gsz) {
            const uint4 r = pairs[2 * (size_t)i];
            if ((r.x | r.y) & TOPBIT) continue;
            const bool dx = is_dirty(r.x, ra), dy = is_dirty(r.y, ra);
            if (dx != dy) mark_dirty(dx ? r.y : r.x, ra, ctl, nxt);
        }
        grid.sync();
        lb = le;
        le += __ldcg(nxt);
        if (le > n / 3u + 1024u) {  // (uniform: every thread computes the same le)
            all = true;
            break;
        }
    }
    if (all) {  // one giant component: everything is dirty
        grid.sync();  // everybody has left the loop before the list is rebuilt
        if (gtid == 0) ctl->n_dirty = 0, ctl->dirty_all = 1;
        grid.sync();
        for (uint32_t x = gtid; x < n; x += gsz) {
            const uint32_t f = __float_as_uint(A0[perm[x]].w);
            if (f & F_STATIC) continue;
            atomicOr(&ra.dirty_bits[x >> 5], 1u << (x & 31u));
            ra.dlist[atomicAdd(&ctl->n_dirty, 1u)] = x;
        }
        grid.sync();
    }
    const uint32_t nd = __ldcg(&ctl->n_dirty);
    // ---- restore the dirty bodies, collect their pairs ----
    uint32_t cleared = 0;
    for (uint32_t i = gtid; i < nd; i += gsz) {
        const uint32_t x = __ldcg(&ra.dlist[i]);
        const uint32_t s0 = perm[x];  // its slot in the untouched input set
        float4 p = B0.P(s0), v = B0.V(s0);
        const float4 a = A0[s0];
        if (!(__float_as_uint(a.w) & F_INTEGRATED)) integrate_body(p, v, a, dt, d16, d18, d35);
        B1.P(x) = p;
        B1.V(x) = v;
        const float dm = dmax[x];
        if (dm != 0.0f) {  // the pairs within the grown reach are candidates now: commit it, start a fresh record
            reach[x] = fmaxf(reach[x], dm * REACH_SLACK);
            dmax[x] = 0.0f;
        }
        chain_cnt[x] = 0;
        for (uint32_t e = offs[x]; e < offs[x + 1]; ++e) {
            const uint32_t ent = (uint32_t)adj[e], pid = ent & ~TOPBIT;
            const uint4 r = pairs[2 * (size_t)pid];
            // a pair is collected by its first dynamic endpoint (all its dynamic endpoints are dirty)
            if ((ent & TOPBIT) && !(r.x & TOPBIT)) continue;
            ra.dpairs[atomicAdd(&ctl->n_dpairs, 1u)] = pid;
            if (r.w & TOPBIT) {
                pairs[2 * (size_t)pid].w = r.w & ~TOPBIT;
                ++cleared;
            }
            pairs[2 * (size_t)pid + 1] = link_init();
        }
    }
    for (uint32_t i = np0 + gtid; i < np2; i += gsz) {  // pairs appended since K4
        const uint4 r = pairs[2 * (size_t)i];
        const bool d = (!(r.x & TOPBIT) && is_dirty(r.x, ra)) || (!(r.y & TOPBIT) && is_dirty(r.y & ~TOPBIT, ra));
        if (!d) continue;
        ra.dpairs[atomicAdd(&ctl->n_dpairs, 1u)] = i;
        if (r.w & TOPBIT) {
            pairs[2 * (size_t)i].w = r.w & ~TOPBIT;
            ++cleared;
        }
        pairs[2 * (size_t)i + 1] = link_init();
    }
    {  // every escapee commits the reach of the sweep that ran (clean ones keep their sweep record)
        const uint32_t ne = min(ctl->n_escapees, esc_cap);
        for (uint32_t i = gtid; i < ne; i += gsz) {
            const uint32_t x = escapees[i];
            if (!is_dirty(x, ra)) reach[x] = fmaxf(reach[x], dmax[x] * REACH_SLACK);
        }
    }
    cleared = __reduce_add_sync(0xffffffffu, cleared);
    if ((threadIdx.x & 31) == 0 && cleared) atomicAdd(&ctl->n_hits_cleared, cleared);
    grid.sync();
    // ---- chains of the dirty bodies ----
    const uint32_t ndp = __ldcg(&ctl->n_dpairs);
    for (uint32_t j = gtid; j < ndp; j += gsz) {  // degrees
        const uint4 r = pairs[2 * (size_t)__ldcg(&ra.dpairs[j])];
        if (!(r.x & TOPBIT)) atomicAdd(&chain_cnt[r.x], 1u);
        if (!(r.y & TOPBIT)) atomicAdd(&chain_cnt[r.y], 1u);
    }
    grid.sync();
    for (uint32_t i = gtid; i < nd; i += gsz) {  // segments
        const uint32_t x = __ldcg(&ra.dlist[i]);
        const uint32_t d = __ldcg(&chain_cnt[x]);
        ra.deg2[x] = d;
        ra.offs2[x] = d ? atomicAdd(&ctl->dirty_bump, d) : 0u;
    }
    grid.sync();
    for (uint32_t j = gtid; j < ndp; j += gsz) {  // fill
        const uint32_t pid = __ldcg(&ra.dpairs[j]);
        const uint4 r = pairs[2 * (size_t)pid];
        const uint32_t rk_a = r.z & ~TOPBIT, rk_b = r.w & ~TOPBIT;
        if (!(r.x & TOPBIT)) {
            const uint32_t a = r.x;
            ra.adj2[__ldcg(&ra.offs2[a]) + atomicSub(&chain_cnt[a], 1u) - 1u] = ((unsigned long long)rk_b << 32) | pid;
        }
        if (!(r.y & TOPBIT)) {
            const uint32_t b = r.y;
            ra.adj2[__ldcg(&ra.offs2[b]) + atomicSub(&chain_cnt[b], 1u) - 1u] =
                ((unsigned long long)rk_a << 32) | TOPBIT | pid;
        }
    }
    grid.sync();
    for (uint32_t i = gtid; i < nd; i += gsz) {  // sort + link (short chains)
        const uint32_t x = __ldcg(&ra.dlist[i]);
        const uint32_t d = __ldcg(&ra.deg2[x]);
        if (d == 0) continue;
        if (d > CHAIN_INLINE) {
            const uint32_t hs = atomicAdd(&ctl->n_dheavy, 1u);
            if (hs < ra.dheavy_cap)
                ra.dheavy[hs] = x;
            else
                atomicOr(&ctl->step_flags, SF_HEAVY_OVERFLOW);
            continue;
        }
        unsigned long long* seg = ra.adj2 + __ldcg(&ra.offs2[x]);
        // the segment was filled by other SMs in the previous phase: fetch it from L2 once, sort a private copy
        unsigned long long buf[CHAIN_INLINE];
        for (uint32_t k = 0; k < d; ++k) buf[k] = __ldcg(&seg[k]);
        for (uint32_t k = 1; k < d; ++k) {
            const unsigned long long v = buf[k];
            uint32_t j = k;
            while (j > 0 && buf[j - 1] > v) {
                buf[j] = buf[j - 1];
                --j;
            }
            buf[j] = v;
        }
        for (uint32_t k = 1; k < d; ++k) {
            const uint32_t prev = (uint32_t)buf[k - 1], cur = (uint32_t)buf[k] & ~TOPBIT;
            pair_link(pairs, prev & ~TOPBIT)[prev >> 31] = cur;
            atomicAdd(&pair_link(pairs, cur)[2], 1u);
        }
    }
    grid.sync();
    {  // heavy chains: one warp per body (same sorter as k_sort_heavy)
        __shared__ unsigned long long wbuf[COOP_THREADS / 32][WSORT_SMEM];
        const uint32_t nh = min(__ldcg(&ctl->n_dheavy), ra.dheavy_cap);
        for (uint32_t hi = gtid >> 5; hi < nh; hi += gsz >> 5) {
            const uint32_t x = __ldcg(&ra.dheavy[hi]);
            warp_sort_and_link(ra.adj2 + __ldcg(&ra.offs2[x]), __ldcg(&ra.deg2[x]), wbuf[threadIdx.x >> 5], pairs);
        }
    }
    grid.sync();
    // ---- leave the dirty marks clean for the next round; per-sweep counters back to zero ----
    for (uint32_t i = gtid; i < nd; i += gsz) {
        const uint32_t x = __ldcg(&ra.dlist[i]);
        atomicAnd(&ra.dirty_bits[x >> 5], ~(1u << (x & 31u)));
    }
    if (gtid == 0) {
        ctl->n_contacts -= min(ctl->n_contacts, ctl->n_hits_cleared);
        ctl->frontier_n[0] = ctl->frontier_n[1] = ctl->frontier_n[2] = 0;
        ctl->n_repairs += 1;
        ctl->n_resweep += ndp;
    }
}

// after the repair sweep: the next k_requery decides again
__global__ void k_repair_done(Ctl* ctl) { ctl->repair_needed = 0; }

// Lazy candidates: a watch pair (between the inner and the outer margin) can only intersect at its turn if one
// of its bodies moved.  After a sweep, every watch pair whose SWEPT boxes (snapshot (+) reach) touch joins the sweep
// set and a repair sweep runs; pairs that stay out provably never intersect (exactness argument).
__global__ void __launch_bounds__(256)
    k_verify_watch(uint2* __restrict__ watch, uint32_t watch_cap, const float4* __restrict__ sbox,
                   const float* __restrict__ dmax, const float* __restrict__ reach,
                   const uint32_t* __restrict__ moved_bits, uint4* __restrict__ pairs, uint32_t pair_cap, Ctl* ctl) {
    const uint32_t nw = min(ctl->n_watch, watch_cap);
    for (uint32_t wi = blockIdx.x * blockDim.x + threadIdx.x; wi < nw; wi += gridDim.x * blockDim.x) {
        const uint2 wp = watch[wi];
        if (wp.x == NONE) continue;  // joined the sweep in an earlier round
        const uint32_t a = wp.x & ~TOPBIT, b = wp.y & ~TOPBIT;
        const bool ma = (moved_bits[a >> 5] >> (a & 31u)) & 1u, mb = (moved_bits[b >> 5] >> (b & 31u)) & 1u;
        if (!(ma || mb)) continue;
        const float ra = ma ? fmaxf(reach[a], dmax[a] * REACH_SLACK) : 0.0f;
        const float rb = mb ? fmaxf(reach[b], dmax[b] * REACH_SLACK) : 0.0f;
        const float4 alo = sbox[2 * (size_t)a], ahi = sbox[2 * (size_t)a + 1];
        const float4 blo = sbox[2 * (size_t)b], bhi = sbox[2 * (size_t)b + 1];
        if (!inflated_overlap(alo, ahi, blo, bhi, ra + rb)) continue;
        watch[wi].x = NONE;
        const uint4 rec = make_uint4(wp.x, wp.y, __float_as_uint(ahi.w), __float_as_uint(bhi.w));  // a = lower rank
        cg::coalesced_group cgp = cg::coalesced_threads();
        uint32_t base = 0;
        if (cgp.thread_rank() == 0) base = atomicAdd(&ctl->n_pairs, cgp.size());
        base = cgp.shfl(base, 0);
        const uint32_t slot = base + cgp.thread_rank();
        if (slot < pair_cap) {
            pairs[2 * (size_t)slot] = rec;
            pairs[2 * (size_t)slot + 1] = link_init();
        } else {
            atomicOr(&ctl->step_flags, SF_PAIR_OVERFLOW);
        }
        ctl->repair_needed = 1u;
        atomicAdd(&ctl->n_activated, 1u);
    }
}

// End of step: a last k_requery has run; if it still found pairs the step is inexact (async) or the host
// runs more repair rounds (sync).  Adapt the margin, reset per-step counters, clear escapee marks.
__global__ void k_finish_step(Ctl* ctl, Ctl* snapshot_out, int latch) {
    Ctl c = *ctl;
    if (c.repair_needed) c.step_flags |= SF_MARGIN;
    if (latch == 1) {  // async stepping: failures are latched, the host looks later
        if (c.step_flags &
            (SF_PAIR_OVERFLOW | SF_MARGIN | SF_HEAVY_OVERFLOW | SF_ESCAPEE_OVERFLOW | SF_INTERNAL | SF_EXCHANGE_OVERFLOW)) {
            c.steps_inexact += 1;
            c.sticky_flags |= c.step_flags;
        }
        c.steps_total += 1;
    }
    // the live set of an accepted step is the input of the next one (ghosts and departed bodies were dropped)
    if (latch == 1 || (latch == 0 && c.step_flags == 0)) c.n_in = c.n;
    c.unverified_total += c.n_unverified;
    *snapshot_out = c;  // what the host reads (stats of THIS step)
    if (latch == 0 && c.step_flags == SF_MARGIN) return;  // sync mode: the host continues the repair loop
    // (latch == 2: the host gave up on this step; just reset)
    // margin: grow when the step needed repair sweeps, relax slowly otherwise
    // a small margin keeps the candidate set (and every pair-proportional kernel) small; a repair sweep
    // costs about one more sweep.  Grow only when the pre-enqueued repair rounds were nearly used up.
    float m = c.margin;
    // bodies that out-run the outer margin cost a neighbourhood query each and a repair sweep; a repair sweep
    // is tolerated (watch-pair activations cause one anyway), a second one is not
    if (c.n_requeried && c.n_repairs > c.repair_tol + (c.n_activated ? 0u : 1u))
        m *= 1.25f;
    else if (c.n_requeried == 0)
        m *= 0.98f;
    c.margin = fminf(fmaxf(m, c.margin_min), c.margin_max);
    float mi = c.margin_in;
    if (c.n_repairs > c.repair_tol)
        mi = fmaxf(mi * 1.5f, 0.004f);   // too many repair sweeps: let more pairs into the first sweep
    else if (c.n_repairs < c.repair_tol)
        mi *= 0.9f;
    else if (c.repair_tol == 0)
        mi *= 0.97f;  // no repairs wanted at all: relax slowly
    c.margin_in = c.lazy ? fminf(mi, c.margin) : c.margin;
    c.grid = make_grid(c.lo, c.hi, c.ext_max, c.margin, c.cell_cfg, c.max_cells);
    c.n_pairs = 0, c.n_snapshot = 0, c.n_contacts = 0, c.n_rounds = 0, c.max_disp_bits = 0;
    c.frontier_n[0] = 0, c.frontier_n[1] = 0, c.frontier_n[2] = 0, c.n_heavy = 0, c.step_flags = 0;
    c.n_escapees = 0, c.max_abs_disp_bits = 0, c.repair_needed = 0, c.n_repairs = 0, c.n_pairs_swept = 0;
    c.n_pairs_k4 = 0, c.n_dirty = 0, c.n_dpairs = 0, c.dirty_bump = 0, c.n_dheavy = 0, c.dirty_all = 0;
    c.n_hits_cleared = 0, c.n_resweep = 0, c.dirty_level[0] = c.dirty_level[1] = c.dirty_level[2] = 0;
    c.n_watch = 0, c.n_activated = 0, c.n_requeried = 0;
    c.n_recv = 0, c.n_sent[0] = 0, c.n_sent[1] = 0, c.n_migrated_out = 0, c.n_migrated_in = 0, c.n_unverified = 0;
    c.taint_flag[0] = 0, c.taint_flag[1] = 0, c.taint_flag[2] = 0;
    *ctl = c;
}


}  // namespace gp

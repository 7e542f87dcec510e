// gp_repair_kernels.cuh — verification of the speculative pair set, component-local repair, end of step.
// Part of the physics step; see gp_kernels.cuh for the map of kernels to reference lines.
#pragma once
#include "gp_common.cuh"
#include "gp_mirror_kernels.cuh"
#include "gp_pairfind_kernels.cuh"
#include "gp_chain_kernels.cuh"
#include "gp_sweep_kernels.cuh"

namespace gp {
// Every function here is a PHASE of k_contacts (gp_contact_kernel.cuh): called by all threads of the cooperative grid,
// separated from the next phase by a grid barrier.  Arrays that are written inside the kernel are read through plain
// or .cg loads, never through the read-only path.

// ---------------------------------------------------------------------------------------------
// Repair loop of the speculative broad phase.
//  requery   : for every escapee x, find the bodies y with swept(x) ^ swept(y) != {} that are NOT
//              candidates yet, append those pairs, raise ctl->repair_needed.  One warp per escapee; the
//              grid rows its query box covers are contiguous runs of the cell-sorted AABB array.
//  k_tiny_*  : (gated) the usual repair: one warp per small connected component the new pairs touch re-executes
//              the component's pairs in the reference's order (see below).
//  k_repair_local : (gated, when a component is too large for that) undo and re-chain only the connected
//              components the new pairs touch; the repair sweep then covers those pairs only.
// When k_requery appends nothing the last sweep saw every pair that could intersect at its turn: the
// result equals the reference's full O(N^2) sweep.
// ---------------------------------------------------------------------------------------------
// Candidate predicate of the whole step: overlap(snap x, snap y, r_x + r_y), r_x = max(mu*extent(x), reach[x]).
// K4 built the set for reach = 0; a sweep in which x moved by dmax[x] > reach[x] extends it.  The pairs that
// pass with the grown reaches but not with the old ones are exactly the new ones (no duplicates).
constexpr float REACH_SLACK = 1.0001f;  // covers the rounding of the measured displacement

// The grid rows an escapee's query box covers are fetched by ONE lane each (all cell-table loads in flight at once),
// a warp scan flattens their slots into a list, and the lanes then test one candidate each: three dependent memory
// stages per escapee instead of three per ROW (a row holds one or two bodies, so a row-at-a-time walk left 30 of 32
// lanes idle; measured 119 us per step at 16 M bodies before).
constexpr uint32_t RQ_LIST = 224;
struct RequeryList {
    uint32_t q[RQ_LIST];
};

__device__ void requery_phase(const ContactArgs& a, RequeryList& rl) {
    Ctl* ctl = a.ctl;
    const float4* box = a.sbox;
    const float4* sbox = a.sbox;
    const uint32_t* LB = a.LB;
    const float *dmax = a.dmax, *reach = a.reach;
    const uint32_t *esc_bits = a.esc_bits, *escapees = a.escapees;
    uint4* pairs = a.pairs;
    const uint32_t esc_cap = a.esc_cap, pair_cap = a.pair_cap;
    const uint32_t ne = min(__ldcg(&ctl->n_escapees), esc_cap);
    if (ne == 0) return;
    const GridParams g = ctl->grid;
    const float mu = ctl->margin;
    const float ext_max = ctl->ext_max;
    const float partner_reach = fmaxf(mu * ext_max, __uint_as_float(__ldcg(&ctl->max_abs_disp_bits)) * REACH_SLACK);
    const uint32_t n_small = ctl->n_small, n = ctl->n;
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t ei = warp; ei < ne; ei += nwarps) {
        const uint32_t x = __ldcg(&escapees[ei]);
        const float rx_prev = __ldcg(&reach[x]);
        const float rx_now = __ldcg(&dmax[x]) * REACH_SLACK;
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
            const float ry_prev = __ldcg(&reach[y]), ry_now = __ldcg(&dmax[y]) * REACH_SLACK;
            const float rb_old = fmaxf(mb, ry_prev), rb_new = fmaxf(rb_old, ry_now);
            if (inflated_overlap(alo, ahi, blo, bhi, ra_old + rb_old)) return;   // found in an earlier round / by K4
            if (!inflated_overlap(alo, ahi, blo, bhi, ra_new + rb_new)) return;  // still out of reach
            // y is an escapee that grew too: its own query reports the pair (the lower slot wins)
            if (y < x && ry_now > ry_prev && ((__ldcg(&esc_bits[y >> 5]) >> (y & 31u)) & 1u)) return;
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
        const int ny = c1[1] - c0[1] + 1, nrows = ny * (c1[2] - c0[2] + 1);
        bool flat = nrows <= 32;  // (uniform: every lane derived the same box)
        uint32_t rq0 = 0, rlen = 0, roff = 0, T = 0;
        if (flat) {
            if (lane < nrows) {
                const int y = c0[1] + lane % ny, z = c0[2] + lane / ny;
                const uint32_t rowk = (uint32_t)g.dx * ((uint32_t)y + (uint32_t)g.dy * (uint32_t)z);
                rq0 = LB[rowk + (uint32_t)c0[0]];
                rlen = LB[rowk + (uint32_t)c1[0] + 1u] - rq0;
            }
            uint32_t inc = rlen;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            roff = inc - rlen;
            T = __shfl_sync(0xffffffffu, inc, 31);
            flat = T <= RQ_LIST;
        }
        if (flat) {
            __syncwarp();
            for (uint32_t k = 0; k < rlen; ++k) rl.q[roff + k] = rq0 + k;
            __syncwarp();
            for (uint32_t t = lane; t < T; t += 32) test(rl.q[t]);
            __syncwarp();
        } else {
            for (int z = c0[2]; z <= c1[2]; ++z)
                for (int y = c0[1]; y <= c1[1]; ++y) {
                    const uint32_t rowk = (uint32_t)g.dx * ((uint32_t)y + (uint32_t)g.dy * (uint32_t)z);
                    const uint32_t q0 = LB[rowk + (uint32_t)c0[0]], q1 = LB[rowk + (uint32_t)c1[0] + 1u];
                    for (uint32_t q = q0 + lane; q < q1; q += 32) test(q);
                }
        }
        for (uint32_t q = n_small + lane; q < n; q += 32) test(q);
    }
}

// ---------------------------------------------------------------------------------------------
// Component-local repair (a phase of k_contacts with grid barriers of its own).  The pairs that verification
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

__device__ void repair_local_phase(const ContactArgs& a, unsigned long long (*wbuf)[WSORT_SMEM], cg::grid_group& grid) {
    uint4* pairs = a.pairs;
    const uint32_t pair_cap = a.pair_cap, esc_cap = a.esc_cap;
    const uint32_t* offs = a.offs;
    const unsigned long long* adj = a.adj;
    uint32_t* chain_cnt = a.chain_cnt;
    const uint32_t* perm = a.perm;
    const Bodies B0 = a.B0, B1 = a.B1;
    const float4* A0 = a.A0;
    float *dmax = a.dmax, *reach = a.reach;
    const uint32_t* escapees = a.escapees;
    const RepairArrays ra = a.ra;
    Ctl* ctl = a.ctl;
    const float dt = a.dt, d16 = a.d16, d18 = a.d18, d35 = a.d35;
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    const uint32_t n = ctl->n;
    const uint32_t np0 = __ldcg(&ctl->n_pairs_k4), np1 = __ldcg(&ctl->n_pairs_swept),
                   np2 = min(__ldcg(&ctl->n_pairs), pair_cap);
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
        const uint32_t ne = min(__ldcg(&ctl->n_escapees), esc_cap);
        for (uint32_t i = gtid; i < ne; i += gsz) {
            const uint32_t x = __ldcg(&escapees[i]);
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

// ---------------------------------------------------------------------------------------------
// Tiny repair (the usual case).  Verification typically appends a few hundred pairs, each in a component of a handful
// of bodies.  k_tiny_walk: one WARP per new pair walks its component (K5 adjacency + per-body lists of the pairs
// appended since K4); the warp of the component's lowest new pair writes the component down.  k_tiny_exec: one warp
// per component restores the bodies, sorts the component's pairs by (rank_a, rank_b) -- the reference's order itself
// -- and one lane simply executes them in that order on a shared-memory copy of the bodies.  No chains, no rounds,
// no grid barriers.  If any component does not fit (more than TINY_MAX_BODIES bodies / TINY_MAX_PAIRS pairs) the walk
// raises tiny_fallback BEFORE anything was modified: k_tiny_exec does nothing and the general repair
// (k_repair_local + k_sweep) runs as before.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t TINY_MAX_BODIES = 64, TINY_MAX_PAIRS = 192;
constexpr uint32_t TINY_REC = 2 + TINY_MAX_BODIES + TINY_MAX_PAIRS + 30;  // u32 per component record (1152 B)

// per-body lists of the pairs appended since K4: xhead[body] -> entry (pid | side<<31), xnext[2*pid + side] -> next.
// Also: every escapee commits the reach of the sweep that ran (its query was answered by k_requery).
__device__ void tiny_link_phase(const ContactArgs& a) {
    Ctl* ctl = a.ctl;
    const uint4* pairs = a.pairs;
    uint32_t *xhead = a.xhead, *xnext = a.xnext;
    const float* dmax = a.dmax;
    float* reach = a.reach;
    const uint32_t* escapees = a.escapees;
    const uint32_t pair_cap = a.pair_cap, esc_cap = a.esc_cap, rec_cap = a.rec_cap;
    const uint32_t np1 = __ldcg(&ctl->n_pairs_swept), np2 = min(__ldcg(&ctl->n_pairs), pair_cap);
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    if (gtid == 0 && np2 - np1 > rec_cap) ctl->tiny_fallback = 1u;
    for (uint32_t j = np1 + gtid; j < np2; j += gsz) {
        const uint4 r = pairs[2 * (size_t)j];
        if (!(r.x & TOPBIT)) xnext[2 * (size_t)j] = atomicExch(&xhead[r.x], j);
        if (!(r.y & TOPBIT)) xnext[2 * (size_t)j + 1] = atomicExch(&xhead[r.y], j | TOPBIT);
    }
    const uint32_t ne = min(__ldcg(&ctl->n_escapees), esc_cap);
    for (uint32_t i = gtid; i < ne; i += gsz) {
        const uint32_t x = __ldcg(&escapees[i]);
        reach[x] = fmaxf(reach[x], __ldcg(&dmax[x]) * REACH_SLACK);
    }
}

struct TinyWalk {
    uint32_t bodies[TINY_MAX_BODIES];
    uint32_t pid[TINY_MAX_PAIRS];
};

__device__ __forceinline__ int tiny_find(const uint32_t* bodies, uint32_t nb, uint32_t x) {
    for (uint32_t k = 0; k < nb; ++k)
        if (bodies[k] == x) return (int)k;
    return -1;
}

__device__ void tiny_walk_phase(const ContactArgs& a, TinyWalk& t) {
    Ctl* ctl = a.ctl;
    const uint4* pairs = a.pairs;
    const uint32_t* offs = a.offs;
    const unsigned long long* adj = a.adj;
    const uint32_t *xhead = a.xhead, *xnext = a.xnext;
    uint32_t* recs = a.recs;
    const uint32_t pair_cap = a.pair_cap, rec_cap = a.rec_cap;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t np1 = __ldcg(&ctl->n_pairs_swept), np2 = min(__ldcg(&ctl->n_pairs), pair_cap);
    if (np2 - np1 > rec_cap) return;  // (k_tiny_link raised tiny_fallback)
    const uint32_t warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t inew = np1 + warp0; inew < np2; inew += nwarps) {
        __syncwarp();
        uint32_t nb = 0, npp = 0, min_new = inew;
        bool over = false;
        {
            const uint4 r = pairs[2 * (size_t)inew];
            if (lane == 0) {
                if (!(r.x & TOPBIT)) t.bodies[nb++] = r.x;
                if (!(r.y & TOPBIT) && (nb == 0 || t.bodies[0] != r.y)) t.bodies[nb++] = r.y;
            }
            nb = __shfl_sync(0xffffffffu, nb, 0);
        }
        __syncwarp();
        for (uint32_t fi = 0; fi < nb && !over; ++fi) {
            const uint32_t x = t.bodies[fi];
            // two sources of pairs of x: the K5 adjacency (pairs K4 emitted), then the list of later pairs
            const uint32_t e0 = offs[x], e1 = offs[x + 1];
            uint32_t xl = __ldcg(&xhead[x]);  // (same value in all lanes)
            for (uint32_t base = e0; !over; base += 32) {
                uint32_t ent = NONE;
                bool valid = false;
                if (base < e1) {  // adjacency chunk: one entry per lane
                    if (base + lane < e1) ent = (uint32_t)adj[base + lane], valid = true;
                } else if (xl != NONE) {  // one list entry at a time (lane 0 holds it)
                    if (lane == 0) ent = xl, valid = true;
                    xl = __ldcg(&xnext[2 * (size_t)(xl & ~TOPBIT) + (xl >> 31)]);
                } else {
                    break;
                }
                const uint32_t pid = ent & ~TOPBIT;
                uint4 r = make_uint4(0, 0, 0, 0);
                if (valid) r = pairs[2 * (size_t)pid];
                const bool x_is_b = (ent & TOPBIT) != 0;
                const uint32_t partner = x_is_b ? r.x : r.y;
                // a pair is collected from its first dynamic endpoint
                const bool emit = valid && !(x_is_b && !(r.x & TOPBIT));
                const uint32_t em = __ballot_sync(0xffffffffu, emit);
                if (npp + __popc(em) > TINY_MAX_PAIRS) {
                    over = true;
                    break;
                }
                if (emit) t.pid[npp + __popc(em & ((1u << lane) - 1u))] = pid;
                npp += __popc(em);
                if (valid && pid >= np1 && pid < min_new) min_new = pid;
                // new dynamic partners join the component (one at a time: the list stays duplicate-free)
                uint32_t cm = __ballot_sync(0xffffffffu, valid && !(partner & TOPBIT));
                while (cm) {
                    const int L = __ffs(cm) - 1;
                    cm &= cm - 1;
                    const uint32_t y = __shfl_sync(0xffffffffu, partner, L);
                    if (tiny_find(t.bodies, nb, y) >= 0) continue;
                    if (nb == TINY_MAX_BODIES) {
                        over = true;
                        break;
                    }
                    __syncwarp();
                    if (lane == 0) t.bodies[nb] = y;
                    ++nb;
                    __syncwarp();
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) min_new = min(min_new, __shfl_xor_sync(0xffffffffu, min_new, o));
        uint32_t* rec = recs + (size_t)(inew - np1) * TINY_REC;
        if (lane == 0) atomicMax(&ctl->dbg_tiny[4], nb), atomicMax(&ctl->dbg_tiny[5], npp);
        if (over) {
            if (lane == 0) ctl->tiny_fallback = 1u, rec[0] = 0u, atomicAdd(&ctl->dbg_tiny[3], 1u);
            continue;
        }
        __syncwarp();
        if (min_new != inew) {  // the warp of the component's lowest new pair owns it
            if (lane == 0) rec[0] = 0u;
            continue;
        }
        if (lane == 0) rec[0] = nb, rec[1] = npp;
        for (uint32_t k = lane; k < nb; k += 32) rec[2 + k] = t.bodies[k];
        for (uint32_t k = lane; k < npp; k += 32) rec[2 + TINY_MAX_BODIES + k] = t.pid[k];
    }
}

constexpr uint32_t TINY_STATICS = 8;
struct TinyExec {
    uint32_t bodies[TINY_MAX_BODIES];
    uint32_t pid[TINY_MAX_PAIRS], rx[TINY_MAX_PAIRS], ry[TINY_MAX_PAIRS], rw[TINY_MAX_PAIRS];
    uint32_t order[TINY_MAX_PAIRS];
    unsigned long long key[TINY_MAX_PAIRS];
    // where a / b live: < 128 index among `bodies`; 128 + i: cached static partner i; 255: static, read from global
    uint8_t la[TINY_MAX_PAIRS], lb[TINY_MAX_PAIRS];
    uint8_t hit[TINY_MAX_PAIRS];
    uint32_t sid[TINY_STATICS];  // the component's static partners (usually just the floor), cached like the bodies
    float4 SP[TINY_STATICS], SV[TINY_STATICS], SLo[TINY_STATICS], SHi[TINY_STATICS];
    float4 P[TINY_MAX_BODIES], V[TINY_MAX_BODIES], Lo[TINY_MAX_BODIES], Hi[TINY_MAX_BODIES], Snap[TINY_MAX_BODIES];
    float dm[TINY_MAX_BODIES];       // displacement record of this re-execution (track_move, kept on chip)
    uint8_t bflag[TINY_MAX_BODIES];  // 1: moved, 2: left its margin
};
constexpr int TINY_EXEC_WARPS = 8;  // warps of each CTA that execute components (the shared-memory copy is 6 KB per warp)

__device__ __forceinline__ float4 tiny_fetch(const float4* dyn, const float4* stat, const float4* glob, uint32_t l) {
    return l < 128u ? dyn[l] : l < 255u ? stat[l - 128u] : ldcg4(glob);
}

// track_move on the shared-memory copy (same comparisons, same NaN behaviour); global effects are applied afterwards
__device__ __forceinline__ void tiny_track(TinyExec& t, uint32_t l, const BodyRW& x, float& maxrel, float& maxabs) {
    const float4 lo_snap = t.Snap[l];
    const V3 lo_now = x.omin + x.pos;
    const float d = fmaxf(fabsf(lo_now.x - lo_snap.x), fmaxf(fabsf(lo_now.y - lo_snap.y), fabsf(lo_now.z - lo_snap.z)));
    if (d == 0.0f) return;
    const float e = fmaxf(x.omax.x - x.omin.x, fmaxf(x.omax.y - x.omin.y, x.omax.z - x.omin.z));
    maxrel = fmaxf(maxrel, d / e);
    if (!(d <= t.dm[l])) t.dm[l] = d;
    uint8_t f = t.bflag[l] | 1;
    if (!(d <= ESCAPE_FRAC * box_margin(lo_snap))) {
        f |= 2;
        maxabs = fmaxf(maxabs, d);
    }
    t.bflag[l] = f;
}

__device__ void tiny_exec_phase(const ContactArgs& a, TinyExec* tw) {
    if ((threadIdx.x >> 5) >= TINY_EXEC_WARPS) return;  // (no barrier inside this phase)
    Ctl* ctl = a.ctl;
    uint4* pairs = a.pairs;
    const uint32_t* recs = a.recs;
    const uint32_t* perm = a.perm;
    const Bodies B0 = a.B0, B1 = a.B1;
    const float4* A0 = a.A0;
    const float4* box = a.sbox;
    float *dmax = a.dmax, *reach = a.reach;
    uint32_t *moved_bits = a.moved_bits, *esc_bits = a.esc_bits, *escapees = a.escapees;
    const uint32_t pair_cap = a.pair_cap, esc_cap = a.esc_cap;
    const float dt = a.dt, d16 = a.d16, d18 = a.d18, d35 = a.d35;
    const uint32_t lane = threadIdx.x & 31;
    TinyExec& t = tw[threadIdx.x >> 5];
    const uint32_t np1 = __ldcg(&ctl->n_pairs_swept), np2 = min(__ldcg(&ctl->n_pairs), pair_cap);
    const uint32_t warp0 = blockIdx.x * TINY_EXEC_WARPS + (threadIdx.x >> 5), nwarps = gridDim.x * TINY_EXEC_WARPS;
    for (uint32_t inew = np1 + warp0; inew < np2; inew += nwarps) {
        const uint32_t* rec = recs + (size_t)(inew - np1) * TINY_REC;
        const uint32_t nb = __ldcg(&rec[0]);
        if (nb == 0) continue;
        const uint32_t npp = __ldcg(&rec[1]);
        __syncwarp();
        // ---- restore the bodies into shared memory, commit their reach, start a fresh displacement record ----
        for (uint32_t bl = lane; bl < nb; bl += 32) {
            const uint32_t x = __ldcg(&rec[2 + bl]);
            t.bodies[bl] = x;
            const uint32_t s0 = perm[x];
            float4 p = B0.P(s0), v = B0.V(s0);
            const float4 a = A0[s0];
            if (!(__float_as_uint(a.w) & F_INTEGRATED)) integrate_body(p, v, a, dt, d16, d18, d35);
            t.P[bl] = p, t.V[bl] = v, t.Lo[bl] = ldcg4(&B1.Lo(x)), t.Hi[bl] = ldcg4(&B1.Hi(x));
            t.Snap[bl] = box[2 * (size_t)x];
            const float dm = __ldcg(&dmax[x]);
            if (dm != 0.0f) reach[x] = fmaxf(__ldcg(&reach[x]), dm * REACH_SLACK);
            t.dm[bl] = 0.0f, t.bflag[bl] = 0;
        }
        __syncwarp();
        // ---- pairs: sort keys, local body indices, hit flags off ----
        uint32_t cleared = 0;
        for (uint32_t k = lane; k < npp; k += 32) {
            const uint32_t pid = __ldcg(&rec[2 + TINY_MAX_BODIES + k]);
            const uint4 r = __ldcg(&pairs[2 * (size_t)pid]);
            t.pid[k] = pid, t.rx[k] = r.x, t.ry[k] = r.y, t.rw[k] = r.w, t.hit[k] = 0;
            t.key[k] = ((unsigned long long)(r.z & ~TOPBIT) << 32) | (r.w & ~TOPBIT);
            t.la[k] = (r.x & TOPBIT) ? (uint8_t)255 : (uint8_t)tiny_find(t.bodies, nb, r.x);
            t.lb[k] = (r.y & TOPBIT) ? (uint8_t)255 : (uint8_t)tiny_find(t.bodies, nb, r.y);
            cleared += (r.w & TOPBIT) ? 1u : 0u;
        }
        __syncwarp();
        {   // static partners: de-duplicate (serially, there are very few) and cache them
            uint32_t ns = 0;
            for (uint32_t k0 = 0; k0 < npp; k0 += 32) {
                const uint32_t k = k0 + lane;
                uint32_t sx = NONE;  // (a pair has at most one static endpoint)
                bool a_side = false;
                if (k < npp) {
                    if (t.rx[k] & TOPBIT) sx = t.rx[k] & ~TOPBIT, a_side = true;
                    else if (t.ry[k] & TOPBIT) sx = t.ry[k] & ~TOPBIT;
                }
                uint32_t m = __ballot_sync(0xffffffffu, sx != NONE);
                while (m) {
                    const uint32_t L = (uint32_t)__ffs(m) - 1u;
                    m &= m - 1u;
                    const uint32_t y = __shfl_sync(0xffffffffu, sx, L);
                    int f = tiny_find(t.sid, ns, y);
                    __syncwarp();
                    if (f < 0 && ns < TINY_STATICS) {
                        if (lane == 0) t.sid[ns] = y;
                        f = (int)ns++;
                    }
                    __syncwarp();
                    if (lane == L) (a_side ? t.la[k] : t.lb[k]) = f < 0 ? (uint8_t)255 : (uint8_t)(128 + f);
                }
            }
            if (lane < ns) {
                const uint32_t y = t.sid[lane];
                t.SP[lane] = ldcg4(&B1.P(y)), t.SV[lane] = ldcg4(&B1.V(y));
                t.SLo[lane] = ldcg4(&B1.Lo(y)), t.SHi[lane] = ldcg4(&B1.Hi(y));
            }
        }
        __syncwarp();
        for (uint32_t k = lane; k < npp; k += 32) {  // rank of a key = its place in the reference's order
            const unsigned long long kk = t.key[k];
            uint32_t c = 0;
            for (uint32_t q = 0; q < npp; ++q) c += t.key[q] < kk;
            t.order[c] = k;
        }
        cleared = __reduce_add_sync(0xffffffffu, cleared);
        __syncwarp();
        // ---- one lane executes the component's pairs in the reference's order, entirely on chip (static partners
        //      are read-only in global memory) ----
        float maxd = 0.0f, maxabs = 0.0f;
        if (lane == 0) {
            uint32_t hits = 0;
            for (uint32_t oi = 0; oi < npp; ++oi) {
                const uint32_t k = t.order[oi];
                const uint32_t ia = t.rx[k] & ~TOPBIT, ib = t.ry[k] & ~TOPBIT;
                const uint32_t la = t.la[k], lb = t.lb[k];
                BodyRW ba, bb;
                ba.is_static = la >= 128u, bb.is_static = lb >= 128u;
                const float4 pa = tiny_fetch(t.P, t.SP, &B1.P(ia), la), pb = tiny_fetch(t.P, t.SP, &B1.P(ib), lb);
                const float4 loa = tiny_fetch(t.Lo, t.SLo, &B1.Lo(ia), la), hia = tiny_fetch(t.Hi, t.SHi, &B1.Hi(ia), la);
                const float4 lob = tiny_fetch(t.Lo, t.SLo, &B1.Lo(ib), lb), hib = tiny_fetch(t.Hi, t.SHi, &B1.Hi(ib), lb);
                ba.pos = xyz(pa), ba.mass = pa.w, ba.omin = xyz(loa), ba.omax = xyz(hia);
                bb.pos = xyz(pb), bb.mass = pb.w, bb.omin = xyz(lob), bb.omax = xyz(hib);
                const V3 amin = ba.omin + ba.pos, amax = ba.omax + ba.pos;
                const V3 bmin = bb.omin + bb.pos, bmax = bb.omax + bb.pos;
                const bool hit = amin.x < bmax.x && amax.x > bmin.x && amin.y < bmax.y && amax.y > bmin.y &&
                                 amin.z < bmax.z && amax.z > bmin.z;
                if (!hit) continue;
                ++hits;
                t.hit[k] = 1;
                const float4 va = tiny_fetch(t.V, t.SV, &B1.V(ia), la), vb = tiny_fetch(t.V, t.SV, &B1.V(ib), lb);
                ba.vel = xyz(va), ba.rest = va.w;
                bb.vel = xyz(vb), bb.rest = vb.w;
                const bool moved = resolve_pair(ba, bb, amin, amax, bmin, bmax);
                if (!ba.is_static) {
                    t.P[la] = make_float4(ba.pos.x, ba.pos.y, ba.pos.z, ba.mass);
                    t.V[la] = make_float4(ba.vel.x, ba.vel.y, ba.vel.z, ba.rest);
                    if (moved) tiny_track(t, la, ba, maxd, maxabs);
                }
                if (!bb.is_static) {
                    t.P[lb] = make_float4(bb.pos.x, bb.pos.y, bb.pos.z, bb.mass);
                    t.V[lb] = make_float4(bb.vel.x, bb.vel.y, bb.vel.z, bb.rest);
                    if (moved) tiny_track(t, lb, bb, maxd, maxabs);
                }
            }
            atomicAdd(&ctl->n_contacts, hits - cleared);  // (unsigned wrap-around = signed difference)
            atomicAdd(&ctl->n_resweep, npp);
            if (maxd > 0.0f) atomicMax(&ctl->max_disp_bits, __float_as_uint(maxd));
            if (maxabs > 0.0f) atomicMax(&ctl->max_abs_disp_bits, __float_as_uint(maxabs));
            atomicAdd(&ctl->dbg_tiny[2], 1u);
        }
        __syncwarp();
        // ---- results and their side effects go out in parallel ----
        for (uint32_t bl = lane; bl < nb; bl += 32) {
            const uint32_t x = t.bodies[bl];
            __stcg(&B1.P(x), t.P[bl]);
            __stcg(&B1.V(x), t.V[bl]);
            dmax[x] = t.dm[bl];
            const uint32_t bit = 1u << (x & 31u);
            const uint8_t f = t.bflag[bl];
            if (f & 1) atomicOr(&moved_bits[x >> 5], bit);
            if (f & 2) {
                const uint32_t prev = atomicOr(&esc_bits[x >> 5], bit);
                if (!(prev & bit)) {  // first escape of this body in this step
                    const uint32_t s = atomicAdd(&ctl->n_escapees, 1u);
                    if (s < esc_cap)
                        escapees[s] = x;
                    else
                        atomicOr(&ctl->step_flags, SF_ESCAPEE_OVERFLOW);
                }
            }
        }
        for (uint32_t k = lane; k < npp; k += 32) {
            const uint32_t w0 = t.rw[k], w1 = (w0 & ~TOPBIT) | (t.hit[k] ? TOPBIT : 0u);
            if (w1 != w0) pairs[2 * (size_t)t.pid[k]].w = w1;
        }
    }
}

// all components fitted: the repair is done (the general repair kernels behind this one become no-ops)
__device__ void tiny_done(Ctl* ctl, uint32_t pair_cap) {
    if (ctl->tiny_fallback == 0) {
        ctl->repair_needed = 0;
        ctl->n_repairs += 1;
        ctl->n_pairs_swept = min(ctl->n_pairs, pair_cap);
        ctl->dbg_tiny[0] += 1;
    } else {
        ctl->dbg_tiny[1] += 1;
    }
    ctl->tiny_fallback = 0;
}

// Lazy candidates: a watch pair (between the inner and the outer margin) can only intersect at its turn if one
// of its bodies moved.  After a sweep, every watch pair whose SWEPT boxes (snapshot (+) reach) touch joins the sweep
// set and a repair sweep runs; pairs that stay out provably never intersect (exactness argument).
__device__ void verify_watch_phase(const ContactArgs& a) {
    Ctl* ctl = a.ctl;
    uint2* watch = a.watch;
    const float4* sbox = a.sbox;
    const float *dmax = a.dmax, *reach = a.reach;
    const uint32_t* moved_bits = a.moved_bits;
    uint4* pairs = a.pairs;
    const uint32_t pair_cap = a.pair_cap, watch_cap = a.pair_cap;
    const uint32_t nw = min(ctl->n_watch, watch_cap);
    for (uint32_t wi = blockIdx.x * blockDim.x + threadIdx.x; wi < nw; wi += gridDim.x * blockDim.x) {
        const uint2 wp = __ldcg(&watch[wi]);
        if (wp.x == NONE) continue;  // joined the sweep in an earlier round
        const uint32_t xa = wp.x & ~TOPBIT, xb = wp.y & ~TOPBIT;
        const bool ma = (__ldcg(&moved_bits[xa >> 5]) >> (xa & 31u)) & 1u, mb = (__ldcg(&moved_bits[xb >> 5]) >> (xb & 31u)) & 1u;
        if (!(ma || mb)) continue;
        const float ra = ma ? fmaxf(__ldcg(&reach[xa]), __ldcg(&dmax[xa]) * REACH_SLACK) : 0.0f;
        const float rb = mb ? fmaxf(__ldcg(&reach[xb]), __ldcg(&dmax[xb]) * REACH_SLACK) : 0.0f;
        const float4 alo = sbox[2 * (size_t)xa], ahi = sbox[2 * (size_t)xa + 1];
        const float4 blo = sbox[2 * (size_t)xb], bhi = sbox[2 * (size_t)xb + 1];
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
__device__ void finish_step(Ctl* ctl, Ctl* snapshot_out, int latch) {
    Ctl c = *ctl;
    if (c.repair_needed) c.step_flags |= SF_MARGIN;
    if (latch == 1) {  // async stepping: failures are latched, the host looks later
        if (c.step_flags &
            (SF_PAIR_OVERFLOW | SF_MARGIN | SF_HEAVY_OVERFLOW | SF_ESCAPEE_OVERFLOW | SF_INTERNAL | SF_EXCHANGE_OVERFLOW |
             SF_PEER_TIMEOUT)) {
            c.steps_inexact += 1;
            c.sticky_flags |= c.step_flags;
        }
        c.steps_total += 1;
    }
    // the live set of an accepted step is the input of the next one (ghosts and departed bodies were dropped)
    if (latch == 1 || (latch == 0 && c.step_flags == 0)) c.n_in = c.n;
    c.unverified_total += c.n_unverified;
    c.dbg_counts[0] += c.n_escapees, c.dbg_counts[1] += c.n_requeried, c.dbg_counts[2] += c.n_activated;
    c.dbg_counts[3] += c.n_repairs;
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
    c.n_hits_cleared = 0, c.n_resweep = 0, c.tiny_fallback = 0, c.dirty_level[0] = c.dirty_level[1] = c.dirty_level[2] = 0;
    c.n_watch = 0, c.n_activated = 0, c.n_requeried = 0;
    c.n_recv = 0, c.n_sent[0] = 0, c.n_sent[1] = 0, c.n_migrated_out = 0, c.n_migrated_in = 0, c.n_unverified = 0;
    c.taint_flag[0] = 0, c.taint_flag[1] = 0, c.taint_flag[2] = 0;
    *ctl = c;
}
// stand-alone form: the host resets a step it gave up on (latch 2), or finishes a step without bodies
__global__ void k_finish_step(Ctl* ctl, Ctl* snapshot_out, int latch) { finish_step(ctl, snapshot_out, latch); }

}  // namespace gp

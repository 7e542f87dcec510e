// gp_cellsort_kernels.cuh — K1 cell keys + histogram (slab: migration / halo packing), K3 scatter into cell order, slab exchange helpers.
// Part of the physics step; see gp_kernels.cuh for the map of kernels to reference lines.
#pragma once
#include "gp_common.cuh"
#include "gp_mirror_kernels.cuh"

namespace gp {

// ---------------------------------------------------------------------------------------------
// Data layout.  The body state lives in two sets of five float4 arrays (P = pos|mass, V = vel|restitution,
// A = acc|flags, Lo = min offset|id, Hi = max offset|rank; 80 B/body).  Every step physically re-sorts the
// bodies into uniform-grid cell order: K1 computes the cell key of the integrated position, K2 sorts
// (key, slot), K3 gathers the previous set through the permutation, integrates on the fly and writes the
// next set + the world AABBs in cell order.  The contact sweep (K6) then works on bodies whose partners sit
// a few slots away: its accesses hit L2 instead of random DRAM sectors.  Bodies move a fraction of a cell per
// step, so the permutation is nearly the identity and the gather is nearly coalesced.
// ---------------------------------------------------------------------------------------------
// K1: integrate (registers only) + world AABB centre -> cell key.  R 80 B (P,V,A,Lo,Hi) + W 4 B (key).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t body_key(const GridParams& g, const float4 p, const float4 lo, const float4 hi,
                                             uint32_t flags) {
    if (flags & F_BIG) return g.ncells;
    const float lx = lo.x + p.x, ly = lo.y + p.y, lz = lo.z + p.z;
    const float hx = hi.x + p.x, hy = hi.y + p.y, hz = hi.z + p.z;
    return cell_key(g, (lx + hx) * 0.5f, (ly + hy) * 0.5f, (lz + hz) * 0.5f);
}

// Exchange buffer of a slab face: 16 B header {count, -, -, -} + records of five float4 (P, V, A, Lo, Hi) holding
// the POST-integration state, flagged F_INTEGRATED (+ F_GHOST for halo copies; without it the record transfers
// ownership = migration).
constexpr uint32_t XREC = 5;  // float4 per record

// K1 on a slab (SLAB = true) adds the multi-GPU duties: drop last step's ghosts, hand bodies whose centre left the
// slab to the neighbour (they stay here as ghosts for this step) and copy bodies near a face into the neighbour's
// halo.  A CTA walks K1_TILES consecutive tiles of 256 bodies and stages its outgoing records in shared memory (the
// bodies near a face are spread over almost every tile of the grid); it then reserves its slots with ONE global
// atomic per face (the send counter is a single address) and writes the records as one contiguous burst of 16 B
// stores, which matters when the buffer is the neighbour GPU's memory (peer-to-peer transport: contiguous stores of a
// warp travel as large NVLink packets).
constexpr int K1_TILES = 4;  // per CTA on a slab; a plain world uses one tile per CTA
constexpr uint32_t SEND_STAGE = 160;  // records per face a CTA can stage (about 2.5x the mean of uniform_16M on 8 GPUs with K1_TILES = 4)

__device__ __forceinline__ void stage_record(bool want, float4* __restrict__ stage, uint32_t* sm_cnt,
                                             float4* __restrict__ buf, uint32_t xcap, uint32_t* counter, Ctl* ctl,
                                             const float4 p, const float4 v, float4 a, const float4 lo, const float4 hi,
                                             uint32_t flags_out) {
    const uint32_t bal = __ballot_sync(0xffffffffu, want);
    if (!bal) return;
    const uint32_t lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(sm_cnt, __popc(bal));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (!want) return;
    const uint32_t s = base + __popc(bal & ((1u << lane) - 1u));
    a.w = __uint_as_float(flags_out);
    float4* r;
    if (s < SEND_STAGE) {
        r = stage + (size_t)s * XREC;
    } else {  // the stage is full (a very dense face): straight to the buffer
        const uint32_t g = atomicAdd(counter, 1u);
        if (g >= xcap) {
            atomicOr(&ctl->step_flags, SF_EXCHANGE_OVERFLOW);
            return;
        }
        r = buf + 1 + (size_t)g * XREC;
    }
    r[0] = p, r[1] = v, r[2] = a, r[3] = lo, r[4] = hi;
}

__device__ __forceinline__ void flush_stage(const float4* __restrict__ stage, uint32_t cnt, float4* __restrict__ buf,
                                            uint32_t xcap, uint32_t base, Ctl* ctl) {
    cnt = min(cnt, SEND_STAGE);
    if (base + cnt > xcap) {
        if (threadIdx.x == 0) atomicOr(&ctl->step_flags, SF_EXCHANGE_OVERFLOW);
        cnt = base < xcap ? xcap - base : 0u;
    }
    float4* dst = buf + 1 + (size_t)base * XREC;
    for (uint32_t e = threadIdx.x; e < cnt * XREC; e += blockDim.x) dst[e] = stage[e];
}

template <bool SLAB>
__global__ void __launch_bounds__(256)
    k_keys(const Bodies B0, float4* __restrict__ A, uint32_t* __restrict__ key, uint32_t* __restrict__ lrank,
           uint32_t* __restrict__ cell_cnt, Ctl* __restrict__ ctl, float4* __restrict__ send_l,
           float4* __restrict__ send_r, float dt, float d16, float d18, float d35) {
    __shared__ float4 stage[SLAB ? 2 : 1][SLAB ? SEND_STAGE * XREC : 1];
    __shared__ uint32_t sm_cnt[2], sm_base[2];
    const uint32_t n_in = ctl->n_in;
    constexpr int TILES = SLAB ? K1_TILES : 1;
    const uint32_t first = blockIdx.x * (256u * TILES);
    if (first >= n_in) return;  // whole CTA beyond the bodies (grids are sized for capacity)
    const GridParams g = ctl->grid;
    // slab parameters once, into registers (ctl is also written in this kernel: the compiler would re-load them)
    float x_lo = 0.f, x_hi = 0.f, halo = 0.f, mu = 0.f;
    bool has_l = false, has_r = false;
    uint32_t xcap_l = 0, xcap_r = 0;
    if (SLAB) {
        x_lo = ctl->x_lo, x_hi = ctl->x_hi, halo = ctl->halo, mu = ctl->margin;
        has_l = ctl->has_left != 0 && send_l != nullptr, has_r = ctl->has_right != 0 && send_r != nullptr;
        xcap_l = ctl->xcap[0], xcap_r = ctl->xcap[1];
        if (threadIdx.x < 2) sm_cnt[threadIdx.x] = 0;
        __syncthreads();
    }
#pragma unroll 1
    for (int t = 0; t < TILES; ++t) {
        const uint32_t i = first + t * 256u + threadIdx.x;
        const bool live = i < n_in;
        float4 p = make_float4(0, 0, 0, 0), v = p, a = p, lo = p, hi = p;
        uint32_t flags = 0, k = 0;
        if (live) {
            p = B0.P(i), v = B0.V(i), a = A[i], lo = B0.Lo(i), hi = B0.Hi(i);
            flags = __float_as_uint(a.w);
            if (!(flags & F_INTEGRATED)) integrate_body(p, v, a, dt, d16, d18, d35);
            k = body_key(g, p, lo, hi, flags);
        }
        if (SLAB) {
            bool to_l = false, to_r = false, own_l = false, own_r = false;
            if (live) {
                if (flags & F_GHOST) {
                    k = g.ncells + 1u;  // last step's halo copy: its owner sends a fresh one if it is still near
                } else if (!(flags & F_BIG)) {  // big (static) bodies are replicated on every rank and never travel
                    const float4 wlo = make_float4(lo.x + p.x, lo.y + p.y, lo.z + p.z, 0.f);
                    const float4 whi = make_float4(hi.x + p.x, hi.y + p.y, hi.z + p.z, 0.f);
                    const float cx = (wlo.x + whi.x) * 0.5f;
                    const float m = mu * box_extent(wlo, whi);  // the margin K3 will store for this body
                    own_l = has_l && cx < x_lo;
                    own_r = has_r && !(cx < x_hi);
                    to_l = own_l || (has_l && wlo.x - m < x_lo + halo);
                    to_r = own_r || (has_r && whi.x + m > x_hi - halo);
                }
            }
            const uint32_t fl = (flags | F_INTEGRATED) & ~F_GHOST;
            stage_record(to_l, stage[0], &sm_cnt[0], send_l, xcap_l, &ctl->n_sent[0], ctl, p, v, a, lo, hi,
                         own_l ? fl : (fl | F_GHOST));
            stage_record(to_r, stage[SLAB ? 1 : 0], &sm_cnt[1], send_r, xcap_r, &ctl->n_sent[1], ctl, p, v, a, lo, hi,
                         own_r ? fl : (fl | F_GHOST));
            if (own_l || own_r) {
                A[i].w = __uint_as_float(flags | F_GHOST);  // ownership moved: here it is a halo copy from now on
                atomicAdd(&ctl->n_migrated_out, 1u);
            }
        }
        if (live) {
            key[i] = k;
            // K2 is a one-pass radix sort whose single digit is the whole cell key: this is its histogram.  The bodies
            // arrive in last step's cell order, so the atomics of a warp fall on a few neighbouring L2 lines.  Dead
            // bodies (last step's halo copies) are dropped, not sorted: counting them would be ~10^5 atomics on ONE
            // address.
            if (k <= g.ncells) lrank[i] = atomicAdd(&cell_cnt[k], 1u);
        }
    }
    if (SLAB) {
        __syncthreads();
        if (threadIdx.x < 2) {
            const uint32_t c = min(sm_cnt[threadIdx.x], SEND_STAGE);
            sm_base[threadIdx.x] = c ? atomicAdd(&ctl->n_sent[threadIdx.x], c) : 0u;
        }
        __syncthreads();
        if (send_l) flush_stage(stage[0], sm_cnt[0], send_l, xcap_l, sm_base[0], ctl);
        if (send_r) flush_stage(stage[SLAB ? 1 : 0], sm_cnt[1], send_r, xcap_r, sm_base[1], ctl);
    }
}

// the counts travel in the buffer header
__global__ void k_seal_sends(float4* send_l, float4* send_r, const Ctl* ctl) {
    if (send_l) send_l[0] = make_float4(__uint_as_float(min(ctl->n_sent[0], ctl->xcap[0])), 0.f, 0.f, 0.f);
    if (send_r) send_r[0] = make_float4(__uint_as_float(min(ctl->n_sent[1], ctl->xcap[1])), 0.f, 0.f, 0.f);
}

// ---- peer-to-peer transport (one process per GPU, NVLink): K1 writes the records STRAIGHT into the neighbour's
// receive buffer (peer-mapped through CUDA IPC), so the transfer overlaps K1 itself and no copy kernel runs.
// Synchronisation is a pair of step counters per face in the receiver's memory, written with release.sys stores and
// polled with acquire.sys loads (bounded: a library kernel never spins forever).  Receive buffers are double-buffered
// by step parity; `credit` tells the sender that the buffer of two steps ago has been consumed.
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// wait until *f >= want for both flags (nullptr = no neighbour on that side).  Bounded (a library kernel never spins
// forever), but generously: ranks are separate processes, and ordinary host skew (a download, a GC pause, late
// enqueueing) must not expire it.  An expired wait is FATAL for the world: records that were migrating in the missed
// message are gone, so peer_dead is latched, k_append appends nothing from then on, and gp_sync reports
// GP_ERR_PEER_TIMEOUT until the caller uploads again.
__global__ void k_wait_flags(const uint32_t* f0, const uint32_t* f1, uint32_t want, Ctl* ctl,
                             unsigned long long timeout_ns) {
    const uint32_t* f = threadIdx.x == 0 ? f0 : f1;
    if (threadIdx.x > 1 || !f) return;
    if (*(volatile uint32_t*)&ctl->peer_dead) return;
    const unsigned long long t0 = globaltimer_ns();
    for (;;) {
        if (ld_acquire_sys(f) >= want) return;
        __nanosleep(200);
        if (globaltimer_ns() - t0 > timeout_ns) break;
    }
    atomicOr(&ctl->step_flags, SF_PEER_TIMEOUT);
    atomicExch(&ctl->peer_dead, 1u);
}
__global__ void k_post_flags(uint32_t* f0, uint32_t* f1, uint32_t v) {
    __threadfence_system();
    if (f0) st_release_sys(f0, v);
    if (f1) st_release_sys(f1, v);
}
// header (count) into the neighbour's buffer, then the ready flag; all of K1's record stores precede this kernel
__global__ void k_seal_p2p(float4* peer_l, float4* peer_r, uint32_t* ready_l, uint32_t* ready_r, uint32_t step,
                           const Ctl* ctl) {
    if (peer_l) peer_l[0] = make_float4(__uint_as_float(min(ctl->n_sent[0], ctl->xcap[0])), 0.f, 0.f, 0.f);
    if (peer_r) peer_r[0] = make_float4(__uint_as_float(min(ctl->n_sent[1], ctl->xcap[1])), 0.f, 0.f, 0.f);
    __threadfence_system();
    if (ready_l) st_release_sys(ready_l, step);
    if (ready_r) st_release_sys(ready_r, step);
}

// received records join the input set behind the bodies that were already here; their cell keys enter the
// histogram before the scan
__global__ void __launch_bounds__(256)
    k_append(const float4* __restrict__ recv_l, const float4* __restrict__ recv_r, const Bodies B0,
             float4* __restrict__ A, uint32_t* __restrict__ key, uint32_t* __restrict__ lrank,
             uint32_t* __restrict__ cell_cnt, Ctl* ctl) {
    if (ctl->peer_dead) {  // a neighbour timed out: whatever is in the buffers is not this step's message
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            ctl->n_recv = 0;
            atomicOr(&ctl->step_flags, SF_PEER_TIMEOUT);
        }
        return;
    }
    const uint32_t cl = recv_l ? min(__float_as_uint(__ldcg(recv_l).x), ctl->xcap[0]) : 0u;
    const uint32_t cr = recv_r ? min(__float_as_uint(__ldcg(recv_r).x), ctl->xcap[1]) : 0u;
    const uint32_t n_in = ctl->n_in;
    uint32_t total = cl + cr;
    if (n_in + total > ctl->cap) {
        total = ctl->cap - n_in;
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&ctl->step_flags, SF_EXCHANGE_OVERFLOW);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) ctl->n_recv = total;
    const GridParams g = ctl->grid;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const float4* r = (t < cl) ? recv_l + 1 + (size_t)t * XREC : recv_r + 1 + (size_t)(t - cl) * XREC;
        // (L2 loads: the buffer may have been written by the neighbour GPU)
        const float4 p = __ldcg(r), v = __ldcg(r + 1), a = __ldcg(r + 2), lo = __ldcg(r + 3), hi = __ldcg(r + 4);
        const uint32_t s = n_in + t;
        B0.P(s) = p, B0.V(s) = v, A[s] = a, B0.Lo(s) = lo, B0.Hi(s) = hi;
        const uint32_t flags = __float_as_uint(a.w);
        if (!(flags & F_GHOST)) atomicAdd(&ctl->n_migrated_in, 1u);
        const uint32_t k = body_key(g, p, lo, hi, flags);  // the record is already integrated
        key[s] = k;
        lrank[s] = atomicAdd(&cell_cnt[k], 1u);
    }
}

// ---------------------------------------------------------------------------------------------
// K2: exclusive scan of the cell histogram in place (scan.cuh) -> LB[c] = first slot of cell c, for
// c in [0, ncells+2] (key ncells = big list, key ncells+1 = dead); LB[ncells+2] = bodies entering the step.
// K3: scatter the bodies into cell order (slot = LB[key] + arrival rank within the cell), integrate, emit
// world AABBs.  R 8 (key, rank) + R 80 (own record, coalesced) + R 4 (cell start) + W 80 (next set) + W 32 (AABB)
// + W 4 (perm, scattered) + W 16 (per-slot constants, coalesced).
// The order inside a cell is the arrival order of the atomics: results do not depend on it (the contact
// sweep orders by the reference's iteration rank, never by slot).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    k_scatter_cells(const uint32_t* __restrict__ key, const uint32_t* __restrict__ lrank,
                    const Bodies B0, const float4* __restrict__ A0, const Bodies B1, float4* __restrict__ A1,
                    float4* __restrict__ sbox, uint32_t* __restrict__ chain_cnt, uint32_t* __restrict__ perm,
                    float* __restrict__ dmax, float* __restrict__ reach, uint32_t* __restrict__ moved_bits,
                    uint32_t* __restrict__ esc_bits, uint32_t* __restrict__ xhead, const uint32_t* __restrict__ LB,
                    Ctl* ctl, uint32_t* __restrict__ slot_of /* nullable: body id -> slot, kept for dirty-set updates */,
                    uint32_t* __restrict__ seed_bits /* nullable (slab worlds): seeds of the exactness proof */,
                    uint32_t* __restrict__ uf_parent /* nullable (slab worlds): the proof's union-find, reset here */,
                    float dt, float d16, float d18, float d35) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t ncells = ctl->grid.ncells;
    if (i == 0) {
        ctl->n_small = LB[ncells];
        ctl->n = LB[ncells + 1u];
    }
    if (i >= ctl->n_in + ctl->n_recv) return;
    // Per-slot state that starts every step at a constant: written by INDEX, not by destination slot -- the live
    // slots [0, n) are a prefix of the indices handled here, and a coalesced store costs a fraction of a scattered
    // 4-byte one (partial sectors).  xhead: no pair appended after K4 touches this body yet (k_tiny_link).
    chain_cnt[i] = 0, dmax[i] = 0.0f, reach[i] = 0.0f, xhead[i] = NONE;
    if (uf_parent) uf_parent[i] = i;  // every body its own component (saves the proof a pass and a grid barrier)
    if ((i & 31u) == 0u) moved_bits[i >> 5] = 0u, esc_bits[i >> 5] = 0u;
    const uint32_t k = key[i];
    if (k > ncells) return;  // dead (a ghost of the last step, a body that migrated away): dropped here
    const uint32_t s = LB[k] + lrank[i];
    float4 p, v, lo, hi, a = A0[i];
    // 256-bit loads / stores (LDG/STG.E.ENL2.256): the 64 B body record moves in two instructions, the 32 B AABB
    // record in one, whole sectors each (measured: 0.75 -> 0.60 ms at 16 M bodies against 128-bit accesses)
    ld256(&B0.P(i), p, v);
    ld256(&B0.Lo(i), lo, hi);
    uint32_t flags = __float_as_uint(a.w);
    if (flags & F_INTEGRATED) {
        flags &= ~F_INTEGRATED;
        a.w = __uint_as_float(flags);
    } else {
        integrate_body(p, v, a, dt, d16, d18, d35);
    }
    A1[s] = a;
    st256(&B1.P(s), p, v);
    st256(&B1.Lo(s), lo, hi);
    // world AABB = offsets + position (physics.rs:117: corner + pos; min/max commute with the rounding)
    // AABB record: lo.w = the body's outer margin fl(mu * extent), sign bit = static; hi.w = rank
    const float4 wlo = make_float4(lo.x + p.x, lo.y + p.y, lo.z + p.z, 0.0f);
    const float4 whi = make_float4(hi.x + p.x, hi.y + p.y, hi.z + p.z, hi.w);
    const float m = ctl->margin * box_extent(wlo, whi);
    const float4 wlo_m = make_float4(wlo.x, wlo.y, wlo.z,
                                     __uint_as_float((__float_as_uint(m) & ~TOPBIT) | ((flags & F_STATIC) ? TOPBIT : 0u)));
    st256(&sbox[2 * (size_t)s], wlo_m, whi);
    // Slab worlds: a body whose margin pokes beyond what the neighbour mirrors may have a partner this rank has never
    // seen -- a seed of the per-step exactness proof (taint_phase).  Decided here, where the box is in registers, so
    // the proof need not read 32 B per body again; bodies that out-run their margin are added from the escapee list.
    if (seed_bits && !(flags & F_STATIC)) {
        const bool out_r = ctl->has_right && !(whi.x + m <= ctl->x_hi + ctl->halo);
        const bool out_l = ctl->has_left && !(wlo.x - m >= ctl->x_lo - ctl->halo);
        if (out_l || out_r) atomicOr(&seed_bits[s >> 5], 1u << (s & 31u));
    }
    perm[s] = i;  // slot in the input set: the repair kernels re-integrate from there
    if (slot_of) slot_of[__float_as_uint(lo.w)] = s;  // (only for worlds whose host sends per-index updates)
}


}  // namespace gp

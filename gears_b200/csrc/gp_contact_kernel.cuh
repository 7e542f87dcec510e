// gp_contact_kernel.cuh — K6: the whole contact phase of a step as ONE persistent cooperative kernel.
// Part of the physics step; see gp_kernels.cuh for the map of kernels to reference lines.
//
//   sweep      the ordered contact sweep (check_and_resolve_collision, physics.rs:474-499) over the dependency levels
//   verify     k_requery-style neighbourhood queries of the escapees + the watch list (speculative broad phase)
//   repair     while verification found pairs: per-component warp repair, or the component-local cooperative repair
//              + a list-mode sweep when a component is too large; then verify again
//   taint      (slab worlds) the per-step exactness proof
//   finish     margin adaptation, counters, stats snapshot
//
// Round 1 ran these as 21-22 launches per step, five of them cooperative, most of them gated no-ops that were
// enqueued "just in case" (a cooperative launch costs ~25 us even when it returns at once; measured with
// GP_PROF_FINE=1: 0.19 ms of no-ops and launch gaps per step at 16 M bodies, 0.15 ms at 2 M).  Here the phases are
// device functions separated by grid barriers, the repair loop runs only while it is needed, and the tail of the
// sweep (levels with a handful of pairs) is run by ONE CTA without grid barriers.
#pragma once
#include "gp_common.cuh"
#include "gp_sweep_kernels.cuh"
#include "gp_repair_kernels.cuh"
#include "gp_slab_kernels.cuh"

namespace gp {

// ---------------------------------------------------------------------------------------------
// The sweep.  A frontier holds the pairs that are at the head of both chains.  Per pair: exact intersects with
// CURRENT positions (physics.rs:482-489), resolve (physics.rs:176-299), then pop both chains; a pair whose pending
// count drops to 0 joins the next frontier.  One grid barrier per dependency level while a level has work for the
// whole grid; levels with at most TAIL_MAX pairs are run by CTA 0 alone (block barriers only).
// skip (first sweep only): a pair without the S_snapshot flag whose two bodies have not been moved by an earlier
// pair of this sweep is tested on exactly the positions K4 tested (same operands, same comparisons) -- it misses
// again, and its bodies need not be fetched at all (84 % of the candidates of the uniform scene).  "Moved" = the bit
// in moved_bits, which the first sweep sets for EVERY position write (track_move alone sets it only when the box
// corner changed, which can differ by rounding).
// ---------------------------------------------------------------------------------------------
constexpr uint32_t TAIL_MAX = 2 * COOP_THREADS;

struct SweepAcc {
    uint32_t contacts = 0;
    float maxd = 0.0f, maxabs = 0.0f;
};

// test + resolve of one pair at its turn.  LEVEL0: the pair is at the head of all its chains, so no earlier pair of this
// sweep has touched its bodies: without the S_snapshot flag it misses again, whatever moved_bits says.
// (Loads are staged -- moved bits, then boxes, then velocities only for a hit: fetching everything a contact could need
// in one round trip was measured slower, the extra live registers spill at 768 threads per SM.)
template <bool SKIP, bool LEVEL0>
__device__ __forceinline__ void pair_execute(const ContactArgs& a, FrontierStage& fs, uint32_t pid, const uint4 rec,
                                             SweepAcc& acc) {
    const Bodies B = a.B1;
    const uint32_t ia = rec.x & ~TOPBIT, ib = rec.y & ~TOPBIT;
    if (SKIP && !(rec.z & TOPBIT)) {
        if (LEVEL0) return;
        const uint32_t wa = __ldcg(&a.moved_bits[ia >> 5]), wb = __ldcg(&a.moved_bits[ib >> 5]);
        if (!(((wa >> (ia & 31u)) | (wb >> (ib & 31u))) & 1u)) return;
    }
    BodyRW x, y;
    x.is_static = (rec.x & TOPBIT) != 0;
    y.is_static = (rec.y & TOPBIT) != 0;
    const float4 pa = ldcg4(&B.P(ia)), pb = ldcg4(&B.P(ib));
    const float4 loa = ldcg4(&B.Lo(ia)), hia = ldcg4(&B.Hi(ia)), lob = ldcg4(&B.Lo(ib)), hib = ldcg4(&B.Hi(ib));
    x.pos = xyz(pa), x.mass = pa.w, x.omin = xyz(loa), x.omax = xyz(hia);
    y.pos = xyz(pb), y.mass = pb.w, y.omin = xyz(lob), y.omax = xyz(hib);
    const V3 amin = x.omin + x.pos, amax = x.omax + x.pos;
    const V3 bmin = y.omin + y.pos, bmax = y.omax + y.pos;
    // physics.rs:159-164 on the CURRENT positions
    const bool hit = amin.x < bmax.x && amax.x > bmin.x && amin.y < bmax.y && amax.y > bmin.y && amin.z < bmax.z &&
                     amax.z > bmin.z;
    if (!hit) return;
    ++acc.contacts;
    const float4 va = ldcg4(&B.V(ia)), vb = ldcg4(&B.V(ib));
    x.vel = xyz(va), x.rest = va.w;
    y.vel = xyz(vb), y.rest = vb.w;
    const bool moved = resolve_pair(x, y, amin, amax, bmin, bmax);
    if (!x.is_static) {
        __stcg(&B.P(ia), make_float4(x.pos.x, x.pos.y, x.pos.z, x.mass));
        __stcg(&B.V(ia), make_float4(x.vel.x, x.vel.y, x.vel.z, x.rest));
    }
    if (!y.is_static) {
        __stcg(&B.P(ib), make_float4(y.pos.x, y.pos.y, y.pos.z, y.mass));
        __stcg(&B.V(ib), make_float4(y.vel.x, y.vel.y, y.vel.z, y.rest));
    }
    if (moved) {
        if (!x.is_static)
            track_move(x, ia, __ldcg(&a.sbox[2 * (size_t)ia]), __ldcg(&a.dmax[ia]), a.dmax, a.moved_bits, a.esc_bits,
                       a.escapees, a.esc_cap, a.ctl, &fs, acc.maxd, acc.maxabs);
        if (!y.is_static)
            track_move(y, ib, __ldcg(&a.sbox[2 * (size_t)ib]), __ldcg(&a.dmax[ib]), a.dmax, a.moved_bits, a.esc_bits,
                       a.escapees, a.esc_cap, a.ctl, &fs, acc.maxd, acc.maxabs);
        if (SKIP) {  // every position write counts (see above)
            if (!x.is_static) atomicOr(&a.moved_bits[ia >> 5], 1u << (ia & 31u));
            if (!y.is_static) atomicOr(&a.moved_bits[ib >> 5], 1u << (ib & 31u));
        }
    }
    a.pairs[2 * (size_t)pid].w = rec.w | TOPBIT;  // S_sweep flag; the repair: these bodies were written
}
// wake the successors on both chains once they are at the head everywhere (static endpoints have no chain: their
// link stays NONE)
__device__ __forceinline__ void pair_wake(const ContactArgs& a, FrontierStage& fs, const uint4 link, uint32_t* fout,
                                          uint32_t* fcount) {
    if (link.x != NONE && atomicSub(&pair_link(a.pairs, link.x)[2], 1u) == 1u) stage_push(fs, link.x, fout, fcount);
    if (link.y != NONE && atomicSub(&pair_link(a.pairs, link.y)[2], 1u) == 1u) stage_push(fs, link.y, fout, fcount);
}

__device__ __forceinline__ void flush(const ContactArgs& a, FrontierStage& fs, uint32_t* fout, uint32_t* fcount) {
    stage_flush(fs, fout, fcount, a.escapees, a.esc_cap, a.ctl);
}

// The items of one level: thread t of nthreads, two items per iteration (both records are fetched before either pair
// is processed).  No barrier inside: the warps of a CTA drift apart and hide each other's dependent loads; a push that
// finds the CTA's stage full goes straight to the global list.
template <bool SKIP, bool LEVEL0>
__device__ __forceinline__ void sweep_items(const ContactArgs& a, FrontierStage& fs, const uint32_t* fin, uint32_t nf,
                                            uint32_t t, uint32_t nthreads, uint32_t* fout, uint32_t* fcount,
                                            SweepAcc& acc) {
    for (uint32_t f = t; f < nf; f += 2u * nthreads) {
        const uint32_t f2 = f + nthreads;
        const bool two = f2 < nf;
        const uint32_t p0 = __ldcg(&fin[f]), p1 = two ? __ldcg(&fin[f2]) : 0u;
        const uint4 r0 = __ldcg(&a.pairs[2 * (size_t)p0]), l0 = __ldcg(&a.pairs[2 * (size_t)p0 + 1]);
        uint4 r1 = r0, l1 = l0;
        if (two) r1 = __ldcg(&a.pairs[2 * (size_t)p1]), l1 = __ldcg(&a.pairs[2 * (size_t)p1 + 1]);
        pair_execute<SKIP, LEVEL0>(a, fs, p0, r0, acc);
        pair_wake(a, fs, l0, fout, fcount);
        if (two) {
            pair_execute<SKIP, LEVEL0>(a, fs, p1, r1, acc);
            pair_wake(a, fs, l1, fout, fcount);
        }
    }
}

// LIST: repair sweep over the pairs the component-local repair collected (ra.dpairs); SKIP: see above
// Returns the number of pairs whose edges (with the whole watch list) the idle CTAs united for the slab proof while
// CTA 0 ran tail levels (first sweep of a slab world only; 0: none).
template <bool LIST, bool SKIP>
__device__ uint32_t sweep_phase(const ContactArgs& a, FrontierStage& fs, cg::grid_group& grid, PhaseClock& pc) {
    Ctl* ctl = a.ctl;
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    if (threadIdx.x == 0) fs.cnt = 0, fs.ecnt = 0;
    __syncthreads();
    const uint32_t np = min(__ldcg(&ctl->n_pairs), a.pair_cap);
    SweepAcc acc;
    uint32_t united = 0;
    // Level 0 = the pairs at the head of all their chains, found by one coalesced scan of the link records.  First
    // sweep: two thirds of them are linked to nothing and five in six carry no S_snapshot flag -- with the sweep
    // shortcut such a pair needs neither a test nor a wake-up, so it is not even listed.
    const uint32_t nl = LIST ? __ldcg(&ctl->n_dpairs) : np;
    for (uint32_t li = gtid; li < nl; li += gsz) {
        const uint32_t pid = LIST ? __ldcg(&a.ra.dpairs[li]) : li;
        const uint4 link = __ldcg(&a.pairs[2 * (size_t)pid + 1]);
        if (link.z != 0) continue;
        if (!LIST && SKIP && (link.x & link.y) == NONE && !(__ldcg(&a.pairs[2 * (size_t)pid].z) & TOPBIT)) continue;
        stage_push(fs, pid, a.frontier0, &ctl->frontier_n[0]);
    }
    flush(a, fs, a.frontier0, &ctl->frontier_n[0]);
    grid.sync();
    pc.stamp(ctl, LIST ? 8 : 0, LIST ? -1 : 0, 0);
    uint32_t rounds = 0;
    while (true) {
        const uint32_t ci = rounds % 3u, co = (rounds + 1u) % 3u, cz = (rounds + 2u) % 3u;
        const uint32_t nf = __ldcg(&ctl->frontier_n[ci]);
        if (nf == 0) break;
        if (nf <= TAIL_MAX) {
            // ---- tail: CTA 0 runs the levels alone (block barriers) until the frontier is empty or large again ----
            if (blockIdx.x == 0) {
                uint32_t r = rounds, n_now = nf;
                while (n_now != 0 && n_now <= TAIL_MAX) {
                    const uint32_t* fin = (r & 1u) ? a.frontier1 : a.frontier0;
                    uint32_t* fout = (r & 1u) ? a.frontier0 : a.frontier1;
                    // the stage holds at most 2 * TAIL_MAX <= FB_CAP pushes: the overflow path is never taken
                    if (!LIST && r == 0)
                        sweep_items<SKIP, true>(a, fs, fin, n_now, threadIdx.x, blockDim.x, fout, nullptr, acc);
                    else
                        sweep_items<SKIP, false>(a, fs, fin, n_now, threadIdx.x, blockDim.x, fout, nullptr, acc);
                    __syncthreads();
                    const uint32_t c = fs.cnt;
                    for (uint32_t i = threadIdx.x; i < c; i += blockDim.x) __stcg(&fout[i], fs.buf[i]);
                    __syncthreads();
                    if (threadIdx.x == 0) fs.cnt = 0;
                    n_now = c;
                    ++r;
                    __syncthreads();
                }
                if (threadIdx.x == 0) {
                    ctl->frontier_n[r % 3u] = n_now;
                    ctl->frontier_n[(r + 1u) % 3u] = 0u;
                    ctl->frontier_n[(r + 2u) % 3u] = 0u;
                    ctl->sweep_round = r;
                    ctl->dbg_tail_rounds += r - rounds;
                }
            } else if (!LIST && a.taint_bits && united == 0u) {
                proof_unions(a, 0u, np, true, (blockIdx.x - 1u) * blockDim.x + threadIdx.x,
                             (gridDim.x - 1u) * blockDim.x);
            }
            if (!LIST && a.taint_bits && gridDim.x > 1u) united = np;  // (uniform: every CTA takes this path)
            grid.sync();
            pc.stamp(ctl, LIST ? 8 : 2);
            rounds = __ldcg(&ctl->sweep_round);
            continue;
        }
        // counter cz was consumed in the previous round (all reads are behind that round's barrier)
        // and is next filled in the round after this one (behind this round's barrier)
        if (gtid == 0) ctl->frontier_n[cz] = 0;
        const uint32_t* fin = (rounds & 1u) ? a.frontier1 : a.frontier0;
        uint32_t* fout = (rounds & 1u) ? a.frontier0 : a.frontier1;
        if (!LIST && rounds == 0)
            sweep_items<SKIP, true>(a, fs, fin, nf, gtid, gsz, fout, &ctl->frontier_n[co], acc);
        else
            sweep_items<SKIP, false>(a, fs, fin, nf, gtid, gsz, fout, &ctl->frontier_n[co], acc);
        flush(a, fs, fout, &ctl->frontier_n[co]);
        ++rounds;
        grid.sync();
        pc.stamp(ctl, LIST ? 8 : 1, LIST ? -1 : (int)rounds, nf);
    }
    flush(a, fs, a.frontier0, &ctl->frontier_n[0]);  // (escapees staged by the tail levels; the frontier stage is empty)
    // per-thread partials -> warp -> CTA (the stage is idle now) -> ONE atomic per CTA and counter
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc.contacts += __shfl_xor_sync(0xffffffffu, acc.contacts, o);
        acc.maxd = fmaxf(acc.maxd, __shfl_xor_sync(0xffffffffu, acc.maxd, o));
        acc.maxabs = fmaxf(acc.maxabs, __shfl_xor_sync(0xffffffffu, acc.maxabs, o));
    }
    __syncthreads();
    if (threadIdx.x < 3) fs.buf[threadIdx.x] = 0u;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {  // (positive floats order like their bit patterns)
        if (acc.contacts) atomicAdd(&fs.buf[0], acc.contacts);
        if (acc.maxd > 0.0f) atomicMax(&fs.buf[1], __float_as_uint(acc.maxd));
        if (acc.maxabs > 0.0f) atomicMax(&fs.buf[2], __float_as_uint(acc.maxabs));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (fs.buf[0]) atomicAdd(&ctl->n_contacts, fs.buf[0]);
        if (fs.buf[1]) atomicMax(&ctl->max_disp_bits, fs.buf[1]);
        if (fs.buf[2]) atomicMax(&ctl->max_abs_disp_bits, fs.buf[2]);
    }
    if (gtid == 0) {
        ctl->n_rounds = max(ctl->n_rounds, rounds);
        ctl->n_pairs_swept = np;
        if (!LIST) ctl->n_pairs_k4 = np;
        ctl->frontier_n[0] = ctl->frontier_n[1] = ctl->frontier_n[2] = 0;
    }
    return united;
}

// ---------------------------------------------------------------------------------------------
// The kernel.  start_phase 0: a fresh step (sweep first); 1: the host continues the repair loop of a synchronous step
// whose candidate set was still growing after max_rounds rounds.
// ---------------------------------------------------------------------------------------------
union ContactSmem {
    FrontierStage fs;
    TinyWalk walk[COOP_THREADS / 32];
    TinyExec exec[TINY_EXEC_WARPS];
    unsigned long long wbuf[COOP_THREADS / 32][WSORT_SMEM];
    RequeryList rq[COOP_THREADS / 32];
};

__global__ void __launch_bounds__(COOP_THREADS, 1) k_contacts(const ContactArgs a) {
    extern __shared__ __align__(16) unsigned char contact_smem_raw[];
    ContactSmem& sm = *reinterpret_cast<ContactSmem*>(contact_smem_raw);
    cg::grid_group grid = cg::this_grid();
    Ctl* ctl = a.ctl;
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
    PhaseClock pc;
    pc.start(a.profile != 0);
    uint32_t united = 0;
    if (a.start_phase == 0) {
        united = sweep_phase<false, true>(a, sm.fs, grid, pc);
        grid.sync();
        pc.stamp(ctl, 3);
        requery_phase(a, sm.rq[threadIdx.x >> 5]);
        if (a.profile) {
            grid.sync();
            pc.stamp(ctl, 4);
        }
        verify_watch_phase(a);
        grid.sync();
        pc.stamp(ctl, 5);
    }
    uint32_t round = 0;
    // (uniform: every thread reads the same word behind a grid barrier)
    while (__ldcg(&ctl->repair_needed) != 0 && round < a.max_rounds) {
        bool repaired = false;
        if (a.tiny) {
            // the usual case: every component the new pairs touch is small; one warp re-executes each
            tiny_link_phase(a);
            grid.sync();
            tiny_walk_phase(a, sm.walk[threadIdx.x >> 5]);
            grid.sync();
            pc.stamp(ctl, 6);
            const bool fits = __ldcg(&ctl->tiny_fallback) == 0;
            if (fits) tiny_exec_phase(a, sm.exec);
            grid.sync();
            if (gtid == 0) tiny_done(ctl, a.pair_cap);
            repaired = fits;
            grid.sync();
            pc.stamp(ctl, 7);
        }
        if (!repaired) {
            // a component beyond 64 bodies / 192 pairs (a pile): component-local cooperative repair + list sweep
            repair_local_phase(a, sm.wbuf, grid);
            grid.sync();
            pc.stamp(ctl, 9);
            sweep_phase<true, false>(a, sm.fs, grid, pc);
            grid.sync();
            if (gtid == 0) ctl->repair_needed = 0;
            grid.sync();
            pc.stamp(ctl, 8);
        }
        requery_phase(a, sm.rq[threadIdx.x >> 5]);
        verify_watch_phase(a);
        grid.sync();
        pc.stamp(ctl, 10);
        ++round;
    }
    if (a.taint_bits) {
        taint_phase(a, grid, united);
        grid.sync();
        pc.stamp(ctl, 11);
        taint_cleanup(a);
    }
    if (gtid == 0) finish_step(ctl, a.ctl_snap, a.latch);
}

}  // namespace gp

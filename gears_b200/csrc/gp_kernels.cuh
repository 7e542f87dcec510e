// gp_kernels.cuh — the physics step as sm_100a CUDA kernels.
//
// Reference behaviour reproduced (paths relative to /root/reference):
//   K1  k_keys            RigidBody::update_pos (in registers)   crates/gears-ecs/src/components/physics.rs:405-462
//                         + world AABB centre -> cell key + cell histogram; slabs: migration / halo packing
//   K2  (scan.cuh)        (new) in-place scan of the histogram = cell start table (one-pass bucket sort by cell)
//   K3  k_scatter_cells   update_pos again on the fly + world AABB (get_rotated_aabb, physics.rs:84-135, via per-body
//                         min/max offsets); bodies are physically re-sorted into cell order
//   K4  k_find_pairs      CollisionBox::intersects                physics.rs:148-165  (margin-inflated superset, TMA-staged
//                         tiles, flattened tests, per-warp staged bursts; exact snapshot test for S_snapshot)
//   K5  k_fill_chains / k_sort_chains / k_sort_heavy   (new) per-body contact chains in the reference's pair order
//   K6  k_sweep           check_and_resolve_collision             physics.rs:474-499 -> intersects + resolve (physics.rs:176-299)
//                         in the order of update_systems.rs:408-487
//       k_requery / k_verify_watch / k_restore / k_recount        (new) verification + repair of the speculative pair set
//       k_taint_closure   (new) multi-GPU: which owned bodies are provably independent of unseen bodies
//
// Ordering argument: the reference visits pairs (i,j), i<j in iteration order,
// lexicographically.  Only pairs that share a DYNAMIC body interact (static bodies are never written,
// physics.rs:245-270).  For a body x the pairs involving x, in reference order, are exactly its
// partners sorted by partner rank.  So executing a pair when it is at the head of both of its
// dynamic endpoints' chains (Kahn peeling, one round per dependency level) reproduces the sequential
// sweep bit for bit; pairs within a round touch disjoint dynamic bodies.
//
// All arithmetic: IEEE binary32, compiled with --fmad=false (Rust never contracts), IEEE div/sqrt.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "scan.cuh"

namespace cg = cooperative_groups;

namespace gp {

// ---- flags kept in A.w ----
constexpr uint32_t F_STATIC = 1u;
constexpr uint32_t F_BIG = 2u;     // extent above the grid cell: brute-force list
constexpr uint32_t F_GHOST = 4u;   // multi-GPU halo copy: resolved here but owned elsewhere
constexpr uint32_t F_INTEGRATED = 8u;  // multi-GPU: record received this step, already integrated by its sender

constexpr uint32_t TOPBIT = 0x80000000u;
constexpr uint32_t NONE = 0xffffffffu;

// Pair record, 32 B = one sector: pairs[2*pid] = (a | static<<31, b | static<<31, rank_a | snapshot<<31,
// rank_b | hit<<31); pairs[2*pid+1] = (next pair in a's chain, next pair in b's chain, pending, unused).
// The sweep follows the next links, so it never touches the per-body chain arrays.
__device__ __forceinline__ uint32_t* pair_link(uint4* pairs, uint32_t pid) {
    return reinterpret_cast<uint32_t*>(pairs + 2 * (size_t)pid + 1);
}
__device__ __forceinline__ uint4 link_init() { return make_uint4(NONE, NONE, 0u, 0u); }

// physics.rs:7-20
constexpr float POSITION_CORRECTION_FACTOR = 0.4f;
constexpr float POSITION_CORRECTION_SLOP = 0.01f;
constexpr float FRICTION_COEFFICIENT = 0.8f;
constexpr float VELOCITY_THRESHOLD = 0.05f;
constexpr float GRAVITY_ACCELERATION = 28.0f;
constexpr float FALL_MULTIPLIER = 1.75f;
constexpr float APEX_MULTIPLIER = 0.6f;

struct GridParams {
    float ox, oy, oz;   // origin of cell (1,1,1) (cell 0 and dim-1 are an always-empty apron)
    float inv_cell;
    int dx, dy, dz;     // dims including the apron
    uint32_t ncells;    // dx*dy*dz; key ncells = big list
};

// Device-resident control block: everything the kernels of one step communicate through, so that a
// step needs no host round trip.
struct Ctl {
    // device-resident body counts: kernels never take n from the host, so a step needs no host round trip even
    // when bodies arrive / leave between steps (multi-GPU slabs)
    uint32_t n_in;            // bodies entering the step = live bodies of the last step
    uint32_t n_recv;          // + records received from the slab neighbours this step (appended behind n_in)
    // multi-GPU slab along x: this rank owns the bodies whose AABB centre x lies in [x_lo, x_hi)
    float x_lo, x_hi;
    float halo;               // coverage H: the neighbour sends every body whose inflated AABB comes within H of the face
    uint32_t has_left, has_right;
    uint32_t xcap[2];         // records per exchange buffer (left / right face; both sides of a face agree)
    uint32_t cap;             // body capacity of the arrays
    uint32_t n_sent[2];       // records packed for the left / right neighbour this step
    uint32_t n_migrated_out, n_migrated_in;  // stats of the step
    uint32_t n_unverified;    // owned bodies whose result could depend on a body this rank does not know (this step)
    uint32_t unverified_total;  // summed over the async steps since the last sync
    uint32_t taint_flag[3];
    uint32_t n;               // live bodies of the current step (set by K3 after the sort dropped dead keys)
    float margin;             // relative margin mu of the current step: body x's AABB is inflated by
                              // mu * (largest extent of x) in K4, and K6 verifies it never moves further
    float margin_min, margin_max;
    float margin_in;          // relative INNER margin: pairs within it enter the sweep at once; pairs between the
                              // inner and the outer margin only go on the watch list (lazy candidates)
    uint32_t lazy;            // 0: margin_in == margin always (no watch list; GP_FLAG_NO_RETRY)
    uint32_t repair_tol;      // margin policy: repair sweeps per step that are tolerated before the inner margin grows
    uint32_t n_watch;         // watch-list entries written by K4
    uint32_t n_activated;     // watch pairs that had to join the sweep (stats)
    uint32_t n_requeried;     // pairs beyond the outer margin found by k_requery (stats; grows the outer margin)
    uint32_t n_pairs;         // sweep pairs (K4 inner pairs + activated + requeried; may exceed capacity: overflow)
    uint32_t n_snapshot;
    uint32_t n_contacts;
    uint32_t n_rounds;
    uint32_t max_disp_bits;   // max over bodies of (per-axis AABB displacement / body extent) during the sweep
    uint32_t frontier_n[3];   // rotating: round r consumes [r%3], fills [(r+1)%3], clears [(r+2)%3]
    uint32_t n_heavy;
    uint32_t step_flags;      // SF_* of the current step
    uint32_t n_small;         // bodies in the grid (sorted prefix); n - n_small = big list
    uint32_t sticky_flags;    // OR of step_flags over async steps
    uint32_t steps_inexact;
    uint32_t steps_total;
    // speculative broad phase: bodies that left their margin during the sweep, and the repair loop
    uint32_t n_escapees;
    uint32_t max_abs_disp_bits;  // largest absolute per-axis displacement of any escapee this sweep
    uint32_t repair_needed;      // set by k_requery when it appended pairs; gates the repair kernels
    uint32_t n_repairs;          // repair sweeps executed this step
    uint32_t n_pairs_swept;      // pairs covered by the last sweep (k_requery appends after them)
    uint32_t n_pairs_k4;         // pairs emitted by K4 (the per-body adjacency of K5 covers exactly these)
    // component-local repair (k_repair_local): bodies / pairs whose result the new pairs can influence
    uint32_t n_dirty, n_dpairs, dirty_bump, n_dheavy, dirty_all, n_hits_cleared;
    uint32_t dirty_level[3];     // bodies appended to the dirty list per BFS level (rotating)
    uint32_t n_resweep;          // stats: pairs re-swept by repair rounds this step
    uint32_t n_escapees_total;   // stats: escapees of the first sweep
    // uniform grid of the current step; re-derived by k_finish_step whenever the margin changes
    GridParams grid;
    float lo[3], hi[3];       // bounds of grid-class body centres (from upload / config)
    float ext_max;            // upper bound of every grid-class body extent
    float cell_cfg;           // user cell size (0 = auto)
    uint32_t max_cells;       // capacity of the cell table
};

// The 27-neighbourhood search is complete iff cell >= (largest grid-class extent) * (1 + 2*margin); the slack
// covers the rounding of centre and cell index.  Cells grow until the table fits.
__host__ __device__ inline GridParams make_grid(const float* lo, const float* hi, float ext_max, float margin,
                                                float cell_cfg, uint32_t max_cells) {
    float maxabs = 0.f;
    for (int k = 0; k < 3; ++k) maxabs = fmaxf(maxabs, fmaxf(fabsf(lo[k]), fabsf(hi[k])));
    const float need = (ext_max * (1.0f + 2.0f * margin)) * 1.002f + 8.0f * maxabs * 1.2e-7f;
    float cell = cell_cfg > need ? cell_cfg : need;
    GridParams g;
    for (;;) {
        double ex[3];
        for (int k = 0; k < 3; ++k) ex[k] = floor(((double)hi[k] - (double)lo[k]) / (double)cell) + 7.0;
        if (ex[0] * ex[1] * ex[2] <= (double)max_cells && ex[0] < 2.0e9 && ex[1] < 2.0e9 && ex[2] < 2.0e9) {
            g.dx = (int)ex[0], g.dy = (int)ex[1], g.dz = (int)ex[2];
            break;
        }
        cell *= 1.25f;
    }
    g.inv_cell = 1.0f / cell;
    g.ox = lo[0] - 2.0f * cell, g.oy = lo[1] - 2.0f * cell, g.oz = lo[2] - 2.0f * cell;
    g.ncells = (uint32_t)g.dx * (uint32_t)g.dy * (uint32_t)g.dz;
    return g;
}
constexpr uint32_t SF_PAIR_OVERFLOW = 1u;
constexpr uint32_t SF_MARGIN = 2u;       // the candidate set was still growing when the repair rounds ran out
constexpr uint32_t SF_MARGIN_MAX = 4u;   // (unused)
constexpr uint32_t SF_HEAVY_OVERFLOW = 16u;
constexpr uint32_t SF_ESCAPEE_OVERFLOW = 32u;
constexpr uint32_t SF_EXCHANGE_OVERFLOW = 128u;  // a slab exchange buffer or the body arrays were too small
constexpr uint32_t SF_INTERNAL = 64u;     // a bounded wait expired (TMA copy never landed): result invalid

// Body record (array of structures, 64 B = one DRAM burst): P = pos|mass, V = vel|restitution, Lo = min offset|id,
// Hi = max offset|rank.  The contact sweep touches bodies at random; with the four float4 in one 64 B line a body
// costs one burst instead of four.  Streaming kernels read the 4 x 16 B of a record with consecutive LDG.128.
struct Bodies {
    float4* b;
    __device__ __forceinline__ float4& P(uint32_t i) const { return b[4 * (size_t)i]; }
    __device__ __forceinline__ float4& V(uint32_t i) const { return b[4 * (size_t)i + 1]; }
    __device__ __forceinline__ float4& Lo(uint32_t i) const { return b[4 * (size_t)i + 2]; }
    __device__ __forceinline__ float4& Hi(uint32_t i) const { return b[4 * (size_t)i + 3]; }
};

struct V3 {
    float x, y, z;
};
__device__ __forceinline__ V3 mk(float x, float y, float z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 operator/(V3 a, float s) { return mk(a.x / s, a.y / s, a.z / s); }
// cgmath 0.18 dot: (x*x' + y*y') + z*z'   — no FMA (--fmad=false)
__device__ __forceinline__ float dot3(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ float len3(V3 a) { return sqrtf(dot3(a, a)); }
__device__ __forceinline__ V3 cross3(V3 a, V3 b) {
    return mk((a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x));
}
__device__ __forceinline__ V3 xyz(float4 v) { return mk(v.x, v.y, v.z); }

__device__ __forceinline__ float box_extent(const float4 lo, const float4 hi) {
    return fmaxf(hi.x - lo.x, fmaxf(hi.y - lo.y, hi.z - lo.z));
}

__device__ __forceinline__ uint32_t cell_key(const GridParams& g, float cx, float cy, float cz) {
    // __float2int_rd saturates and maps NaN to 0; the clamp keeps every body inside [1, dim-2]
    int ix = __float2int_rd((cx - g.ox) * g.inv_cell) + 1;
    int iy = __float2int_rd((cy - g.oy) * g.inv_cell) + 1;
    int iz = __float2int_rd((cz - g.oz) * g.inv_cell) + 1;
    ix = min(max(ix, 1), g.dx - 2);
    iy = min(max(iy, 1), g.dy - 2);
    iz = min(max(iz, 1), g.dz - 2);
    return (uint32_t)ix + (uint32_t)g.dx * ((uint32_t)iy + (uint32_t)g.dy * (uint32_t)iz);
}

// ---------------------------------------------------------------------------------------------
// Upload: derive per-body AABB offsets.  off_min/off_max = component-wise min/max over the 8
// corners of rot*(corner*scale) (physics.rs:93-119 without the translation).  Because rounding is
// monotone, min_i fl(r_i + p) == fl(min_i r_i + p): adding pos afterwards gives the reference's AABB.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void corner_offsets(V3 bmin, V3 bmax, float4 q /*v.xyz, s*/, V3 sc, V3& omin, V3& omax) {
    const V3 qv = mk(q.x, q.y, q.z);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        V3 c = mk((i & 1) ? bmax.x : bmin.x, (i & 2) ? bmax.y : bmin.y, (i & 4) ? bmax.z : bmin.z);
        V3 s = mk(c.x * sc.x, c.y * sc.y, c.z * sc.z);
        V3 t = cross3(qv, s) + (s * q.w);
        V3 r = (cross3(qv, t) * 2.0f) + s;
        if (i == 0) {
            omin = r;
            omax = r;
        } else {
            omin = mk(fminf(omin.x, r.x), fminf(omin.y, r.y), fminf(omin.z, r.z));
            omax = mk(fmaxf(omax.x, r.x), fmaxf(omax.y, r.y), fmaxf(omax.z, r.z));
        }
    }
}

// Packs host-layout staging arrays into the device SoA.  idx == nullptr: body i = thread i.
__global__ void k_pack(uint32_t k, const uint32_t* __restrict__ idx, const float* __restrict__ pos3,
                       const float* __restrict__ rot4, const float* __restrict__ scale3, const float* __restrict__ vel3,
                       const float* __restrict__ acc3, const float* __restrict__ bmin3, const float* __restrict__ bmax3,
                       const float* __restrict__ mass, const float* __restrict__ rest, const uint8_t* __restrict__ stat,
                       const uint32_t* __restrict__ entity, const uint32_t* __restrict__ rank, Bodies B,
                       float4* A, float* ext, int shape_only,
                       const uint32_t* __restrict__ slot_of, float4* __restrict__ rawbox /* by body id */) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= k) return;
    const uint32_t i = idx ? slot_of[idx[t]] : t;  // bodies are re-sorted every step: caller index -> current slot
    const V3 bmin = mk(bmin3[3 * t], bmin3[3 * t + 1], bmin3[3 * t + 2]);
    const V3 bmax = mk(bmax3[3 * t], bmax3[3 * t + 1], bmax3[3 * t + 2]);
    const float4 q = rot4 ? make_float4(rot4[4 * t], rot4[4 * t + 1], rot4[4 * t + 2], rot4[4 * t + 3])
                          : make_float4(0.f, 0.f, 0.f, 1.f);
    const V3 sc = scale3 ? mk(scale3[3 * t], scale3[3 * t + 1], scale3[3 * t + 2]) : mk(1.f, 1.f, 1.f);
    V3 omin, omax;
    corner_offsets(bmin, bmax, q, sc, omin, omax);
    const uint32_t fl_static = stat[t] ? F_STATIC : 0u;
    if (shape_only) {
        float4 p = B.P(i), v = B.V(i), a = A[i], lo = B.Lo(i), hi = B.Hi(i);
        p.w = mass[t];
        v.w = rest[t];
        a.w = __uint_as_float((__float_as_uint(a.w) & ~F_STATIC) | fl_static);
        lo = make_float4(omin.x, omin.y, omin.z, lo.w);
        hi = make_float4(omax.x, omax.y, omax.z, hi.w);
        B.P(i) = p, B.V(i) = v, A[i] = a, B.Lo(i) = lo, B.Hi(i) = hi;
    } else {
        B.P(i) = make_float4(pos3[3 * t], pos3[3 * t + 1], pos3[3 * t + 2], mass[t]);
        B.V(i) = make_float4(vel3[3 * t], vel3[3 * t + 1], vel3[3 * t + 2], rest[t]);
        A[i] = make_float4(acc3 ? acc3[3 * t] : 0.f, acc3 ? acc3[3 * t + 1] : 0.f, acc3 ? acc3[3 * t + 2] : 0.f,
                           __uint_as_float(fl_static));
        B.Lo(i) = make_float4(omin.x, omin.y, omin.z, __uint_as_float(i));  // .w = body id (caller index)
        (void)entity;
        B.Hi(i) = make_float4(omax.x, omax.y, omax.z, __uint_as_float(rank ? rank[t] : i));
    }
    if (ext) ext[i] = fmaxf(omax.x - omin.x, fmaxf(omax.y - omin.y, omax.z - omin.z));
    // the un-rotated, un-scaled collision box: what the ray cast and the obstacle grid of the reference use
    const uint32_t id = idx ? idx[t] : t;
    rawbox[2 * (size_t)id] = make_float4(bmin.x, bmin.y, bmin.z, 0.f);
    rawbox[2 * (size_t)id + 1] = make_float4(bmax.x, bmax.y, bmax.z, 0.f);
}

// histogram of body extents in quarter-octave buckets (bucket b: ext <= 2^((b-64)/4)), plus bounds of centres
constexpr int EXT_BUCKETS = 160;
__device__ __forceinline__ int ext_bucket(float e) {
    if (!(e > 0.f)) return 0;
    if (isinf(e) || isnan(e)) return EXT_BUCKETS - 1;
    int b = (int)ceilf(log2f(e) * 4.0f) + 64;
    // make the bucket edge an upper bound despite log2f rounding
    if (b >= 0 && b < EXT_BUCKETS && exp2f((float)(b - 64) * 0.25f) < e) ++b;
    return min(max(b, 0), EXT_BUCKETS - 1);
}
__host__ __device__ inline float ext_bucket_edge(int b) { return exp2f((float)(b - 64) * 0.25f); }

// hist[0 .. EXT_BUCKETS) counts every body, hist[EXT_BUCKETS .. 2 EXT_BUCKETS) the dynamic ones only
__global__ void k_ext_hist(uint32_t n, const float* __restrict__ ext, const float4* __restrict__ A,
                           uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[2 * EXT_BUCKETS];
    for (int i = threadIdx.x; i < 2 * EXT_BUCKETS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int b = ext_bucket(ext[i]);
        atomicAdd(&sh[b], 1u);
        if (!(__float_as_uint(A[i].w) & F_STATIC)) atomicAdd(&sh[EXT_BUCKETS + b], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * EXT_BUCKETS; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

// order-preserving float <-> uint for atomicMin/Max
__device__ __forceinline__ uint32_t f2ord(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & TOPBIT) ? ~u : (u | TOPBIT);
}
__host__ __device__ inline float ord2f(uint32_t u) {
    uint32_t v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(v);
#else
    float f;
    memcpy(&f, &v, 4);
    return f;
#endif
}

// classify by extent threshold and reduce the bounds of grid-class body centres
__global__ void k_classify(uint32_t n, float thr, const Bodies B, float4* A,
                           uint32_t* __restrict__ bounds /*6: min xyz, max xyz (ordered uints)*/) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    if (i < n) {
        float4 a = A[i];
        uint32_t f = __float_as_uint(a.w) & ~F_BIG;
        const float4 l4 = B.Lo(i), h4 = B.Hi(i);
        const bool big = !(fmaxf(h4.x - l4.x, fmaxf(h4.y - l4.y, h4.z - l4.z)) <= thr);
        if (big) f |= F_BIG;
        a.w = __uint_as_float(f);
        A[i] = a;
        if (!big) {
            const float4 p = B.P(i), l = l4, h = h4;
            const float c[3] = {((l.x + p.x) + (h.x + p.x)) * 0.5f, ((l.y + p.y) + (h.y + p.y)) * 0.5f,
                                ((l.z + p.z) + (h.z + p.z)) * 0.5f};
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (isfinite(c[k])) lo[k] = hi[k] = c[k];
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (lo[k] < 3.0e38f) atomicMin(&bounds[k], f2ord(lo[k]));
            if (hi[k] > -3.0e38f) atomicMax(&bounds[3 + k], f2ord(hi[k]));
        }
    }
}

// The body arrays are physically re-sorted into cell order by every step (K3), so the slot of a body changes.
// Lo[slot].w keeps the body id (the caller's index at upload; the global entity id in multi-GPU mode).
__global__ void k_slot_map(uint32_t n, const Bodies B, uint32_t* __restrict__ slot_of) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) slot_of[__float_as_uint(B.Lo(i).w)] = i;
}

// bulk refresh of mutable state from host-layout arrays (gp_set_state: idx == nullptr, every slot gathers the
// row of its body id;  gp_update_dirty: row t belongs to body idx[t])
__global__ void k_set_state(uint32_t k, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ slot_of,
                            const float* __restrict__ pos3, const float* __restrict__ vel3,
                            const float* __restrict__ acc3, Bodies B, float4* A) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= k) return;
    uint32_t i, r;
    if (idx) {
        i = slot_of[idx[t]], r = t;
    } else {
        i = t, r = __float_as_uint(B.Lo(t).w);
    }
    if (pos3) {
        float4 p = B.P(i);
        p.x = pos3[3 * (size_t)r], p.y = pos3[3 * (size_t)r + 1], p.z = pos3[3 * (size_t)r + 2];
        B.P(i) = p;
    }
    if (vel3) {
        float4 v = B.V(i);
        v.x = vel3[3 * (size_t)r], v.y = vel3[3 * (size_t)r + 1], v.z = vel3[3 * (size_t)r + 2];
        B.V(i) = v;
    }
    if (acc3) {
        float4 a = A[i];
        a.x = acc3[3 * (size_t)r], a.y = acc3[3 * (size_t)r + 1], a.z = acc3[3 * (size_t)r + 2];
        A[i] = a;
    }
}

// device -> host layout.  by_id: row = body id (gp_download); else row = slot (gp_download_owned, + ids)
__global__ void k_unpack(uint32_t n, const Bodies B, float* pos3, float* vel3, uint32_t* ids, int by_id) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t id = __float_as_uint(B.Lo(i).w);
    const size_t r = by_id ? id : i;
    if (pos3) {
        const float4 p = B.P(i);
        pos3[3 * r] = p.x, pos3[3 * r + 1] = p.y, pos3[3 * r + 2] = p.z;
    }
    if (vel3) {
        const float4 v = B.V(i);
        vel3[3 * r] = v.x, vel3[3 * r + 1] = v.y, vel3[3 * r + 2] = v.z;
    }
    if (ids) ids[i] = id;
}

// RigidBody::update_pos, physics.rs:405-462 (a.w carries the flags; static bodies are untouched)
__device__ __forceinline__ void integrate_body(float4& p, float4& v, const float4 a, float dt, float d16, float d18,
                                               float d35) {
    if (__float_as_uint(a.w) & F_STATIC) return;
    const bool falling = v.y < 0.0f;
    const bool near_apex = fabsf(v.y) < 2.0f && v.y > 0.0f;
    const float gm = falling ? FALL_MULTIPLIER : (near_apex ? APEX_MULTIPLIER : 1.0f);
    v.y = v.y - ((GRAVITY_ACCELERATION * gm) * dt);
    const bool accelerating = len3(mk(a.x, a.y, a.z)) > 0.01f;
    const float d = accelerating ? d16 : (falling ? d18 : d35);
    if (falling) {
        v.x = v.x * d;
        v.z = v.z * d;
    } else {
        v.x = v.x * d;
        v.y = v.y * d;
        v.z = v.z * d;
    }
    v.x = v.x + a.x * dt;
    v.z = v.z + a.z * dt;
    if (len3(mk(v.x, v.y, v.z)) < 0.005f) {
        v.x = 0.0f, v.y = 0.0f, v.z = 0.0f;
    }
    p.x = p.x + v.x * dt;
    p.y = p.y + v.y * dt;
    p.z = p.z + v.z * dt;
}

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
// wait until *f >= want for both flags (nullptr = no neighbour on that side)
__global__ void k_wait_flags(const uint32_t* f0, const uint32_t* f1, uint32_t want, Ctl* ctl) {
    const uint32_t* f = threadIdx.x == 0 ? f0 : f1;
    if (threadIdx.x > 1 || !f) return;
    for (uint32_t spin = 0; spin < (1u << 23); ++spin) {
        if (ld_acquire_sys(f) >= want) return;
        __nanosleep(100);
    }
    atomicOr(&ctl->step_flags, SF_INTERNAL);  // the neighbour never arrived (~1 s): the step is invalid, not hung
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
// world AABBs.  R 8 (key, rank) + R 80 (own record, coalesced) + W 80 (next set) + W 32 (AABB) + W 12.
// The order inside a cell is the arrival order of the atomics: results do not depend on it (the contact
// sweep orders by the reference's iteration rank, never by slot).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    k_scatter_cells(const uint32_t* __restrict__ key, const uint32_t* __restrict__ lrank,
                    const Bodies B0, const float4* __restrict__ A0, const Bodies B1, float4* __restrict__ A1,
                    float4* __restrict__ sbox, uint32_t* __restrict__ chain_cnt, uint32_t* __restrict__ perm,
                    float* __restrict__ dmax, float* __restrict__ reach, uint32_t* __restrict__ moved_bits,
                    uint32_t* __restrict__ esc_bits, const uint32_t* __restrict__ LB, Ctl* ctl,
                    float dt, float d16, float d18, float d35) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t ncells = ctl->grid.ncells;
    if (i == 0) {
        ctl->n_small = LB[ncells];
        ctl->n = LB[ncells + 1u];
    }
    if (i >= ctl->n_in + ctl->n_recv) return;
    const uint32_t k = key[i];
    if (k > ncells) return;  // dead (a ghost of the last step, a body that migrated away): dropped here
    const uint32_t s = LB[k] + lrank[i];
    float4 p = B0.P(i), v = B0.V(i), a = A0[i];
    const float4 lo = B0.Lo(i), hi = B0.Hi(i);
    uint32_t flags = __float_as_uint(a.w);
    if (flags & F_INTEGRATED) {
        flags &= ~F_INTEGRATED;
        a.w = __uint_as_float(flags);
    } else {
        integrate_body(p, v, a, dt, d16, d18, d35);
    }
    B1.P(s) = p, B1.V(s) = v, A1[s] = a, B1.Lo(s) = lo, B1.Hi(s) = hi;
    // world AABB = offsets + position (physics.rs:117: corner + pos; min/max commute with the rounding)
    // AABB record: lo.w = the body's outer margin fl(mu * extent), sign bit = static; hi.w = rank
    const float4 wlo = make_float4(lo.x + p.x, lo.y + p.y, lo.z + p.z, 0.0f);
    const float4 whi = make_float4(hi.x + p.x, hi.y + p.y, hi.z + p.z, hi.w);
    const float m = ctl->margin * box_extent(wlo, whi);
    sbox[2 * (size_t)s] = make_float4(wlo.x, wlo.y, wlo.z,
                                      __uint_as_float((__float_as_uint(m) & ~TOPBIT) | ((flags & F_STATIC) ? TOPBIT : 0u)));
    sbox[2 * (size_t)s + 1] = whi;
    chain_cnt[s] = 0;
    perm[s] = i;  // slot in the input set: k_restore re-integrates from there
    dmax[s] = 0.0f, reach[s] = 0.0f;
    if ((s & 31u) == 0u) moved_bits[s >> 5] = 0u, esc_bits[s >> 5] = 0u;  // every word of [0, n) has one such slot
}

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

// ---------------------------------------------------------------------------------------------
// K5: contact chains.  offs = exclusive scan of chain_cnt (host launches scan.cuh).  k_fill_chains
// drops, for every pair and dynamic endpoint x, the entry (partner rank << 32 | side << 31 | pair id) into x's
// segment; k_sort_chains orders each segment by partner rank = the reference's visiting order for x,
// links every pair to its successor on x's side (pair record: next_a / next_b) and counts, per pair, on how
// many chains it is not yet at the head (pending).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    k_fill_chains(const uint4* __restrict__ pairs, uint32_t pair_cap, const uint32_t* __restrict__ offs,
                  uint32_t* __restrict__ chain_cnt, unsigned long long* __restrict__ adj, const Ctl* ctl,
                  const uint32_t* __restrict__ gate) {
    if (gate && *gate == 0) return;
    const uint32_t np = min(ctl->n_pairs, pair_cap);
    for (uint32_t pid = blockIdx.x * blockDim.x + threadIdx.x; pid < np; pid += gridDim.x * blockDim.x) {
        const uint4 r = pairs[2 * (size_t)pid];
        const uint32_t ra = r.z & ~TOPBIT, rb = r.w & ~TOPBIT;
        if (!(r.x & TOPBIT)) {
            const uint32_t a = r.x;
            const uint32_t s = offs[a] + atomicSub(&chain_cnt[a], 1u) - 1u;
            adj[s] = ((unsigned long long)rb << 32) | pid;  // side 0: this body is the pair's a
        }
        if (!(r.y & TOPBIT)) {
            const uint32_t b = r.y;
            const uint32_t s = offs[b] + atomicSub(&chain_cnt[b], 1u) - 1u;
            adj[s] = ((unsigned long long)ra << 32) | TOPBIT | pid;  // side 1
        }
    }
}

constexpr uint32_t CHAIN_INLINE = 48;  // longer chains go to the block-wide sorter

// entries [first, d) of a sorted chain, stride apart: entry i-1 gets entry i as its successor on this body's side
__device__ __forceinline__ void link_chain(const unsigned long long* __restrict__ seg, uint32_t first, uint32_t d,
                                           uint32_t stride, uint4* __restrict__ pairs) {
    for (uint32_t i = first + 1; i < d; i += stride) {
        const uint32_t prev = (uint32_t)seg[i - 1], cur = (uint32_t)seg[i] & ~TOPBIT;
        pair_link(pairs, prev & ~TOPBIT)[prev >> 31] = cur;
        atomicAdd(&pair_link(pairs, cur)[2], 1u);
    }
}

__global__ void __launch_bounds__(256)
    k_sort_chains(const uint32_t* __restrict__ offs, unsigned long long* __restrict__ adj,
                  uint4* __restrict__ pairs, uint32_t* __restrict__ heavy, uint32_t heavy_cap, Ctl* ctl,
                  const uint32_t* __restrict__ gate) {
    if (gate && *gate == 0) return;
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= ctl->n) return;
    const uint32_t b = offs[x], e = offs[x + 1];
    const uint32_t d = e - b;
    if (d == 0) return;
    if (d > CHAIN_INLINE) {
        const uint32_t s = atomicAdd(&ctl->n_heavy, 1u);
        if (s < heavy_cap)
            heavy[s] = x;
        else
            atomicOr(&ctl->step_flags, SF_HEAVY_OVERFLOW);
        return;
    }
    // insertion sort in place (segments are short and L1-resident)
    for (uint32_t i = b + 1; i < e; ++i) {
        const unsigned long long v = adj[i];
        uint32_t j = i;
        while (j > b && adj[j - 1] > v) {
            adj[j] = adj[j - 1];
            --j;
        }
        adj[j] = v;
    }
    link_chain(adj + b, 0u, d, 1u, pairs);
}

// A chain longer than CHAIN_INLINE is sorted by ONE WARP: bitonic network in its all-ascending form (the first
// sub-step of each stage mirrors the partner: l = i ^ (k-1), the rest use l = i ^ j), so every compare-exchange moves
// the smaller key to the lower index and the virtual +inf padding at [d, m) never has to move.  Up to WSORT_SMEM
// entries the chain is staged in the warp's shared-memory buffer; longer chains are sorted in place in global memory
// (through L2: the lanes exchange data there).  Then the warp links the chain (successor pointers + pending counts).
constexpr uint32_t WSORT_SMEM = 128;
__device__ __forceinline__ void warp_sort_and_link(unsigned long long* seg, uint32_t d, unsigned long long* wbuf,
                                                   uint4* __restrict__ pairs) {
    const uint32_t lane = threadIdx.x & 31;
    uint32_t m = 1;
    while (m < d) m <<= 1;
    if (d <= WSORT_SMEM) {
        for (uint32_t i = lane; i < d; i += 32) wbuf[i] = __ldcg(&seg[i]);
        __syncwarp();
        for (uint32_t k = 2; k <= m; k <<= 1)
            for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                const uint32_t mask = (j == (k >> 1)) ? (k - 1u) : j;
                for (uint32_t i = lane; i < m; i += 32) {
                    const uint32_t l = i ^ mask;
                    if (l > i && l < d) {
                        const unsigned long long vi = wbuf[i], vl = wbuf[l];
                        if (vi > vl) wbuf[i] = vl, wbuf[l] = vi;
                    }
                }
                __syncwarp();
            }
        for (uint32_t i = lane; i < d; i += 32) __stcg(&seg[i], wbuf[i]);
        for (uint32_t i = lane + 1; i < d; i += 32) {
            const uint32_t prev = (uint32_t)wbuf[i - 1], cur = (uint32_t)wbuf[i] & ~TOPBIT;
            pair_link(pairs, prev & ~TOPBIT)[prev >> 31] = cur;
            atomicAdd(&pair_link(pairs, cur)[2], 1u);
        }
        __syncwarp();
        return;
    }
    for (uint32_t k = 2; k <= m; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            const uint32_t mask = (j == (k >> 1)) ? (k - 1u) : j;
            for (uint32_t i = lane; i < m; i += 32) {
                const uint32_t l = i ^ mask;
                if (l > i && l < d) {
                    const unsigned long long vi = __ldcg(&seg[i]), vl = __ldcg(&seg[l]);
                    if (vi > vl) __stcg(&seg[i], vl), __stcg(&seg[l], vi);
                }
            }
            __syncwarp();
        }
    for (uint32_t i = lane + 1; i < d; i += 32) {
        const uint32_t prev = (uint32_t)__ldcg(&seg[i - 1]), cur = (uint32_t)__ldcg(&seg[i]) & ~TOPBIT;
        pair_link(pairs, prev & ~TOPBIT)[prev >> 31] = cur;
        atomicAdd(&pair_link(pairs, cur)[2], 1u);
    }
    __syncwarp();
}

// one warp per heavy body (a dense pile makes EVERY chain heavy: a million of them)
__global__ void __launch_bounds__(256)
    k_sort_heavy(const uint32_t* __restrict__ offs, unsigned long long* __restrict__ adj, uint4* __restrict__ pairs,
                 const uint32_t* __restrict__ heavy, uint32_t heavy_cap, const Ctl* ctl,
                 const uint32_t* __restrict__ gate) {
    __shared__ unsigned long long wbuf[8][WSORT_SMEM];
    if (gate && *gate == 0) return;
    const uint32_t nh = min(ctl->n_heavy, heavy_cap);
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t hi = warp; hi < nh; hi += nwarps) {
        const uint32_t x = heavy[hi];
        const uint32_t b = offs[x], d = offs[x + 1] - b;
        warp_sort_and_link(adj + b, d, wbuf[threadIdx.x >> 5], pairs);
    }
}

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
            uint32_t esc_cap, const uint32_t* __restrict__ gate, const uint32_t* __restrict__ list) {
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
                BodyRW a, b;
                a.is_static = (rec.x & TOPBIT) != 0;
                b.is_static = (rec.y & TOPBIT) != 0;
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
                    }
                    pairs[2 * (size_t)pid].w = rec.w | TOPBIT;  // S_sweep flag; k_restore: these bodies were written
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

// ONE cooperative launch: clear, seed, propagate to the fixed point (a grid barrier per round), count.
constexpr uint32_t TAINT_MAX_ROUNDS = 64;
__global__ void __launch_bounds__(COOP_THREADS, 1)
    k_taint_closure(const float4* __restrict__ sbox, const float* __restrict__ dmax, const float* __restrict__ reach,
                    const uint32_t* __restrict__ moved_bits, uint32_t* __restrict__ taint_bits,
                    const uint4* __restrict__ pairs, uint32_t pair_cap, const uint2* __restrict__ watch,
                    const float4* __restrict__ A, Ctl* ctl) {
    cg::grid_group grid = cg::this_grid();
    const uint32_t n = ctl->n;
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    for (uint32_t wd = gtid; wd < (n + 31u) / 32u; wd += gsz) taint_bits[wd] = 0u;
    if (gtid == 0) ctl->taint_flag[0] = 0u, ctl->taint_flag[1] = 0u, ctl->taint_flag[2] = 0u;
    grid.sync();
    // seeds: bodies whose reach pokes beyond what the neighbours mirror
    for (uint32_t i = gtid; i < n; i += gsz) {
        const float4 lo = sbox[2 * (size_t)i], hi = sbox[2 * (size_t)i + 1];
        if (box_static(lo)) continue;  // static bodies never change: nothing to get wrong
        float r = box_margin(lo);
        if ((moved_bits[i >> 5] >> (i & 31u)) & 1u) r = fmaxf(r, fmaxf(reach[i], dmax[i] * REACH_SLACK));
        const bool out_r = ctl->has_right && !(hi.x + r <= ctl->x_hi + ctl->halo);
        const bool out_l = ctl->has_left && !(lo.x - r >= ctl->x_lo - ctl->halo);
        if (out_l || out_r) atomicOr(&taint_bits[i >> 5], 1u << (i & 31u));
    }
    grid.sync();
    const uint32_t np = min(ctl->n_pairs, pair_cap), nw = min(ctl->n_watch, pair_cap);
    uint32_t it = 0;
    bool open = true;
    while (open && it < TAINT_MAX_ROUNDS) {
        // three rotating flags: round `it` raises flag it%3 and clears flag (it+1)%3, which was last READ at the end
        // of round it-2, i.e. behind the barrier of round it-1 (with two flags a fast thread would clear the flag a
        // slow thread is still about to read, and the grid would disagree about leaving the loop)
        uint32_t* flag = &ctl->taint_flag[it % 3u];
        if (gtid == 0) ctl->taint_flag[(it + 1u) % 3u] = 0u;
        for (uint32_t i = gtid; i < np; i += gsz) {
            const uint4 r = pairs[2 * (size_t)i];
            taint_edge(r.x, r.y, taint_bits, flag);
        }
        for (uint32_t i = gtid; i < nw; i += gsz) {
            const uint2 r = watch[i];
            if (r.x != NONE) taint_edge(r.x, r.y, taint_bits, flag);  // (activated entries are among the pairs)
        }
        grid.sync();
        open = __ldcg(flag) != 0u;
        ++it;
    }
    // owned dynamic bodies that are not provably independent of unseen bodies
    uint32_t c = 0;
    for (uint32_t i = gtid; i < n; i += gsz) {
        const uint32_t f = __float_as_uint(A[i].w);
        if (!(f & (F_GHOST | F_STATIC)) && (open || ((__ldcg(&taint_bits[i >> 5]) >> (i & 31u)) & 1u))) ++c;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&ctl->n_unverified, c);
}

// owned bodies (no halo copies; replicated big statics only on the rank that has no left neighbour) -> host layout
__global__ void __launch_bounds__(256)
    k_compact_owned(const Ctl* ctl, const Bodies B, const float4* __restrict__ A, int is_first_rank, uint32_t* counter,
                    uint32_t* __restrict__ ids, float* __restrict__ pos3, float* __restrict__ vel3,
                    const uint32_t* __restrict__ taint_bits, uint8_t* __restrict__ unverified) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ctl->n_in) return;  // between steps n_in = the live bodies of the last step
    const uint32_t f = __float_as_uint(A[i].w);
    if (f & F_GHOST) return;
    if ((f & F_BIG) && !is_first_rank) return;
    cg::coalesced_group g = cg::coalesced_threads();
    uint32_t base = 0;
    if (g.thread_rank() == 0) base = atomicAdd(counter, g.size());
    base = g.shfl(base, 0);
    const uint32_t s = base + g.thread_rank();
    if (!ids) return;  // counting pass
    const float4 p = B.P(i), v = B.V(i);
    ids[s] = __float_as_uint(B.Lo(i).w);
    pos3[3 * (size_t)s] = p.x, pos3[3 * (size_t)s + 1] = p.y, pos3[3 * (size_t)s + 2] = p.z;
    vel3[3 * (size_t)s] = v.x, vel3[3 * (size_t)s + 1] = v.y, vel3[3 * (size_t)s + 2] = v.z;
    if (taint_bits && unverified) unverified[s] = (taint_bits[i >> 5] >> (i & 31u)) & 1u;  // of the LAST step
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

// ---------------------------------------------------------------------------------------------
// "Next" rows: consumers of the device-resident body state right after the step.
// ---------------------------------------------------------------------------------------------
// N1  pass 3 of update_physics (update_systems.rs:489-527): Instance::to_raw (gears-renderer instance.rs:22-31) for
// every body in one launch instead of one queue.write_buffer per entity.  cgmath 0.18 semantics, no FMA:
// model = T(pos) * M4(q) * S(scale), normal = M3(q); column j of A*R = ((a*R[j][0] + b*R[j][1]) + c*R[j][2]) + d*R[j][3].
__device__ __forceinline__ void mat4_mul(const float* A, const float* R, float* O) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float acc = A[0 * 4 + i] * R[j * 4 + 0];
            acc = acc + A[1 * 4 + i] * R[j * 4 + 1];
            acc = acc + A[2 * 4 + i] * R[j * 4 + 2];
            acc = acc + A[3 * 4 + i] * R[j * 4 + 3];
            O[j * 4 + i] = acc;
        }
}
__global__ void __launch_bounds__(128)
    k_instances(uint32_t n, const Bodies B, const float* __restrict__ rot4, const float* __restrict__ scale3,
                float* __restrict__ out25) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t id = __float_as_uint(B.Lo(i).w);
    const float4 p = B.P(i);
    const float x = rot4 ? rot4[4 * (size_t)id] : 0.f, y = rot4 ? rot4[4 * (size_t)id + 1] : 0.f;
    const float z = rot4 ? rot4[4 * (size_t)id + 2] : 0.f, w = rot4 ? rot4[4 * (size_t)id + 3] : 1.f;
    const float sx = scale3 ? scale3[3 * (size_t)id] : 1.f, sy = scale3 ? scale3[3 * (size_t)id + 1] : 1.f;
    const float sz = scale3 ? scale3[3 * (size_t)id + 2] : 1.f;
    const float x2 = x + x, y2 = y + y, z2 = z + z;
    const float xx2 = x2 * x, xy2 = x2 * y, xz2 = x2 * z, yy2 = y2 * y, yz2 = y2 * z, zz2 = z2 * z;
    const float sy2 = y2 * w, sz2 = z2 * w, sx2 = x2 * w;
    const float r3[9] = {1.0f - yy2 - zz2, xy2 + sz2, xz2 - sy2, xy2 - sz2, 1.0f - xx2 - zz2,
                         yz2 + sx2,        xz2 + sy2, yz2 - sx2, 1.0f - xx2 - yy2};
    const float T[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, p.x, p.y, p.z, 1};
    const float R[16] = {r3[0], r3[1], r3[2], 0, r3[3], r3[4], r3[5], 0, r3[6], r3[7], r3[8], 0, 0, 0, 0, 1};
    const float S[16] = {sx, 0, 0, 0, 0, sy, 0, 0, 0, 0, sz, 0, 0, 0, 0, 1};
    float TR[16], M[16];
    mat4_mul(T, R, TR);
    mat4_mul(TR, S, M);
    float* o = out25 + 25 * (size_t)id;
#pragma unroll
    for (int k = 0; k < 16; ++k) o[k] = M[k];
#pragma unroll
    for (int k = 0; k < 9; ++k) o[16 + k] = r3[k];
}

// N3  Weapon::ray_intersects_aabb (gears-ecs interactive.rs:94-158), batched: thread = (ray, target)
__global__ void __launch_bounds__(256)
    k_raycast(uint32_t nrays, const float* __restrict__ origin3, const float* __restrict__ dir3, uint32_t ntargets,
              const uint32_t* __restrict__ targets, const uint32_t* __restrict__ slot_of, const Bodies B,
              const float4* __restrict__ rawbox, uint8_t* __restrict__ hits) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)nrays * ntargets) return;
    const uint32_t r = (uint32_t)(t / ntargets), k = (uint32_t)(t % ntargets);
    const uint32_t id = targets ? targets[k] : k;
    const float4 p = B.P(slot_of[id]);
    const float4 bmin = rawbox[2 * (size_t)id], bmax = rawbox[2 * (size_t)id + 1];
    const float o[3] = {origin3[3 * r], origin3[3 * r + 1], origin3[3 * r + 2]};
    const float d[3] = {dir3[3 * r], dir3[3 * r + 1], dir3[3 * r + 2]};
    const float mn[3] = {p.x + bmin.x, p.y + bmin.y, p.z + bmin.z};
    const float mx[3] = {p.x + bmax.x, p.y + bmax.y, p.z + bmax.z};
    float tmin = 0.0f, tmax = 3.40282347e+38f;
    bool hit = true;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (!hit) break;
        if (fabsf(d[i]) < 1e-8f) {
            if (o[i] < mn[i] || o[i] > mx[i]) hit = false;
        } else {
            const float inv = 1.0f / d[i];
            float t1 = (mn[i] - o[i]) * inv, t2 = (mx[i] - o[i]) * inv;
            if (t1 > t2) {
                const float tt = t1;
                t1 = t2, t2 = tt;
            }
            tmin = fmaxf(tmin, t1);
            tmax = fminf(tmax, t2);
            if (tmin > tmax) hit = false;
        }
    }
    hits[t] = (hit && tmax >= 0.0f && tmin <= tmax) ? 1 : 0;
}

// N4  PathfindingGrid::add_obstacle (gears-ecs pathfinding.rs:296-338, GridPos::from_world_pos :57-63): rasterise the
// collision boxes of k bodies into a dense bitmap; one warp per obstacle, lanes stride over its cells.
__device__ __forceinline__ int f2i_sat(float f) { return __float2int_rz(f); }  // saturating, NaN -> 0 (= Rust `as i32`)
__global__ void __launch_bounds__(256)
    k_obstacle_grid(uint32_t k, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ slot_of, const Bodies B,
                    const float4* __restrict__ rawbox, float cell, int ox, int oy, int oz, uint32_t dx, uint32_t dy,
                    uint32_t dz, uint32_t* __restrict__ bitmap, int* __restrict__ bounds /*6, pre-set to +-INT*/) {
    const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (wid >= k) return;
    const uint32_t id = idx ? idx[wid] : wid;
    const float4 p = B.P(slot_of[id]);
    const float4 bmin = rawbox[2 * (size_t)id], bmax = rawbox[2 * (size_t)id + 1];
    const int lo[3] = {f2i_sat(roundf((p.x + bmin.x) / cell)), f2i_sat(roundf((p.y + bmin.y) / cell)),
                       f2i_sat(roundf((p.z + bmin.z) / cell))};
    const int hi[3] = {f2i_sat(roundf((p.x + bmax.x) / cell)), f2i_sat(roundf((p.y + bmax.y) / cell)),
                       f2i_sat(roundf((p.z + bmax.z) / cell))};
    if (lane == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {  // the reference folds the corner cells into its bounds unconditionally
            atomicMin(&bounds[a], lo[a]);
            atomicMax(&bounds[3 + a], hi[a]);
        }
    }
    // clip to the bitmap window
    const long long x0 = max((long long)lo[0], (long long)ox), x1 = min((long long)hi[0], (long long)ox + dx - 1);
    const long long y0 = max((long long)lo[1], (long long)oy), y1 = min((long long)hi[1], (long long)oy + dy - 1);
    const long long z0 = max((long long)lo[2], (long long)oz), z1 = min((long long)hi[2], (long long)oz + dz - 1);
    if (x0 > x1 || y0 > y1 || z0 > z1) return;
    const unsigned long long nx = x1 - x0 + 1, ny = y1 - y0 + 1, nz = z1 - z0 + 1, tot = nx * ny * nz;
    for (unsigned long long c = lane; c < tot; c += 32) {
        const unsigned long long cx = c % nx, cy = (c / nx) % ny, cz = c / (nx * ny);
        const unsigned long long bit = (unsigned long long)(x0 - ox + cx) +
                                       (unsigned long long)dx * ((y0 - oy + cy) + (unsigned long long)dy * (z0 - oz + cz));
        atomicOr(&bitmap[bit >> 5], 1u << (bit & 31u));
    }
}

// ---- pair-set export (gp_pairs): compact flagged pairs into (key = rank_a<<32|rank_b, value = pair id) ----
__global__ void k_export_pairs(const uint4* __restrict__ pairs, uint32_t np, int which,
                               unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals, uint32_t* counter) {
    const uint32_t pid = blockIdx.x * blockDim.x + threadIdx.x;
    if (pid >= np) return;
    const uint4 r = pairs[2 * (size_t)pid];
    const bool in = which == 0 ? (r.z & TOPBIT) != 0 : (r.w & TOPBIT) != 0;
    if (!in) return;
    cg::coalesced_group g = cg::coalesced_threads();
    uint32_t base = 0;
    if (g.thread_rank() == 0) base = atomicAdd(counter, g.size());
    base = g.shfl(base, 0);
    const uint32_t s = base + g.thread_rank();
    keys[s] = ((unsigned long long)(r.z & ~TOPBIT) << 32) | (r.w & ~TOPBIT);
    vals[s] = pid;
}

__global__ void k_gather_pairs(const uint4* __restrict__ pairs, const uint32_t* __restrict__ order, uint32_t m,
                               const Bodies B, uint2* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const uint4 r = pairs[2 * (size_t)order[i]];
    // pair records hold slots of the current (cell-sorted) set; report body ids
    out[i] = make_uint2(__float_as_uint(B.Lo(r.x & ~TOPBIT).w), __float_as_uint(B.Lo(r.y & ~TOPBIT).w));
}

}  // namespace gp

// gp_common.cuh — constants of the reference, the device-resident control block, the grid, vector helpers.
// Part of the physics step; see gp_kernels.cuh for the map of kernels to reference lines.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

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
constexpr float MAX_HORIZONTAL_VELOCITY = 20.0f;  // cap_velocity only (physics.rs:7-8)
constexpr float MAX_VERTICAL_VELOCITY = 40.0f;
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
    uint32_t tiny_fallback;      // the per-component repair met a component it does not handle: the general repair runs
    uint32_t peer_dead;          // latched: a slab neighbour timed out (k_wait_flags); nothing is appended any more
    uint32_t sweep_round;        // level the sweep continues with after CTA 0 ran the tail levels alone
    uint32_t dbg_tail_rounds;    // since creation: levels run by CTA 0 alone
    unsigned long long dbg_level_ns[16], dbg_level_n[16];  // first sweep: time and pairs per dependency level (GP_PROF_FINE)
    unsigned long long dbg_counts[4];  // since creation: escapees, requeried pairs, activated watch pairs, repair rounds
    unsigned long long dbg_phase_ns[12];  // since creation: time per phase of k_contacts as seen by thread 0 (GP_PROF_FINE)
    uint32_t n_escapees_total;   // stats: escapees of the first sweep
    uint32_t dbg_tiny[6];        // since creation: rounds done by the tiny repair, rounds that fell back, components
                                 // repaired, components too large, largest component seen (bodies, pairs)
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
constexpr uint32_t SF_INTERNAL = 64u;     // (unused since the TMA pair find was retired)
constexpr uint32_t SF_PEER_TIMEOUT = 256u;  // slab exchange: a neighbour never arrived; the world is dead until re-uploaded

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

// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256): two adjacent float4, 32 B-aligned.
__device__ __forceinline__ void ld256(const float4* p, float4& a, float4& b) {
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p)
                 : "memory");
}
__device__ __forceinline__ void st256(float4* p, const float4 a, const float4 b) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w),
                 "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w)
                 : "memory");
}

// scratch of the component-local repair
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

// everything the contact kernel (k_contacts) works on
struct ContactArgs {
    Bodies B1;                   // the step's output set, in cell order: swept in place
    Bodies B0;                   // the untouched input set (a repair re-integrates from it) ...
    const float4* A0;            // ... and its acc|flags
    const float4* A1;            // acc|flags of the output set
    const float4* sbox;          // AABB records of the step (written by K3, read-only here)
    const uint32_t* LB;          // cell start table (read-only here)
    uint4* pairs;
    uint32_t pair_cap;
    uint2* watch;
    uint32_t *frontier0, *frontier1;
    Ctl *ctl, *ctl_snap;
    float *dmax, *reach;
    uint32_t *moved_bits, *esc_bits, *escapees;
    uint32_t esc_cap;
    const uint32_t* offs;        // K5 adjacency (read-only here)
    const unsigned long long* adj;
    uint32_t* chain_cnt;
    const uint32_t* perm;
    uint32_t *xhead, *xnext;
    RepairArrays ra;
    uint32_t* recs;              // component records of the per-component repair (aliases ra.adj2)
    uint32_t rec_cap;
    uint32_t* taint_bits;        // slab worlds: per-step exactness proof (nullptr otherwise)
    uint32_t* seed_bits;         // ... and the seeds K3 marked for it
    uint32_t* uf_parent;         // ... and its union-find forest, reset by K3 (parent[i] = i)
    float dt, d16, d18, d35;
    uint32_t start_phase;        // 0: sweep first; 1: continue the repair loop
    uint32_t max_rounds;         // repair rounds before the step is declared still-growing (SF_MARGIN)
    uint32_t tiny;               // per-component warp repair in front of the cooperative one
    int latch;                   // finish_step mode
    uint32_t profile;            // GP_PROF_FINE: thread 0 keeps per-phase times in ctl->dbg_phase_ns
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// phase clock of k_contacts (thread 0 of the grid only): charge the time since the last stamp to `phase`
struct PhaseClock {
    unsigned long long t;
    bool on;
    __device__ __forceinline__ void start(bool enabled) {
        on = enabled && blockIdx.x == 0 && threadIdx.x == 0;
        if (on) t = globaltimer_ns();
    }
    __device__ __forceinline__ void stamp(Ctl* ctl, int phase, int level = -1, uint32_t level_pairs = 0) {
        if (!on) return;
        const unsigned long long now = globaltimer_ns();
        ctl->dbg_phase_ns[phase] += now - t;
        if (level >= 0) {
            const int l = level < 15 ? level : 15;
            ctl->dbg_level_ns[l] += now - t;
            ctl->dbg_level_n[l] += level_pairs;
        }
        t = now;
    }
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


}  // namespace gp

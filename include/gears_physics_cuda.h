/* gears_physics_cuda.h — C ABI of libgears_physics_cuda.so (the `gears-physics-cuda-sys` boundary).
 *
 * B200-native (sm_100a) replacement for passes 1 and 2 of Gears' `update_physics` system
 *   reference crates/gears-app/src/systems/update_systems.rs:362-487
 * and the component math it calls
 *   reference crates/gears-ecs/src/components/physics.rs:84-299,405-462.
 * The reference has no FFI of its own (every crate is #![forbid(unsafe_code)]); the functions below
 * are what a `gears-physics-cuda-sys` crate binds (rust/gears-physics-cuda-sys).
 *
 * Conventions
 *  - plain pointers and sizes only; all device memory is owned by the gp_world handle;
 *  - every function returns a gp_status (0 = ok, <0 = error, >0 = warning); never aborts, never throws;
 *  - a handle may be used from any thread but by one thread at a time (not re-entrant);
 *  - host arrays are tightly packed: vec3 = 3 floats, quaternion = 4 floats in cgmath order
 *    (v.x, v.y, v.z, s) as laid out by `Quaternion<f32>` (reference transforms.rs:17-22);
 *  - body i of an upload is "body index i" in every later call and in every pair output;
 *  - rank[i] is the position of body i in the reference's iteration order, i.e. the index of its
 *    entity in the Vec returned by World::get_entities_with_component (reference
 *    crates/gears-ecs/src/lib.rs:422-433).  Ranks must be distinct.  NULL means rank[i] = i.
 *  - arithmetic is IEEE binary32 without FMA contraction, so results are bit-identical to the
 *    reference's scalar Rust for any schedule that keeps the per-body order of pair updates.
 */
#ifndef GEARS_PHYSICS_CUDA_H
#define GEARS_PHYSICS_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GP_ABI_VERSION 1

typedef struct gp_world gp_world;

typedef enum gp_status {
    GP_OK = 0,
    /* warnings (> 0): the call completed */
    GP_WARN_MARGIN_RETRIED = 1, /* bodies out-ran the broad-phase margin or the pair buffers had to grow:  */
                                /* the step repaired itself / was re-run; the result is exact            */
    /* errors (< 0) */
    GP_ERR_BAD_ARG = -1,
    GP_ERR_CUDA = -2,        /* CUDA runtime error, see gp_last_error                                   */
    GP_ERR_NO_DEVICE = -3,   /* no sm_100 device / library built without a usable kernel image         */
    GP_ERR_CAPACITY = -4,    /* more bodies than max_bodies                                             */
    GP_ERR_PAIR_OVERFLOW = -5, /* candidate pairs exceeded max_pairs (step not applied)                 */
    GP_ERR_MARGIN = -6,      /* a body moved further than margin_max during contact resolution          */
                             /* (step result not guaranteed exact; raise margin_max)                    */
    GP_ERR_NCCL = -7,
    GP_ERR_STATE = -8,       /* call order violation (e.g. step before upload)                          */
    GP_ERR_PEER_TIMEOUT = -9 /* slab worlds: a neighbour rank never arrived at the exchange (GP_PEER_TIMEOUT_MS,
                                default 30 s).  Records that were migrating are lost: the world refuses further
                                steps until gp_upload + gp_comm_init are called again                        */
} gp_status;

/* gp_config.flags */
#define GP_FLAG_RECORD_PAIRS 0x1u /* keep S_snapshot / S_sweep flags so gp_pairs can be called       */
#define GP_FLAG_NO_RETRY 0x2u     /* never re-run a step: report GP_ERR_MARGIN/PAIR_OVERFLOW instead  */
#define GP_FLAG_BIG_STATIC_ONLY 0x8u /* only STATIC bodies may go to the brute-force list (the grid cell grows to  */
                                    /* host every dynamic body): required for slab worlds, which replicate big   */
                                    /* bodies on every rank instead of exchanging them                           */
#define GP_FLAG_PROFILE 0x4u      /* bracket every kernel group of a step with CUDA events on the step */
                                  /* stream; read the sums with gp_kernel_times                        */

typedef struct gp_config {
    uint32_t struct_size; /* = sizeof(gp_config); lets the library version the struct                */
    int32_t device;       /* CUDA device ordinal                                                      */
    uint32_t max_bodies;  /* capacity; 0 = sized by the first gp_upload                               */
    uint64_t max_pairs;   /* candidate-pair capacity; 0 = 16 * max_bodies (min 1 Mi)                   */
    float margin_init;    /* initial (and minimum) OUTER broad-phase margin mu, RELATIVE: every AABB   */
                          /* is inflated by mu * (its largest extent); <=0 = 0.05.  Adapts per step.  */
                          /* (Candidates within a smaller, self-tuning inner margin enter the contact */
                          /* sweep at once, the rest wait on a watch list: DESIGN.md section 1.2)     */
    float margin_max;     /* upper bound for the adaptive mu; <=0 = 2.0                               */
    float cell_size;      /* minimum uniform-grid cell edge; <=0 = auto: at least (largest grid-class  */
                          /* extent) * (1 + 2 mu), and no finer than ~0.3 body per cell; re-derived on */
                          /* the device whenever mu changes                                           */
    uint32_t max_big;     /* bodies larger than the cell go to a brute-force list of at most this     */
                          /* many (0 = 64); the cell grows until the rest fit                         */
    uint32_t flags;
    float grid_min[3], grid_max[3]; /* optional world bounds for the grid; all 0 = from the upload     */
} gp_config;

typedef struct gp_step_stats {
    uint32_t n_bodies;
    uint32_t n_big;          /* bodies on the brute-force list                                        */
    uint64_t n_candidates;   /* broad-phase pairs (margin-inflated AABBs overlap)                     */
    uint64_t n_snapshot;     /* |S_snapshot|: exact overlaps on the post-integration state            */
    uint64_t n_contacts;     /* |S_sweep|: pairs whose exact test was true at their turn              */
    uint32_t n_rounds;       /* dependency levels executed by the contact sweep                       */
    uint32_t n_retries;      /* re-runs for capacity + repair sweeps (component-local) of this step   */
    float margin;            /* relative margin mu used for this step                                 */
    float max_displacement;  /* largest per-axis move of any AABB during the sweep / its extent       */
    float cell_size;
    uint32_t grid_dim[3];
    uint32_t n_heavy;        /* bodies whose contact chain needed the block-wide sort                 */
    uint32_t n_migrated_out, n_migrated_in, n_halo_out, n_halo_in; /* multi-GPU only               */
    uint64_t steps_total;    /* steps since upload                                                    */
    uint64_t steps_inexact;  /* async steps (gp_step_async) that hit GP_ERR_MARGIN/PAIR_OVERFLOW      */
    uint32_t n_unverified;   /* multi-GPU: owned bodies of this step whose result could depend on a   */
                             /* body this rank has not seen (0 = the step is provably identical to    */
                             /* the single-GPU result)                                                */
    uint64_t unverified_total; /* the same, summed over the steps since the last gp_sync             */
    uint32_t n_watch;        /* candidates that only went on the watch list (lazy candidates)         */
    uint32_t n_activated;    /* watch pairs that had to join the sweep                                */
} gp_step_stats;

/* ---- lifetime ---- */
int32_t gp_abi_version(void);
gp_status gp_create(const gp_config* cfg, gp_world** out);
void gp_destroy(gp_world* w);
/* NUL-terminated description of the last error on this handle (or of the last gp_create failure if w is NULL) */
const char* gp_last_error(const gp_world* w);

/* ---- component mirror: host SoA -> device (replaces the per-entity acquire_query/get/write of
 *      update_systems.rs:381-400 and :414-475) ---- */
gp_status gp_upload(gp_world* w, uint32_t n, const float* pos3, const float* rot4 /*nullable: identity*/,
                    const float* scale3 /*nullable: (1,1,1) = no Scale component*/, const float* vel3,
                    const float* acc3 /*nullable: 0*/, const float* box_min3, const float* box_max3,
                    const float* mass, const float* restitution, const uint8_t* is_static,
                    const uint32_t* entity /*nullable: i*/, const uint32_t* rank /*nullable: i*/);

/* Refresh the mutable state of ALL bodies from the host (external writers: user systems and
 * MovementController write velocity/acceleration/position between steps).
 * Any pointer may be NULL = keep the device value. */
gp_status gp_set_state(gp_world* w, const float* pos3, const float* vel3, const float* acc3);

/* Same for k listed bodies (dirty set). */
gp_status gp_update_dirty(gp_world* w, uint32_t k, const uint32_t* idx, const float* pos3, const float* vel3,
                          const float* acc3);

/* Shape / material change of k listed bodies (rotation, scale, box, mass, restitution, static flag).
 * All arrays required except rot4/scale3 (NULL = identity / none). */
gp_status gp_update_shape(gp_world* w, uint32_t k, const uint32_t* idx, const float* box_min3,
                          const float* box_max3, const float* rot4, const float* scale3, const float* mass,
                          const float* restitution, const uint8_t* is_static);

/* ---- the step: pass 1 (update_pos for every body) + pass 2 (ordered pair sweep) ---- */
/* damp3 = { expf(-1.6*dt), expf(-1.8*dt), expf(-3.5*dt) } from the host libm (what Rust's f32::exp
 * calls); NULL = computed here with expf.  Synchronous: returns after the step completed, re-running
 * it transparently when the adaptive margin or the pair capacity was exceeded (unless GP_FLAG_NO_RETRY).
 * stats may be NULL. */
gp_status gp_step(gp_world* w, float dt, const float* damp3, gp_step_stats* stats);

/* Enqueue `nsteps` steps without any host synchronisation.  Failures latch: the first failing step's
 * status is returned by the next gp_sync/gp_step/gp_download and counted in stats.steps_inexact. */
gp_status gp_step_async(gp_world* w, float dt, const float* damp3, uint32_t nsteps);
gp_status gp_sync(gp_world* w, gp_step_stats* stats /*nullable: stats of the last step*/);

/* ---- results: device -> host (replaces the write-back through the RwLock guards) ---- */
gp_status gp_download(gp_world* w, float* pos3 /*nullable*/, float* vel3 /*nullable*/);

/* Pair sets of the LAST step (needs GP_FLAG_RECORD_PAIRS), in the reference's iteration order
 * (ascending (rank_a, rank_b), rank_a < rank_b).  which: 0 = S_snapshot, 1 = S_sweep.
 * pairs2 receives body indices (a, b), a = earlier in iteration order.  *n = number of pairs in the
 * set even if cap is smaller (then only cap pairs are written). */
gp_status gp_pairs(gp_world* w, int32_t which, uint32_t* pairs2, uint64_t cap, uint64_t* n);

/* ---- helpers ---- */
void gp_damping_factors(float dt, float* out3);
/* page-locked host buffers for the mirror (faster gp_upload/gp_set_state/gp_download) */
void* gp_host_alloc(size_t bytes);
void gp_host_free(void* p);
/* device time of the last gp_step / gp_step_async batch in milliseconds (CUDA events on the step stream) */
gp_status gp_last_step_ms(gp_world* w, float* ms);
/* number of kernels launched by this handle since creation (for launch accounting) */
uint64_t gp_kernel_launches(const gp_world* w);

/* Per-kernel-group device time (needs GP_FLAG_PROFILE): sums over the steps enqueued since the last
 * call, measured with CUDA events recorded on the step stream between the groups. */
#define GP_NUM_KERNEL_GROUPS 8
typedef struct gp_kernel_times {
    uint32_t steps;                       /* steps covered by the sums                                */
    float ms[GP_NUM_KERNEL_GROUPS];       /* 0 cell keys + histogram (K1; on a slab also packing/sending),  */
                                          /* 1 slab exchange: wait for the neighbours + append (multi-GPU),  */
                                          /* 2 cell scan (K2), 3 scatter into cell order + integrate + AABBs */
                                          /* (K3), 4 pair find (K4), 5 chain build (K5: scan, fill, sort +   */
                                          /* link), 6 contact sweep + verification + repair rounds (K6),     */
                                          /* 7 finish                                                        */
    uint32_t launches[GP_NUM_KERNEL_GROUPS]; /* kernel launches per group over those steps            */
} gp_kernel_times;
gp_status gp_kernel_times_read(gp_world* w, gp_kernel_times* out);

/* RigidBody::cap_velocity (physics.rs:506-520) on the device, for the bodies idx[0..k) (upload indices; idx == NULL:
 * every body): what MovementController::update_pos does after writing the player's velocity (controllers.rs:307).
 * Asynchronous on the step stream (ordered before the next step); idx is copied before the call returns. */
gp_status gp_cap_velocity(gp_world* w, uint32_t k, const uint32_t* idx);

/* ---- pipelined frames (async H2D / D2H overlapped with the step) ----
 * For hosts that own the state (every frame the ECS side hands over pos / vel / acc of ALL bodies, in upload order,
 * and wants pos / vel back -- SURVEY.md section 3.3: external systems may have written any of them):
 *   gp_frame_submit  enqueues  H2D of the three arrays (copy stream)  ->  whole-world state refresh + one step
 *                    (step stream)  ->  read-back + D2H into state6_out (download stream)  and returns at once.
 *                    Two frames may be in flight: frame t+1 uploads while frame t computes and frame t-1 downloads
 *                    (PCIe is full duplex), so a frame costs max(H2D, kernels, D2H) instead of their sum.
 *                    All host buffers must be pinned (gp_host_alloc) and must stay untouched until the frame has
 *                    been waited for.  A frame REPLACES the mutable state: its result does not depend on the previous
 *                    frame's result, only on what the host passes in (and on shapes / masses set by gp_upload /
 *                    gp_update_shape).
 *   state6_out       n x 6 floats, row = upload index: pos.xyz, vel.xyz.
 *   gp_frame_wait    blocks until at most max_in_flight submitted frames are unfinished (0 = all; then it also
 *                    reports latched step failures like gp_sync and fills stats).
 * Not available on a slab world. */
gp_status gp_frame_submit(gp_world* w, const float* pos3, const float* vel3, const float* acc3, float dt,
                          const float damp3[3], float* state6_out);
gp_status gp_frame_wait(gp_world* w, uint32_t max_in_flight, gp_step_stats* stats);

/* ---- multi-GPU: spatial slabs along x, one gp_world per GPU ----
 * The reference is a single process (no collectives of any kind); this decomposition is new.
 * A rank owns the bodies whose AABB centre x lies in [x_lo, x_hi) (use -INFINITY / +INFINITY at the two ends).
 * Upload each rank's bodies with gp_upload (entity = GLOBAL ids, rank = GLOBAL iteration ranks; big static bodies
 * such as a floor are uploaded on EVERY rank), then attach the worlds:
 *   - one process per GPU:   gp_comm_unique_id on rank 0, broadcast it, gp_comm_init on every rank (collective);
 *                            then gp_step_async / gp_sync as usual: every step exchanges migrating bodies and halo
 *                            copies with the two neighbours on the step stream -- K1 writes the records straight
 *                            into the neighbour's buffer over NVLink (CUDA IPC peer memory, release/acquire step
 *                            flags); ncclSend/ncclRecv when peer access is missing or GP_SLAB_TRANSPORT=nccl;
 *   - one process, n worlds: gp_comm_init_local, then gp_step_group (peer copies instead of NCCL; the worlds may
 *                            even share a device, which is how the slab logic is tested on one GPU).
 * Per-index calls (gp_set_state, gp_update_dirty, gp_update_shape, gp_download) are not available on a slab world;
 * read results with gp_owned_count / gp_download_owned. */
#define GP_UNIQUE_ID_BYTES 128
#define GP_SLAB_NO_VERIFY 0x1u /* skip the per-step proof that owned bodies cannot depend on unseen bodies */
typedef struct gp_slab_config {
    uint32_t struct_size;  /* = sizeof(gp_slab_config)                                                      */
    int32_t rank, nranks;  /* slab index along x, number of slabs                                           */
    float x_lo, x_hi;      /* this rank's slab                                                              */
    float halo;            /* coverage H beyond a face that the neighbour mirrors; <=0 = 4 x the largest    */
                           /* grid-class extent.  Inner slabs must be >= 2 H wide                            */
    uint32_t max_exchange; /* records per face buffer; 0 = 2 x (bodies within 2 H of a face today) + 8192.   */
                           /* Both sides of a face must use the same value (messages have a fixed size)     */
    uint32_t flags;        /* GP_SLAB_*                                                                     */
} gp_slab_config;
gp_status gp_comm_unique_id(uint8_t id[GP_UNIQUE_ID_BYTES]);
gp_status gp_comm_init(gp_world* w, const uint8_t id[GP_UNIQUE_ID_BYTES], const gp_slab_config* cfg);
gp_status gp_comm_init_local(gp_world* const* worlds, int32_t n, const gp_slab_config* cfgs);
gp_status gp_step_group(gp_world* const* worlds, int32_t n, float dt, const float* damp3, uint32_t nsteps);
/* bodies owned by this rank now (after migration); arrays of gp_owned_count elements, in device (slot) order.
 * unverified (nullable): 1 for a body the LAST step could not prove independent of unseen bodies.
 * gp_set_state_owned writes rows back (what external systems changed): row t = the t-th body of the LAST
 * gp_download_owned, k = its count; only valid while no step has run since.  NULL arrays are left as they are. */
gp_status gp_owned_count(gp_world* w, uint32_t* n);
gp_status gp_download_owned(gp_world* w, uint32_t* entity, float* pos3, float* vel3, uint8_t* unverified);
gp_status gp_set_state_owned(gp_world* w, uint32_t k, const float* pos3, const float* vel3, const float* acc3);
/* Re-balance the slab faces between two steps: equal-count quantiles of the owned bodies'
 * centre x over all ranks; a face moves by at most max_shift per call (<= 0: half the halo), so that the bodies that
 * change owner in the next step fit the ordinary migration path; a call that would make a slab narrower than twice
 * the halo changes nothing.  gp_slab_rebalance is COLLECTIVE (every rank of the communicator calls it at the same
 * step); *moved_out (nullable) = 1 if one of this rank's faces moved. */
gp_status gp_slab_rebalance(gp_world* w, float max_shift, uint32_t* moved_out);
gp_status gp_slab_rebalance_local(gp_world* const* worlds, int32_t n, float max_shift, uint32_t* moved_out);

/* ---- next rows: consumers of the body state right after the step ----
 * All three address bodies by upload index (single-GPU worlds only). */
/* N1  pass 3 of update_physics (update_systems.rs:489-527): Instance::to_raw (gears-renderer instance.rs:22-31) of
 * every body in one launch.  out25: n rows of 25 floats = 4x4 model matrix + 3x3 normal matrix, both column-major
 * as cgmath's `.into()` lays them out (= struct InstanceRaw, 100 B).  rot4 / scale3: host arrays in upload order
 * (Pos3.rot in cgmath order; Instance.scale); NULL = identity / (1,1,1). */
gp_status gp_instances(gp_world* w, const float* rot4, const float* scale3, float* out25);
/* N3  Weapon::ray_intersects_aabb (gears-ecs interactive.rs:94-158), batched: hits[r * ntargets + t] = 1 if ray r hits
 * body targets[t] (targets NULL: all bodies, ntargets ignored = n).  Like the reference, the test uses
 * pos + collision_box.min/max and ignores rotation and scale. */
gp_status gp_raycast(gp_world* w, uint32_t nrays, const float* origin3, const float* dir3, uint32_t ntargets,
                     const uint32_t* targets, uint8_t* hits);
/* N4  PathfindingGrid::add_obstacle (gears-ecs pathfinding.rs:296-338) for k bodies (idx NULL: all): sets the bits of
 * the cells GridPos::from_world_pos(pos + box.min) ..= from_world_pos(pos + box.max) (pathfinding.rs:57-63) in a dense
 * bitmap over the window [origin, origin + dims): bit = (x-ox) + dims[0]*((y-oy) + dims[1]*(z-oz)); cells outside the
 * window are dropped.  bounds_out (nullable): min xyz / max xyz cell over all obstacles, as the reference tracks. */
gp_status gp_obstacle_grid(gp_world* w, uint32_t k, const uint32_t* idx, float cell_size, const int32_t origin[3],
                           const uint32_t dims[3], uint32_t* bitmap, int32_t bounds_out[6]);

/* ---- test hooks (used by tests/ only; stable but not part of the drop-in surface) ---- */
gp_status gp_test_radix_sort(int32_t device, uint32_t n, uint32_t key_bits, uint32_t* keys_inout,
                             uint32_t* vals_inout);
gp_status gp_test_radix_sort64(int32_t device, uint32_t n, uint32_t key_bits, uint64_t* keys_inout,
                               uint32_t* vals_inout);
gp_status gp_test_exclusive_scan(int32_t device, uint32_t n, uint32_t* data_inout, uint32_t* total);

#ifdef __cplusplus
}
#endif
#endif /* GEARS_PHYSICS_CUDA_H */

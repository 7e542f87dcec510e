//! Raw bindings to `libgears_physics_cuda.so` — a 1:1 mirror of `include/gears_physics_cuda.h`.
//!
//! The reference engine (benditorok/gears) has no FFI of its own (`#![forbid(unsafe_code)]` in every crate);
//! this `-sys` crate is the only place where `unsafe` enters.  NOT compiled in this repository's environment
//! (no cargo/rustc in the image): keep it in lock step with the header by hand — the symbol list is checked
//! against the built library by `tests/test_capi_symbols.py`.
#![allow(non_camel_case_types)]

use core::ffi::{c_char, c_void};

#[repr(C)]
pub struct gp_world {
    _private: [u8; 0],
}

pub type gp_status = i32;
pub const GP_OK: gp_status = 0;
pub const GP_WARN_MARGIN_RETRIED: gp_status = 1;
pub const GP_ERR_BAD_ARG: gp_status = -1;
pub const GP_ERR_CUDA: gp_status = -2;
pub const GP_ERR_NO_DEVICE: gp_status = -3;
pub const GP_ERR_CAPACITY: gp_status = -4;
pub const GP_ERR_PAIR_OVERFLOW: gp_status = -5;
pub const GP_ERR_MARGIN: gp_status = -6;
pub const GP_ERR_NCCL: gp_status = -7;
pub const GP_ERR_STATE: gp_status = -8;
pub const GP_ERR_PEER_TIMEOUT: gp_status = -9;

pub const GP_FLAG_RECORD_PAIRS: u32 = 0x1;
pub const GP_FLAG_NO_RETRY: u32 = 0x2;
pub const GP_FLAG_PROFILE: u32 = 0x4;
pub const GP_FLAG_BIG_STATIC_ONLY: u32 = 0x8;
pub const GP_SLAB_NO_VERIFY: u32 = 0x1;
pub const GP_UNIQUE_ID_BYTES: usize = 128;
pub const GP_NUM_KERNEL_GROUPS: usize = 8;

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct gp_config {
    pub struct_size: u32,
    pub device: i32,
    pub max_bodies: u32,
    pub max_pairs: u64,
    pub margin_init: f32,
    pub margin_max: f32,
    pub cell_size: f32,
    pub max_big: u32,
    pub flags: u32,
    pub grid_min: [f32; 3],
    pub grid_max: [f32; 3],
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct gp_step_stats {
    pub n_bodies: u32,
    pub n_big: u32,
    pub n_candidates: u64,
    pub n_snapshot: u64,
    pub n_contacts: u64,
    pub n_rounds: u32,
    pub n_retries: u32,
    pub margin: f32,
    pub max_displacement: f32,
    pub cell_size: f32,
    pub grid_dim: [u32; 3],
    pub n_heavy: u32,
    pub n_migrated_out: u32,
    pub n_migrated_in: u32,
    pub n_halo_out: u32,
    pub n_halo_in: u32,
    pub steps_total: u64,
    pub steps_inexact: u64,
    pub n_unverified: u32,
    pub unverified_total: u64,
    pub n_watch: u32,
    pub n_activated: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct gp_slab_config {
    pub struct_size: u32,
    pub rank: i32,
    pub nranks: i32,
    pub x_lo: f32,
    pub x_hi: f32,
    pub halo: f32,
    pub max_exchange: u32,
    pub flags: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct gp_kernel_times {
    pub steps: u32,
    pub ms: [f32; GP_NUM_KERNEL_GROUPS],
    pub launches: [u32; GP_NUM_KERNEL_GROUPS],
}

#[link(name = "gears_physics_cuda")]
unsafe extern "C" {
    pub fn gp_abi_version() -> i32;
    pub fn gp_create(cfg: *const gp_config, out: *mut *mut gp_world) -> gp_status;
    pub fn gp_destroy(w: *mut gp_world);
    pub fn gp_last_error(w: *const gp_world) -> *const c_char;

    pub fn gp_upload(
        w: *mut gp_world, n: u32, pos3: *const f32, rot4: *const f32, scale3: *const f32, vel3: *const f32,
        acc3: *const f32, box_min3: *const f32, box_max3: *const f32, mass: *const f32, restitution: *const f32,
        is_static: *const u8, entity: *const u32, rank: *const u32,
    ) -> gp_status;
    pub fn gp_set_state(w: *mut gp_world, pos3: *const f32, vel3: *const f32, acc3: *const f32) -> gp_status;
    pub fn gp_update_dirty(
        w: *mut gp_world, k: u32, idx: *const u32, pos3: *const f32, vel3: *const f32, acc3: *const f32,
    ) -> gp_status;
    pub fn gp_update_shape(
        w: *mut gp_world, k: u32, idx: *const u32, box_min3: *const f32, box_max3: *const f32, rot4: *const f32,
        scale3: *const f32, mass: *const f32, restitution: *const f32, is_static: *const u8,
    ) -> gp_status;

    pub fn gp_step(w: *mut gp_world, dt: f32, damp3: *const f32, stats: *mut gp_step_stats) -> gp_status;
    pub fn gp_step_async(w: *mut gp_world, dt: f32, damp3: *const f32, nsteps: u32) -> gp_status;
    pub fn gp_sync(w: *mut gp_world, stats: *mut gp_step_stats) -> gp_status;

    pub fn gp_download(w: *mut gp_world, pos3: *mut f32, vel3: *mut f32) -> gp_status;
    pub fn gp_pairs(w: *mut gp_world, which: i32, pairs2: *mut u32, cap: u64, n: *mut u64) -> gp_status;

    pub fn gp_damping_factors(dt: f32, out3: *mut f32);
    pub fn gp_host_alloc(bytes: usize) -> *mut c_void;
    pub fn gp_host_free(p: *mut c_void);
    pub fn gp_last_step_ms(w: *mut gp_world, ms: *mut f32) -> gp_status;
    pub fn gp_kernel_launches(w: *const gp_world) -> u64;
    pub fn gp_kernel_times_read(w: *mut gp_world, out: *mut gp_kernel_times) -> gp_status;

    pub fn gp_comm_unique_id(id: *mut u8) -> gp_status;
    pub fn gp_comm_init(w: *mut gp_world, id: *const u8, cfg: *const gp_slab_config) -> gp_status;
    pub fn gp_comm_init_local(worlds: *const *mut gp_world, n: i32, cfgs: *const gp_slab_config) -> gp_status;
    pub fn gp_step_group(worlds: *const *mut gp_world, n: i32, dt: f32, damp3: *const f32, nsteps: u32) -> gp_status;
    pub fn gp_owned_count(w: *mut gp_world, n: *mut u32) -> gp_status;
    pub fn gp_download_owned(
        w: *mut gp_world, entity: *mut u32, pos3: *mut f32, vel3: *mut f32, unverified: *mut u8,
    ) -> gp_status;

    // next rows: pass-3 instance rows, batched ray-vs-AABB, occupancy grid (see the header for the layouts)
    pub fn gp_instances(w: *mut gp_world, rot4: *const f32, scale3: *const f32, out25: *mut f32) -> gp_status;
    pub fn gp_raycast(
        w: *mut gp_world, nrays: u32, origin3: *const f32, dir3: *const f32, ntargets: u32, targets: *const u32,
        hits: *mut u8,
    ) -> gp_status;
    pub fn gp_obstacle_grid(
        w: *mut gp_world, k: u32, idx: *const u32, cell_size: f32, origin: *const i32, dims: *const u32,
        bitmap: *mut u32, bounds_out: *mut i32,
    ) -> gp_status;

    pub fn gp_slab_rebalance(w: *mut gp_world, max_shift: f32, moved_out: *mut u32) -> gp_status;
    pub fn gp_slab_rebalance_local(worlds: *const *mut gp_world, n: i32, max_shift: f32, moved_out: *mut u32) -> gp_status;
    pub fn gp_set_state_owned(w: *mut gp_world, k: u32, pos3: *const f32, vel3: *const f32, acc3: *const f32) -> gp_status;

    // RigidBody::cap_velocity (physics.rs:506-520) for the bodies idx[0..k) (null: all)
    pub fn gp_cap_velocity(w: *mut gp_world, k: u32, idx: *const u32) -> gp_status;

    // pipelined frames: async H2D / step / D2H, two frames in flight (pinned host buffers)
    pub fn gp_frame_submit(
        w: *mut gp_world, pos3: *const f32, vel3: *const f32, acc3: *const f32, dt: f32, damp3: *const f32,
        state6_out: *mut f32,
    ) -> gp_status;
    pub fn gp_frame_wait(w: *mut gp_world, max_in_flight: u32, stats: *mut gp_step_stats) -> gp_status;

    // test hooks (not part of the drop-in surface)
    pub fn gp_test_radix_sort(device: i32, n: u32, key_bits: u32, keys_inout: *mut u32, vals_inout: *mut u32)
        -> gp_status;
    pub fn gp_test_radix_sort64(device: i32, n: u32, key_bits: u32, keys_inout: *mut u64, vals_inout: *mut u32)
        -> gp_status;
    pub fn gp_test_exclusive_scan(device: i32, n: u32, data_inout: *mut u32, total: *mut u32) -> gp_status;
}

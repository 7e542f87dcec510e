//! `gears-physics-cuda` — drop-in replacement for passes 1 and 2 of Gears' `update_physics` system
//! (reference `crates/gears-app/src/systems/update_systems.rs:362-487`) on an NVIDIA B200.
//!
//! The system keeps the engine's ECS-facing surface: it reads and writes the same `Pos3`, `RigidBody<AABBCollisionBox>`
//! and `Scale` components through the unchanged `World` API, mirrors them into the library's device buffers, runs the
//! step, and scatters positions and velocities back.  Pass 3 (instance upload, `update_systems.rs:489-527`) stays in
//! gears-app unchanged.
//!
//! NOT compiled in this repository's environment (no Rust toolchain in the image).  The same call sequence is what
//! `gears_b200/world.py` (ctypes) exercises in the test-suite.
use std::ffi::CStr;
use std::sync::{Arc, Mutex};
use std::time::Duration;

use gears_ecs::components::physics::{AABBCollisionBox, RigidBody};
use gears_ecs::components::transforms::{Pos3, Scale};
use gears_ecs::query::{ComponentQuery, WorldQueryExt};
use gears_ecs::{Entity, World};
use gears_physics_cuda_sys as sys;

/// Mirrors `SystemError::Execution(String)` of gears-app (`systems/errors.rs:20`); gears-app maps it 1:1.
#[derive(Debug)]
pub struct PhysicsCudaError(pub String);

/// Owns one `gp_world`.  `Send` (every entry point sets the device), not re-entrant: keep it behind a `Mutex`.
pub struct CudaPhysics {
    handle: *mut sys::gp_world,
    /// entities that were actually gathered at the last step, in iteration order (= `rank`); an entity without a
    /// `Pos3` is skipped exactly like the reference skips its pairs (update_systems.rs:421-432)
    entities: Vec<Entity>,
    staging: Staging,
}
unsafe impl Send for CudaPhysics {}

#[derive(Default)]
struct Staging {
    pos: Vec<f32>, rot: Vec<f32>, scale: Vec<f32>, vel: Vec<f32>, acc: Vec<f32>,
    bmin: Vec<f32>, bmax: Vec<f32>, mass: Vec<f32>, rest: Vec<f32>, stat: Vec<u8>, ids: Vec<u32>,
}

impl CudaPhysics {
    pub fn new(device: i32) -> Result<Self, PhysicsCudaError> {
        let cfg = sys::gp_config { struct_size: size_of::<sys::gp_config>() as u32, device, ..Default::default() };
        let mut h = std::ptr::null_mut();
        let st = unsafe { sys::gp_create(&cfg, &mut h) };
        if st < 0 {
            return Err(PhysicsCudaError(last_error(std::ptr::null())));
        }
        Ok(Self { handle: h, entities: Vec::new(), staging: Staging::default() })
    }

    /// One physics step over the bodies `entities` (in the order `World::get_entities_with_component` returned them:
    /// that order IS the reference's pair order, `crates/gears-ecs/src/lib.rs:422-433`).
    pub fn step(&mut self, world: &World, entities: &[Entity], dt: Duration) -> Result<(), PhysicsCudaError> {
        let n = entities.len();
        let s = &mut self.staging;
        for v in [&mut s.pos, &mut s.scale, &mut s.vel, &mut s.acc, &mut s.bmin, &mut s.bmax] { v.clear(); v.reserve(3 * n); }
        s.rot.clear(); s.mass.clear(); s.rest.clear(); s.stat.clear(); s.ids.clear();
        self.entities.clear();
        // gather: the same per-entity query the reference issues (update_systems.rs:381-400, :414-475)
        for &e in entities {
            let q = ComponentQuery::new()
                .read::<RigidBody<AABBCollisionBox>>(vec![e]).read::<Pos3>(vec![e]).read::<Scale>(vec![e]);
            let Some(res) = world.acquire_query(q) else { continue };
            let (Some(rb), Some(p3)) = (res.get::<RigidBody<AABBCollisionBox>>(e), res.get::<Pos3>(e)) else { continue };
            let rb = rb.read().map_err(|e| PhysicsCudaError(format!("Failed to read RigidBody: {e}")))?;
            let p3 = p3.read().map_err(|e| PhysicsCudaError(format!("Failed to read Pos3: {e}")))?;
            let sc = res.get::<Scale>(e).and_then(|s| s.read().ok().map(|g| match *g {
                Scale::Uniform(u) => [u, u, u],
                Scale::NonUniform { x, y, z } => [x, y, z],
            })).unwrap_or([1.0, 1.0, 1.0]);
            s.pos.extend_from_slice(&[p3.pos.x, p3.pos.y, p3.pos.z]);
            s.rot.extend_from_slice(&[p3.rot.v.x, p3.rot.v.y, p3.rot.v.z, p3.rot.s]); // cgmath order
            s.scale.extend_from_slice(&sc);
            s.vel.extend_from_slice(&[rb.velocity.x, rb.velocity.y, rb.velocity.z]);
            s.acc.extend_from_slice(&[rb.acceleration.x, rb.acceleration.y, rb.acceleration.z]);
            s.bmin.extend_from_slice(&[rb.collision_box.min.x, rb.collision_box.min.y, rb.collision_box.min.z]);
            s.bmax.extend_from_slice(&[rb.collision_box.max.x, rb.collision_box.max.y, rb.collision_box.max.z]);
            s.mass.push(rb.mass); s.rest.push(rb.restitution); s.stat.push(rb.is_static as u8); s.ids.push(*e);
            self.entities.push(e); // body i of the upload == self.entities[i]: skipped entities leave no hole
        }
        let n = s.ids.len() as u32;
        // Full mirror every frame: any user system may have written any component.  Large worlds
        // should call gp_upload once and gp_update_dirty / gp_set_state afterwards.
        let st = unsafe {
            sys::gp_upload(self.handle, n, s.pos.as_ptr(), s.rot.as_ptr(), s.scale.as_ptr(), s.vel.as_ptr(),
                s.acc.as_ptr(), s.bmin.as_ptr(), s.bmax.as_ptr(), s.mass.as_ptr(), s.rest.as_ptr(), s.stat.as_ptr(),
                s.ids.as_ptr(), std::ptr::null() /* rank = index in `entities` */)
        };
        self.check(st)?;
        // f32::exp is the platform expf: hand the library the very bits Rust would compute (physics.rs:437)
        let dts = dt.as_secs_f32();
        let damp = [(-1.6f32 * dts).exp(), (-1.8f32 * dts).exp(), (-3.5f32 * dts).exp()];
        let st = unsafe { sys::gp_step(self.handle, dts, damp.as_ptr(), std::ptr::null_mut()) };
        self.check(st)?;
        let st = unsafe { sys::gp_download(self.handle, s.pos.as_mut_ptr(), s.vel.as_mut_ptr()) };
        self.check(st)?;
        // scatter back through the same write locks the reference takes; iterate the GATHERED entities, not the
        // caller's list: after a skipped entity the two no longer line up
        debug_assert_eq!(self.entities.len(), n as usize);
        for (i, &e) in self.entities.iter().enumerate() {
            let q = ComponentQuery::new().write::<RigidBody<AABBCollisionBox>>(vec![e]).write::<Pos3>(vec![e]);
            let Some(res) = world.acquire_query(q) else { continue };
            if let (Some(rb), Some(p3)) = (res.get::<RigidBody<AABBCollisionBox>>(e), res.get::<Pos3>(e)) {
                let mut rb = rb.write().map_err(|e| PhysicsCudaError(format!("Failed to write RigidBody: {e}")))?;
                let mut p3 = p3.write().map_err(|e| PhysicsCudaError(format!("Failed to write Pos3: {e}")))?;
                p3.pos = cgmath::Vector3::new(s.pos[3 * i], s.pos[3 * i + 1], s.pos[3 * i + 2]);
                rb.velocity = cgmath::Vector3::new(s.vel[3 * i], s.vel[3 * i + 1], s.vel[3 * i + 2]);
            }
        }
        Ok(())
    }

    fn check(&self, st: sys::gp_status) -> Result<(), PhysicsCudaError> {
        if st < 0 { Err(PhysicsCudaError(last_error(self.handle))) } else { Ok(()) }
    }
}

impl Drop for CudaPhysics {
    fn drop(&mut self) {
        unsafe { sys::gp_destroy(self.handle) }
    }
}

fn last_error(h: *const sys::gp_world) -> String {
    unsafe { CStr::from_ptr(sys::gp_last_error(h)) }.to_string_lossy().into_owned()
}

/// Process-wide instance used by the drop-in system (one step in flight at a time, like the reference's pass 2).
pub fn shared() -> &'static Mutex<Option<CudaPhysics>> {
    static CELL: Mutex<Option<CudaPhysics>> = Mutex::new(None);
    &CELL
}

/// Passes 1 + 2 of `update_physics` on the GPU.  gears-app calls this in place of its two loops and keeps pass 3.
pub fn physics_passes_1_and_2(world: &Arc<World>, dt: Duration) -> Result<(), PhysicsCudaError> {
    let entities = world.get_entities_with_component::<RigidBody<AABBCollisionBox>>();
    if entities.is_empty() {
        return Ok(());
    }
    let mut guard = shared().lock().map_err(|e| PhysicsCudaError(format!("poisoned: {e}")))?;
    if guard.is_none() {
        *guard = Some(CudaPhysics::new(0)?);
    }
    guard.as_mut().unwrap().step(world, &entities, dt)
}

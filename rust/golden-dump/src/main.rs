//! See Cargo.toml.  Everything printed here comes out of the reference's own functions:
//! `RigidBody::update_pos` (physics.rs:405-462), `RigidBody::check_and_resolve_collision` (:474-499),
//! `CollisionBox::intersects` (:148-165, which is the only public door to `get_rotated_aabb`, :84-135) and
//! `Instance::to_raw` (gears-renderer instance.rs:22-31).  NOT compiled in this repository's environment.
use cgmath::{Quaternion, Vector3};
use gears_ecs::components::physics::{AABBCollisionBox, CollisionBox, RigidBody};
use gears_ecs::components::transforms::Pos3;
use gears_renderer::instance::Instance;

type V = Vector3<f32>;
fn v(x: f32, y: f32, z: f32) -> V { Vector3::new(x, y, z) }
fn hex(x: f32) -> String { format!("\"0x{:08x}\"", x.to_bits()) }
fn hex3(a: V) -> String { format!("[{}, {}, {}]", hex(a.x), hex(a.y), hex(a.z)) }
fn cube(h: V) -> AABBCollisionBox { AABBCollisionBox { min: -h, max: h } }

/// f32 <-> integers that order like the floats (for bisection over bit patterns)
fn ord(x: f32) -> i64 { let b = x.to_bits() as i32; (if b < 0 { i32::MIN - b } else { b }) as i64 }
fn unord(k: i64) -> f32 { let k = k as i32; f32::from_bits((if k < 0 { i32::MIN - k } else { k }) as u32) }

/// One face of the world AABB of (box, pos3, scale), found through `intersects` alone: the probe is an unrotated box
/// at the origin (its world AABB is its own min / max), wide open on the other two axes; on `axis` it either starts
/// at e (upper face: intersects iff e < max) or ends at e (lower face: intersects iff e > min).
fn face(b: &AABBCollisionBox, p: &Pos3, s: Option<&V>, axis: usize, upper: bool) -> f32 {
    let origin = Pos3::new(v(0.0, 0.0, 0.0));
    let hits = |e: f32| {
        let big = 1.0e6f32;
        let (mut lo, mut hi) = ([-big; 3], [big; 3]);
        if upper { lo[axis] = e; hi[axis] = e + 4.0e6 } else { hi[axis] = e; lo[axis] = e - 4.0e6 }
        let probe = AABBCollisionBox { min: v(lo[0], lo[1], lo[2]), max: v(hi[0], hi[1], hi[2]) };
        CollisionBox::intersects(b, p, s, &probe, &origin, None)
    };
    // upper: hits(e) is true for e < face, false from the face on; lower: false up to the face, true above it
    let (mut a, mut z) = (ord(-1.0e5), ord(1.0e5));
    while z - a > 1 {
        let m = (a + z) / 2;
        if hits(unord(m)) == upper { a = m } else { z = m }
    }
    if upper { unord(z) } else { unord(a) }
}

fn rotation_case(name: &str, q: [u32; 4]) -> String {
    let rot = Quaternion::new(f32::from_bits(q[3]), f32::from_bits(q[0]), f32::from_bits(q[1]), f32::from_bits(q[2]));
    let b = AABBCollisionBox { min: v(-1.5, -1.0, -1.5), max: v(1.5, 1.0, 10.0) };
    let p = Pos3::new_with_rot(v(14.0, 1.0, 38.0), rot);
    let s = v(3.0, 3.0, 3.0);
    let lo = v(face(&b, &p, Some(&s), 0, false), face(&b, &p, Some(&s), 1, false), face(&b, &p, Some(&s), 2, false));
    let hi = v(face(&b, &p, Some(&s), 0, true), face(&b, &p, Some(&s), 1, true), face(&b, &p, Some(&s), 2, true));
    let raw = Instance { position: p.pos, rotation: rot, scale: s }.to_raw();
    let model: Vec<String> = raw.model.iter().flatten().map(|x| hex(*x)).collect();
    let normal: Vec<String> = raw.normal.iter().flatten().map(|x| hex(*x)).collect();
    format!("\"{name}\": {{\"world_aabb_min_bits\": {}, \"world_aabb_max_bits\": {}, \"model_bits\": [{}], \"normal_bits\": [{}]}}",
            hex3(lo), hex3(hi), model.join(", "), normal.join(", "))
}

fn integrate(pos: V, vel: V, acc: V, dt: f32) -> (V, V) {
    let mut rb = RigidBody::new(1.0, vel, acc, cube(v(0.5, 0.5, 0.5)));
    let mut p = Pos3::new(pos);
    rb.update_pos(&mut p, dt);
    (p.pos, rb.velocity)
}

struct B { pos: V, vel: V, mass: f32, rest: f32, stat: bool, half: V }
fn resolve(a: &B, b: &B) -> (V, V, V, V) {
    let mk = |x: &B| {
        let mut rb = if x.stat { RigidBody::new_static(cube(x.half)) } else { RigidBody::new(x.mass, x.vel, v(0.0, 0.0, 0.0), cube(x.half)) };
        rb.restitution = x.rest;
        (rb, Pos3::new(x.pos))
    };
    let ((mut ra, mut pa), (mut rb, mut pb)) = (mk(a), mk(b));
    RigidBody::check_and_resolve_collision(&mut ra, &mut pa, None, &mut rb, &mut pb, None);
    (pa.pos, ra.velocity, pb.pos, rb.velocity)
}

fn main() {
    let mut out = Vec::new();
    for (name, pos, vel, acc, dt) in [("KA1", v(0.0, 10.0, 0.0), v(0.0, 0.0, 0.0), v(0.0, 0.0, 0.0), 0.016f32),
                                      ("KA2", v(0.0, 10.0, 0.0), v(0.0, -5.0, 0.0), v(0.0, 0.0, 0.0), 0.016),
                                      ("KA3", v(0.0, 0.0, 0.0), v(0.0, 0.0, 0.0), v(10.0, 0.0, 0.0), 0.1)] {
        let (p, w) = integrate(pos, vel, acc, dt);
        out.push(format!("\"{name}\": {{\"pos_bits\": {}, \"vel_bits\": {}}}", hex3(p), hex3(w)));
    }
    let dt = f32::from_bits(0x3c888889);
    out.push(format!("\"KD\": {{\"damp_bits\": [{}, {}, {}]}}", hex((-1.6f32 * dt).exp()), hex((-1.8f32 * dt).exp()), hex((-3.5f32 * dt).exp())));
    let st = |pos: V, half: V, rest: f32| B { pos, vel: v(0.0, 0.0, 0.0), mass: 0.0, rest, stat: true, half };
    let dy = |pos: V, vel: V, mass: f32, rest: f32, half: V| B { pos, vel, mass, rest, stat: false, half };
    let one = v(1.0, 1.0, 1.0);
    let hf = v(0.5, 0.5, 0.5);
    let slab = v(10.0, 0.5, 10.0);
    let cases = [
        ("KR1", dy(v(1.5, 0.0, 0.0), v(-5.0, 0.0, 0.0), 1.0, 0.2, one), st(v(0.0, 0.0, 0.0), one, 0.2)),
        ("KR2", dy(v(-0.2, 0.0, 0.0), v(1.0, 0.0, 0.0), 1.0, 0.2, hf), dy(v(0.2, 0.0, 0.0), v(-1.0, 0.0, 0.0), 1.0, 0.2, hf)),
        ("KR3", dy(v(0.0, 0.8, 0.0), v(0.0, -10.0, 0.0), 1.0, 0.8, hf), st(v(0.0, 0.0, 0.0), slab, 0.8)),
        ("KR4", dy(v(0.8, 0.0, 0.0), v(-1.0, 0.0, 0.0), 1.0, 0.2, one), st(v(0.0, 0.0, 0.0), one, 0.2)),
        ("KR5", dy(v(0.0, 0.9, 0.0), v(3.0, -2.0, 0.0), 2.0, 0.2, hf), st(v(0.0, 0.0, 0.0), slab, 0.2)),
    ];
    for (name, a, b) in cases.iter() {
        let (pa, va, pb, vb) = resolve(a, b);
        out.push(format!("\"{name}\": {{\"a_pos_bits\": {}, \"a_vel_bits\": {}, \"b_pos_bits\": {}, \"b_vel_bits\": {}}}",
                         hex3(pa), hex3(va), hex3(pb), hex3(vb)));
    }
    out.push(format!("\"rotation\": {{{}, {}}}",
                     rotation_case("quarter_turn_y", [0x00000000, 0x3F3504F3, 0x00000000, 0x3F3504F3]),
                     rotation_case("generic", [0x3E3AF4BA, 0x3EBAF4BA, 0x3F0C378B, 0x3F3AF4BA])));
    println!("{{\n {}\n}}", out.join(",\n "));
}

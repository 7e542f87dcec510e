// gears_physics.hpp — C++ host side above the C ABI: a mirror of the reference's ECS-facing physics surface.
//
// The reference is Rust; no Rust toolchain exists in this build environment, so the compiled-language host side is
// written in C++ (the Rust crates under rust/ carry the same logic for a maintainer to build).  Names, argument
// meaning and error behaviour follow the reference:
//   Pos3, Scale                         crates/gears-ecs/src/components/transforms.rs:17-22, 83-86
//   AABBCollisionBox, RigidBody         crates/gears-ecs/src/components/physics.rs:66-72, 302-388
//   World (the slice of it the step uses: create_entity, add_component, get_entities_with_component, get)
//                                       crates/gears-ecs/src/lib.rs:184-470
//   update_physics(world, dt)           crates/gears-app/src/systems/update_systems.rs:362-487 (passes 1 and 2)
//   SystemResult / SystemError          crates/gears-app/src/systems/errors.rs:5-27
// Nothing here computes physics: update_physics mirrors the components into libgears_physics_cuda (gp_upload /
// gp_set_state), runs gp_step on the B200 and writes positions and velocities back.  No CPU fallback.
#pragma once
#include <chrono>
#include <cmath>
#include <cstdint>
#include <optional>
#include <string>
#include <unordered_map>
#include <variant>
#include <vector>

#include "../../include/gears_physics_cuda.h"

namespace gears {

struct Vector3 {
    float x = 0, y = 0, z = 0;
    bool operator==(const Vector3& o) const { return x == o.x && y == o.y && z == o.z; }
};
struct Quaternion {  // cgmath layout: vector part, then scalar
    Vector3 v;
    float s = 1;
    static Quaternion one() { return Quaternion{}; }
};

struct Pos3 {
    Vector3 pos;
    Quaternion rot = Quaternion::one();
    static Pos3 new_(Vector3 p) { return Pos3{p, Quaternion::one()}; }
    static Pos3 new_with_rot(Vector3 p, Quaternion r) { return Pos3{p, r}; }
};

struct ScaleUniform {
    float s;
};
struct ScaleNonUniform {
    float x, y, z;
};
using Scale = std::variant<ScaleUniform, ScaleNonUniform>;

struct AABBCollisionBox {
    Vector3 min{-0.5f, -0.5f, -0.5f}, max{0.5f, 0.5f, 0.5f};
};

constexpr float DEFAULT_RESTITUTION = 0.2f;  // physics.rs:13

struct RigidBody {
    float mass = 1.0f;
    Vector3 velocity, acceleration;
    AABBCollisionBox collision_box;
    bool is_static = false;
    float restitution = DEFAULT_RESTITUTION;
    static RigidBody new_(float mass, Vector3 velocity, Vector3 acceleration, AABBCollisionBox box) {
        return RigidBody{mass, velocity, acceleration, box, false, DEFAULT_RESTITUTION};
    }
    static RigidBody new_static(AABBCollisionBox box) { return RigidBody{0.0f, {}, {}, box, true, DEFAULT_RESTITUTION}; }
};

using Entity = uint32_t;

// The part of gears_ecs::World the physics system touches.  Iteration order of the RigidBody storage is the
// insertion order here (the reference's DashMap order is arbitrary but fixed for an unchanged map; the step takes
// whatever order get_entities_with_component returns).
class World {
public:
    Entity create_entity() { return next_++; }
    void add_component(Entity e, const Pos3& c) { pos3_[e] = c; }
    void add_component(Entity e, const Scale& c) { scale_[e] = c; }
    void add_component(Entity e, const RigidBody& c) {
        if (!body_.count(e)) order_.push_back(e), ++generation_;
        body_[e] = c;
    }
    const std::vector<Entity>& get_entities_with_rigid_body() const { return order_; }
    Pos3* get_pos3(Entity e) { auto it = pos3_.find(e); return it == pos3_.end() ? nullptr : &it->second; }
    RigidBody* get_rigid_body(Entity e) { auto it = body_.find(e); return it == body_.end() ? nullptr : &it->second; }
    const Scale* get_scale(Entity e) const { auto it = scale_.find(e); return it == scale_.end() ? nullptr : &it->second; }
    uint64_t generation() const { return generation_; }

private:
    Entity next_ = 0;
    uint64_t generation_ = 0;
    std::vector<Entity> order_;
    std::unordered_map<Entity, Pos3> pos3_;
    std::unordered_map<Entity, Scale> scale_;
    std::unordered_map<Entity, RigidBody> body_;
};

// SystemResult<()> : empty optional = Ok(()), a string = SystemError::Execution(msg)
using SystemResult = std::optional<std::string>;

// The drop-in system.  Holds the device world; one step in flight at a time (like the reference's pass 2).
class CudaPhysics {
public:
    explicit CudaPhysics(int device = 0) {
        gp_config cfg{};
        cfg.struct_size = sizeof cfg;
        cfg.device = device;
        if (gp_create(&cfg, &h_) < 0) err_ = gp_last_error(nullptr);
    }
    ~CudaPhysics() {
        if (h_) gp_destroy(h_);
    }
    CudaPhysics(const CudaPhysics&) = delete;
    CudaPhysics& operator=(const CudaPhysics&) = delete;

    // update_physics passes 1 + 2 (update_systems.rs:377-487).  Entities without a Pos3 are skipped, as the
    // reference's `if let (Some(body), Some(pos3))` does (:383-388, :421-432).
    SystemResult update_physics(World& world, std::chrono::duration<float> dt) {
        if (!h_) return "gears-physics-cuda: " + err_;
        const auto& ents = world.get_entities_with_rigid_body();
        if (ents.empty()) return std::nullopt;  // :373-375
        gather(world, ents);
        const uint32_t n = (uint32_t)ids_.size();
        // full mirror every frame: any user system may have written any component between frames
        if (gp_upload(h_, n, pos_.data(), rot_.data(), scale_.data(), vel_.data(), acc_.data(), bmin_.data(), bmax_.data(),
                      mass_.data(), rest_.data(), stat_.data(), ids_.data(), nullptr) < 0)
            return std::string(gp_last_error(h_));
        const float dts = dt.count();  // dt.as_secs_f32()
        const float damp[3] = {std::exp(-1.6f * dts), std::exp(-1.8f * dts), std::exp(-3.5f * dts)};  // f32::exp
        if (gp_step(h_, dts, damp, &stats_) < 0) return std::string(gp_last_error(h_));
        if (gp_download(h_, pos_.data(), vel_.data()) < 0) return std::string(gp_last_error(h_));
        for (uint32_t i = 0; i < n; ++i) {
            Pos3* p = world.get_pos3(ids_[i]);
            RigidBody* b = world.get_rigid_body(ids_[i]);
            p->pos = Vector3{pos_[3 * i], pos_[3 * i + 1], pos_[3 * i + 2]};
            b->velocity = Vector3{vel_[3 * i], vel_[3 * i + 1], vel_[3 * i + 2]};
        }
        return std::nullopt;
    }
    const gp_step_stats& last_stats() const { return stats_; }

private:
    void gather(World& world, const std::vector<Entity>& ents) {
        pos_.clear(), rot_.clear(), scale_.clear(), vel_.clear(), acc_.clear(), bmin_.clear(), bmax_.clear();
        mass_.clear(), rest_.clear(), stat_.clear(), ids_.clear();
        for (Entity e : ents) {
            Pos3* p = world.get_pos3(e);
            RigidBody* b = world.get_rigid_body(e);
            if (!p || !b) continue;
            float sc[3] = {1, 1, 1};
            if (const Scale* s = world.get_scale(e)) {  // update_systems.rs:456-475
                if (auto u = std::get_if<ScaleUniform>(s)) sc[0] = sc[1] = sc[2] = u->s;
                if (auto v = std::get_if<ScaleNonUniform>(s)) sc[0] = v->x, sc[1] = v->y, sc[2] = v->z;
            }
            push3(pos_, p->pos), push3(vel_, b->velocity), push3(acc_, b->acceleration);
            push3(bmin_, b->collision_box.min), push3(bmax_, b->collision_box.max);
            rot_.insert(rot_.end(), {p->rot.v.x, p->rot.v.y, p->rot.v.z, p->rot.s});
            scale_.insert(scale_.end(), sc, sc + 3);
            mass_.push_back(b->mass), rest_.push_back(b->restitution), stat_.push_back(b->is_static ? 1 : 0);
            ids_.push_back(e);
        }
    }
    static void push3(std::vector<float>& v, const Vector3& a) { v.insert(v.end(), {a.x, a.y, a.z}); }

    gp_world* h_ = nullptr;
    std::string err_;
    gp_step_stats stats_{};
    std::vector<float> pos_, rot_, scale_, vel_, acc_, bmin_, bmax_, mass_, rest_;
    std::vector<uint8_t> stat_;
    std::vector<uint32_t> ids_;
};

}  // namespace gears

// System-level tests of the C++ host mirror, written like the reference's own unit tests
// (crates/gears-ecs/src/components/physics.rs:696-920) but going through update_physics on the GPU.
// Exit code 0 = all passed.  Built and run by tests/test_gpu_host_cpp.py on the B200 box.
#include <cstdio>
#include <cstring>

#include "../../gears_b200/host/gears_physics.hpp"

using namespace gears;
using secs = std::chrono::duration<float>;

static int failures = 0;
#define CHECK(cond)                                                        \
    do {                                                                   \
        if (!(cond)) {                                                     \
            std::printf("FAILED %s:%d  %s\n", __FILE__, __LINE__, #cond); \
            ++failures;                                                    \
        }                                                                  \
    } while (0)

static uint32_t bits(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}
static AABBCollisionBox create_box(Vector3 mn, Vector3 mx) { return AABBCollisionBox{mn, mx}; }

int main() {
    CudaPhysics phys(0);

    {  // test_update_pos_applies_gravity (physics.rs:696-708) + known-answer bits
        World w;
        Entity e = w.create_entity();
        w.add_component(e, Pos3::new_(Vector3{0, 10, 0}));
        w.add_component(e, RigidBody::new_(1.0f, {}, {}, create_box({-0.5f, -0.5f, -0.5f}, {0.5f, 0.5f, 0.5f})));
        auto r = phys.update_physics(w, secs(0.016f));
        CHECK(!r.has_value());
        CHECK(w.get_rigid_body(e)->velocity.y < 0.0f);
        CHECK(w.get_pos3(e)->pos.y < 10.0f);
        CHECK(bits(w.get_rigid_body(e)->velocity.y) == 0xbed8e24du);
        CHECK(bits(w.get_pos3(e)->pos.y) == 0x411fe43du);
    }
    {  // test_update_pos_static_body_doesnt_move (:711-722)
        World w;
        Entity e = w.create_entity();
        w.add_component(e, Pos3::new_(Vector3{0, 5, 0}));
        w.add_component(e, RigidBody::new_static(create_box({-1, -1, -1}, {1, 1, 1})));
        CHECK(!phys.update_physics(w, secs(0.016f)).has_value());
        CHECK((w.get_pos3(e)->pos == Vector3{0, 5, 0}));
    }
    {  // test_collision_resolution_separates_objects (:767-800) at system level: dt = 0 keeps pass 1 a no-op in x
        World w;
        Entity a = w.create_entity(), b = w.create_entity();
        w.add_component(a, Pos3::new_(Vector3{0.8f, 0, 0}));
        w.add_component(a, RigidBody::new_(1.0f, Vector3{-1, 0, 0}, {}, create_box({-1, -1, -1}, {1, 1, 1})));
        w.add_component(b, Pos3::new_(Vector3{0, 0, 0}));
        w.add_component(b, RigidBody::new_static(create_box({-1, -1, -1}, {1, 1, 1})));
        CHECK(!phys.update_physics(w, secs(0.0f)).has_value());
        CHECK(w.get_pos3(a)->pos.x > 0.8f);                 // A pushed out
        CHECK((w.get_pos3(b)->pos == Vector3{0, 0, 0}));    // static B unmoved
        CHECK(bits(w.get_pos3(a)->pos.x) == 0x3fa353f8u);   // KR4
        CHECK(phys.last_stats().n_contacts == 1);
    }
    {  // test_collision_detection_touching (:632-642): exactly touching boxes do not collide
        World w;
        Entity a = w.create_entity(), b = w.create_entity();
        w.add_component(a, Pos3::new_(Vector3{0, 0, 0}));
        w.add_component(a, RigidBody::new_static(create_box({-1, -1, -1}, {1, 1, 1})));
        w.add_component(b, Pos3::new_(Vector3{2, 0, 0}));
        w.add_component(b, RigidBody::new_(1.0f, {}, {}, create_box({-1, -1, -1}, {1, 1, 1})));
        CHECK(!phys.update_physics(w, secs(0.0f)).has_value());
        CHECK(phys.last_stats().n_contacts == 0 && phys.last_stats().n_snapshot == 0);
    }
    {  // test_collision_detection_with_scale (:645-661): scale applied through the Scale component
        World w;
        Entity a = w.create_entity(), b = w.create_entity();
        w.add_component(a, Pos3::new_(Vector3{0, 0, 0}));
        w.add_component(a, RigidBody::new_static(create_box({-1, -1, -1}, {1, 1, 1})));
        w.add_component(a, Scale{ScaleUniform{2.0f}});
        w.add_component(b, Pos3::new_(Vector3{2.5f, 0, 0}));
        w.add_component(b, RigidBody::new_(1.0f, {}, {}, create_box({-1, -1, -1}, {1, 1, 1})));
        CHECK(!phys.update_physics(w, secs(0.0f)).has_value());
        CHECK(phys.last_stats().n_snapshot == 1);  // 2x scaled A reaches x = 2 > 1.5
    }
    {  // an entity without Pos3 is skipped silently (update_systems.rs:383-388), an empty world is Ok(())
        World w;
        CHECK(!phys.update_physics(w, secs(0.016f)).has_value());
        Entity a = w.create_entity(), b = w.create_entity();
        w.add_component(a, RigidBody::new_(1.0f, {}, {}, AABBCollisionBox{}));
        w.add_component(b, Pos3::new_(Vector3{0, 3, 0}));
        w.add_component(b, RigidBody::new_(2.0f, {}, {}, AABBCollisionBox{}));
        CHECK(!phys.update_physics(w, secs(0.016f)).has_value());
        CHECK(w.get_pos3(b)->pos.y < 3.0f);
        CHECK(phys.last_stats().n_bodies == 1);
    }
    if (failures == 0) std::printf("HOST_CPP_OK\n");
    return failures ? 1 : 0;
}

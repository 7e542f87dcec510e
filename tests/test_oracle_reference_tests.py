"""Ports of the reference's own physics unit tests (crates/gears-ecs/src/components/physics.rs:558-997)
run against the oracle.  Each test names the reference test it mirrors and keeps its inputs and asserts."""
import numpy as np

BOX1 = dict(box_min=[-1, -1, -1], box_max=[1, 1, 1])
BOXH = dict(box_min=[-0.5, -0.5, -0.5], box_max=[0.5, 0.5, 0.5])


def rb(pos, vel=(0, 0, 0), mass=1.0, static=False, rest=0.2, box=BOXH):
    return dict(pos=pos, vel=vel, mass=0.0 if static else mass, restitution=rest, is_static=static, **box)


# physics.rs:606-616
def test_collision_detection_intersecting(oracle):
    assert oracle.intersects([-1] * 3, [1] * 3, [0, 0, 0], [-1] * 3, [1] * 3, [1.5, 0, 0])


# physics.rs:619-629
def test_collision_detection_not_intersecting(oracle):
    assert not oracle.intersects([-1] * 3, [1] * 3, [0, 0, 0], [-1] * 3, [1] * 3, [5, 0, 0])


# physics.rs:632-642 — pins the strict inequalities
def test_collision_detection_touching(oracle):
    assert not oracle.intersects([-1] * 3, [1] * 3, [0, 0, 0], [-1] * 3, [1] * 3, [2, 0, 0])


# physics.rs:645-661
def test_collision_detection_with_scale(oracle):
    assert oracle.intersects([-1] * 3, [1] * 3, [0, 0, 0], [-1] * 3, [1] * 3, [2.5, 0, 0], a_scale=[2, 1, 1])
    assert not oracle.intersects([-1] * 3, [1] * 3, [0, 0, 0], [-1] * 3, [1] * 3, [2.5, 0, 0])


# physics.rs:664-678
def test_velocity_capping_horizontal(oracle):
    v = oracle.cap_velocity([30, 0, 30])
    assert np.sqrt(v[0] ** 2 + v[2] ** 2) <= 20.0 + 0.001


# physics.rs:681-693
def test_velocity_capping_vertical(oracle):
    v = oracle.cap_velocity([0, 50, 0])
    assert -40.0 <= v[1] <= 40.0


# physics.rs:696-708
def test_update_pos_applies_gravity(oracle):
    p, v = oracle.update_pos([0, 10, 0], [0, 0, 0], [0, 0, 0], False, 0.016)
    assert v[1] < 0 and p[1] < 10


# physics.rs:711-722
def test_update_pos_static_body_doesnt_move(oracle):
    p, v = oracle.update_pos([0, 5, 0], [0, 0, 0], [0, 0, 0], True, 0.016)
    assert np.array_equal(p, np.array([0, 5, 0], np.float32))


# physics.rs:725-742
def test_update_pos_applies_acceleration(oracle):
    p, v = oracle.update_pos([0, 0, 0], [0, 0, 0], [10, 0, 0], False, 0.1)
    assert v[0] > 0 and p[0] > 0


# physics.rs:745-764
def test_update_pos_applies_damping(oracle):
    p, v = oracle.update_pos([0, 0, 0], [10, 0, 0], [0, 0, 0], False, 0.1)
    assert v[0] < 10


# physics.rs:767-800
def test_collision_resolution_separates_objects(oracle):
    a = rb([0.8, 0, 0], vel=[-1, 0, 0], box=BOX1)
    b = rb([0, 0, 0], static=True, box=BOX1)
    hit, ap, av, bp, bv = oracle.check_and_resolve(a, b)
    assert hit and ap[0] > 0.8
    assert np.array_equal(bp, np.zeros(3, np.float32))


# physics.rs:803-829
def test_collision_resolution_reverses_velocity(oracle):
    a = rb([1.5, 0, 0], vel=[-5, 0, 0], box=BOX1)
    b = rb([0, 0, 0], static=True, box=BOX1)
    _, ap, av, _, _ = oracle.check_and_resolve(a, b)
    assert av[0] >= -5.0


# physics.rs:832-857
def test_collision_resolution_both_dynamic(oracle):
    a = rb([-0.2, 0, 0], vel=[1, 0, 0])
    b = rb([0.2, 0, 0], vel=[-1, 0, 0])
    _, ap, av, bp, bv = oracle.check_and_resolve(a, b)
    assert ap[0] < -0.2 and bp[0] > 0.2


# physics.rs:860-892
def test_restitution_affects_bounce(oracle):
    a = rb([0, 0.8, 0], vel=[0, -10, 0], rest=0.8)
    b = rb([0, 0, 0], static=True, rest=0.8, box=dict(box_min=[-10, -0.5, -10], box_max=[10, 0.5, 10]))
    _, ap, av, _, _ = oracle.check_and_resolve(a, b)
    assert av[1] > 0


# physics.rs:895-920 — exact equality of unchanged state
def test_no_collision_when_separated(oracle):
    a = rb([0, 0, 0], vel=[1, 0, 0])
    b = rb([10, 0, 0], vel=[-1, 0, 0])
    hit, ap, av, bp, bv = oracle.check_and_resolve(a, b)
    assert not hit
    assert np.array_equal(ap, np.float32([0, 0, 0])) and np.array_equal(bp, np.float32([10, 0, 0]))
    assert np.array_equal(av, np.float32([1, 0, 0])) and np.array_equal(bv, np.float32([-1, 0, 0]))


# physics.rs:923-943 — rotated box; here we additionally check a 90 degree turn about z swaps x/y extents
def test_rotated_collision_box(oracle):
    h = np.float32(np.sqrt(0.5))
    lo, hi = oracle.world_aabb([-2, -1, -1], [2, 1, 1], [0, 0, 0], rot=[0, 0, h, h])
    assert np.allclose(lo, [-1, -2, -1], atol=1e-6) and np.allclose(hi, [1, 2, 1], atol=1e-6)
    lo, hi = oracle.world_aabb([-2, -1, -1], [2, 1, 1], [3, 4, 5])
    assert np.array_equal(lo, np.float32([1, 3, 4])) and np.array_equal(hi, np.float32([5, 5, 6]))


# physics.rs:955-974
def test_fall_multiplier_applies_when_falling(oracle):
    p, v = oracle.update_pos([0, 10, 0], [0, -5, 0], [0, 0, 0], False, 0.016)
    assert abs(v[1] - (-5.0)) > 28.0 * 0.016


# physics.rs:977-997
def test_min_velocity_threshold(oracle):
    p, v = oracle.update_pos([0, 0, 0], [0.001, 0, 0], [0, 0, 0], False, 0.016)
    assert abs(v[0]) < 0.001


def test_acceleration_y_is_never_integrated(oracle):
    # physics.rs:451-452 add only a.x, a.z
    _, v0 = oracle.update_pos([0, 0, 0], [0, 1, 0], [5, 0, 0], False, 0.016)
    _, v1 = oracle.update_pos([0, 0, 0], [0, 1, 0], [5, -10, 0], False, 0.016)
    assert v0[1] == v1[1]

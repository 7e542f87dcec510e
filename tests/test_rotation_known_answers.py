"""Known answers for the quaternion arithmetic of the hot path (cgmath 0.18 `Quaternion * Vector3`, physics.rs:115;
`Matrix3/4::from(Quaternion)`, instance.rs:23-26), derived WITHOUT the oracle: exact rationals, rounded to binary32
after every operation (tests/golden/derive_rotation_answers.py, trace in tests/golden/ROTATION_DERIVATION.md).
Shapes as the API is used with them: an asymmetric box, a quarter turn about y (examples/src/interactive/map.rs:325-346)
and Scale::Uniform(3.0) (map.rs:599-613).  The oracle is CHECKED against them on the CPU, the CUDA path on the GPU."""
import json
import os

import numpy as np
import pytest

from gears_b200 import scenes

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "rotation_answers.json")))["cases"]
DT = np.float32(1.0) / np.float32(60.0)


def f32(bits):
    return np.array([int(b, 16) for b in bits], np.uint32).view(np.float32)


def hexs(a):
    return ["0x%08x" % int(x) for x in np.ascontiguousarray(a, np.float32).view(np.uint32).reshape(-1)]


@pytest.mark.parametrize("name", sorted(G))
def test_oracle_world_aabb_and_instance(oracle, name):
    g = G[name]
    q = f32(g["quat_bits_v_xyz_s"])
    lo, hi = oracle.world_aabb(g["box_min"], g["box_max"], g["pos"], q, g["scale"])
    assert hexs(lo) == g["world_aabb_min_bits"] and hexs(hi) == g["world_aabb_max_bits"]
    r = oracle.instance_raw(g["pos"], q, g["scale"])
    assert hexs(r[:16]) == g["model_bits"] and hexs(r[16:]) == g["normal_bits"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(G))
def test_gpu_instance_raw(name):
    from gears_b200.world import PhysicsWorld
    g = G[name]
    q = f32(g["quat_bits_v_xyz_s"])
    s = scenes.subset(scenes.uniform(1, L=1.0, floor=False, rank_seed=None), np.arange(1))
    s["pos"][0] = g["pos"]
    s["is_static"][0] = 1
    with PhysicsWorld() as w:
        w.upload(s)
        r = w.instances(q[None, :], np.float32([g["scale"]]))[0]
    assert hexs(r[:16]) == g["model_bits"] and hexs(r[16:]) == g["normal_bits"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(G))
def test_gpu_world_aabb_edges(name):
    """The device's world AABB of the rotated, scaled, asymmetric box has the derived bits: a probe box that TOUCHES
    the derived max.x / min.z face is no snapshot pair (strict inequalities, physics.rs:159-164), one ulp further in it
    is."""
    from gears_b200.world import PhysicsWorld
    g = G[name]
    q = f32(g["quat_bits_v_xyz_s"])
    lo, hi = f32(g["world_aabb_min_bits"]), f32(g["world_aabb_max_bits"])
    cy, cz = np.float32((lo[1] + hi[1]) * 0.5), np.float32((lo[2] + hi[2]) * 0.5)
    cx = np.float32((lo[0] + hi[0]) * 0.5)
    for axis, touch_lo in ((0, hi[0]), (2, None)):
        for inside in (False, True):
            s = scenes.subset(scenes.uniform(2, L=1.0, floor=False, rank_seed=None), np.arange(2))
            s["is_static"][:] = (1, 0)
            s["mass"][0] = 0
            s["box_min"][0], s["box_max"][0], s["pos"][0] = g["box_min"], g["box_max"], g["pos"]
            s["rot"] = np.stack([q, np.float32([0, 0, 0, 1])])
            s["scale"] = np.float32([g["scale"], [1, 1, 1]])
            # probe: a dynamic box at pos = 0 (its world AABB is its box, no rounding) with zero velocity: pass 1 only
            # lets it fall a few millimetres (the target is 6 units tall), x and z stay exactly where they are
            s["pos"][1] = (0, 0, 0)
            s["vel"][1] = (0, 0, 0)
            if axis == 0:   # probe's min.x at (or one ulp below) the target's max.x
                edge = hi[0] if not inside else np.nextafter(hi[0], np.float32(-np.inf))
                s["box_min"][1] = (edge, cy - 1, cz - 1)
                s["box_max"][1] = (edge + np.float32(2), cy + 1, cz + 1)
            else:           # probe's max.z at (or one ulp above) the target's min.z
                edge = lo[2] if not inside else np.nextafter(lo[2], np.float32(np.inf))
                s["box_min"][1] = (cx - 1, cy - 1, edge - np.float32(2))
                s["box_max"][1] = (cx + 1, cy + 1, edge)
            with PhysicsWorld(record_pairs=True) as w:
                w.upload(s)
                w.step(DT)
                snap = w.pairs(0)
            assert (len(snap) == 1) == inside, (name, axis, inside, snap)

"""Next rows of the scope table: instance hand-off (N1), ray cast (N3), obstacle grid (N4).
CPU part: the oracle against hand-derived answers and the reference's own expectations; GPU part: bit-exact
parity of the CUDA kernels with the oracle through the C ABI."""
import numpy as np
import pytest

from gears_b200 import scenes

DT = np.float32(1.0) / np.float32(60.0)


# ---------------- oracle (CPU) ----------------
def test_oracle_instance_raw_identity_and_scale(oracle):
    r = oracle.instance_raw([1, 2, 3], [0, 0, 0, 1], [2, 3, 4])
    model = r[:16].reshape(4, 4).T  # column-major -> rows
    assert np.array_equal(model, np.float32([[2, 0, 0, 1], [0, 3, 0, 2], [0, 0, 4, 3], [0, 0, 0, 1]]))
    assert np.array_equal(r[16:].reshape(3, 3), np.eye(3, dtype=np.float32))


def test_oracle_instance_raw_quarter_turn_about_y(oracle):
    h = np.float32(np.sqrt(0.5))
    r = oracle.instance_raw([0, 0, 0], [0, h, 0, h], [1, 1, 1])
    normal = r[16:].reshape(3, 3).T
    # a rotation of +90 degrees about y maps x -> -z, z -> x
    assert np.allclose(normal @ np.float32([1, 0, 0]), [0, 0, -1], atol=1e-6)
    assert np.allclose(normal @ np.float32([0, 0, 1]), [1, 0, 0], atol=1e-6)


def test_oracle_ray_slab_cases(oracle):
    box = ([-1, -1, -1], [1, 1, 1])
    assert oracle.ray_intersects_aabb([0, 0, -5], [0, 0, 1], [0, 0, 0], *box)
    assert not oracle.ray_intersects_aabb([0, 0, -5], [0, 0, -1], [0, 0, 0], *box)      # behind the origin
    assert not oracle.ray_intersects_aabb([3, 0, -5], [0, 0, 1], [0, 0, 0], *box)       # parallel, outside the slab
    assert oracle.ray_intersects_aabb([1, 0, -5], [0, 0, 1], [0, 0, 0], *box)           # grazing the face counts
    assert oracle.ray_intersects_aabb([0, 0, 0], [1, 0, 0], [0, 0, 0], *box)            # origin inside
    assert oracle.ray_intersects_aabb([-5, -5, -5], [1, 1, 1], [0, 0, 0], *box)


def test_oracle_obstacle_cells_round_half_away_from_zero(oracle):
    lo, hi = oracle.obstacle_cells([0.4, 2.5, -2.5], [-1, -1, -1], [1, 1, 1], 1.0)
    assert list(lo) == [-1, 2, -4] and list(hi) == [1, 4, -2]
    lo, hi = oracle.obstacle_cells([10, 0, 0], [-3, -0.1, -3], [3, 0.1, 3], 2.0)
    assert list(lo) == [4, 0, -2] and list(hi) == [7, 0, 2]   # 3.5 -> 4, 6.5 -> 7, -1.5 -> -2


# ---------------- CUDA vs oracle (GPU) ----------------
def _world_after_steps(s, steps=3):
    from gears_b200.world import PhysicsWorld
    w = PhysicsWorld()
    w.upload(s)
    for _ in range(steps):
        w.step(DT)
    return w


@pytest.mark.gpu
def test_instances_match_oracle_bitwise(oracle):
    n = 3000
    s = scenes.uniform(n, L=25.0, seed=201, rank_seed=202)
    m = s["pos"].shape[0]
    rng = np.random.default_rng(7)
    q = rng.normal(size=(m, 4)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True).astype(np.float32)
    sc = (0.5 + rng.random((m, 3))).astype(np.float32)
    with _world_after_steps(s) as w:
        pos, _ = w.download()
        got = w.instances(q, sc)
        got_id = w.instances()
    for i in range(0, m, 7):
        assert np.array_equal(got[i].view(np.uint32), oracle.instance_raw(pos[i], q[i], sc[i]).view(np.uint32)), i
        assert np.array_equal(got_id[i].view(np.uint32),
                              oracle.instance_raw(pos[i], [0, 0, 0, 1], [1, 1, 1]).view(np.uint32)), i


@pytest.mark.gpu
def test_raycast_matches_oracle(oracle):
    s = scenes.uniform(400, L=12.0, seed=211, rank_seed=212)
    s["box_min"][:50] = (-0.2, -0.9, -0.4)   # asymmetric boxes: the reference uses the raw box, not the rotated one
    s["box_max"][:50] = (0.7, 0.3, 0.6)
    rng = np.random.default_rng(11)
    nr = 40
    org = (rng.random((nr, 3)) * 12).astype(np.float32)
    d = rng.normal(size=(nr, 3)).astype(np.float32)
    d[0] = (0, 0, 1)
    d[1] = (1, 0, 0)          # axis-parallel rays exercise the |dir| < 1e-8 branch
    d[2] = (0, -1, 0)
    with _world_after_steps(s) as w:
        pos, _ = w.download()
        hits = w.raycast(org, d)
        sub = np.array([3, 17, 399, 400], np.uint32)
        hits_sub = w.raycast(org, d, targets=sub)
    m = s["pos"].shape[0]
    ref = np.zeros((nr, m), bool)
    for r in range(nr):
        for t in range(m):
            ref[r, t] = oracle.ray_intersects_aabb(org[r], d[r], pos[t], s["box_min"][t], s["box_max"][t])
    assert ref.any() and not ref.all()
    assert np.array_equal(hits, ref)
    assert np.array_equal(hits_sub, ref[:, sub])


@pytest.mark.gpu
def test_obstacle_grid_matches_oracle(oracle):
    s = scenes.mixed(600, L=40.0, seed=221, rank_seed=222)
    cell = 1.5
    origin, dims = np.int32([-2, -2, -2]), np.uint32([34, 20, 34])   # clips part of the scene on purpose
    with _world_after_steps(s, steps=2) as w:
        pos, _ = w.download()
        statics = np.nonzero(s["is_static"])[0].astype(np.uint32)
        grid, bounds = w.obstacle_grid(cell, origin, dims, idx=statics)
        grid_all, _ = w.obstacle_grid(cell, origin, dims)
    ref = np.zeros((dims[2], dims[1], dims[0]), bool)
    blo, bhi = np.full(3, 2**31 - 1, np.int64), np.full(3, -2**31, np.int64)
    for i in statics:
        lo, hi = oracle.obstacle_cells(pos[i], s["box_min"][i], s["box_max"][i], cell)
        blo, bhi = np.minimum(blo, lo), np.maximum(bhi, hi)
        l = np.maximum(lo.astype(np.int64) - origin, 0)
        h = np.minimum(hi.astype(np.int64) - origin, dims.astype(np.int64) - 1)
        if np.all(l <= h):
            ref[l[2]:h[2] + 1, l[1]:h[1] + 1, l[0]:h[0] + 1] = True
    assert ref.any()
    assert np.array_equal(grid, ref)
    assert np.array_equal(bounds, np.concatenate([blo, bhi]).astype(np.int32))
    assert grid_all.sum() >= grid.sum() and np.all(grid_all[grid])


def world(**kw):
    from gears_b200.world import PhysicsWorld
    return PhysicsWorld(**kw)


def assert_same_bits(a, b, what):
    a = np.ascontiguousarray(a, np.float32).view(np.uint32)
    b = np.ascontiguousarray(b, np.float32).view(np.uint32)
    assert np.array_equal(a, b), f"{what}: {int((a != b).sum())} words differ"


@pytest.mark.gpu
def test_pipelined_frames_equal_set_state_step_download(oracle):
    """gp_frame_submit / gp_frame_wait (N2: async H2D / D2H overlapped with the step): four frames in flight back to
    back give the very bits of gp_set_state + gp_step + gp_download called frame by frame, and of the oracle."""
    s = scenes.uniform(30_000, L=70.0, seed=61, rank_seed=62)
    n = len(s["mass"])
    rng = np.random.default_rng(5)
    frames = []
    with world() as w1:
        w1.upload(s)
        pos, vel = s["pos"].copy(), s["vel"].copy()
        for t in range(4):
            acc = (rng.random((n, 3)).astype(np.float32) - np.float32(0.5)) * np.float32(4.0) * (t % 2)
            w1.set_state(pos, vel, acc)       # what the host hands over for this frame
            w1.step(DT)
            p1, v1 = w1.download()
            r = oracle.step(dict(s, pos=pos, vel=vel, acc=acc), DT, 1, mode=1, want_snapshot=False)
            assert_same_bits(p1, r["pos"], f"frame {t} pos (per-call path vs oracle)")
            frames.append((pos, vel, acc, p1, v1))
            pos, vel = p1.copy(), v1.copy()
    with world() as w2:
        w2.upload(s)
        w2.step(DT)  # the frame path must not depend on what state the world happens to be in
        ins, outs = [], []
        for (pos, vel, acc, _, _) in frames:
            hp, hv, ha = (w2.host_alloc((n, 3)) for _ in range(3))
            hp[:], hv[:], ha[:] = pos, vel, acc
            o = w2.host_alloc((n, 6))
            o[:] = np.float32(7.0)
            w2.frame_submit(hp, hv, ha, DT, o)
            ins.append((hp, hv, ha)), outs.append(o)
        st = w2.frame_wait(0)
        assert st["steps_inexact"] == 0
        for t, (o, (_, _, _, p1, v1)) in enumerate(zip(outs, frames)):
            assert_same_bits(o[:, 0:3], p1, f"frame {t} pos")
            assert_same_bits(o[:, 3:6], v1, f"frame {t} vel")
        # and the world is left in the state of the last frame for the per-call API
        p2, v2 = w2.download()
        assert_same_bits(p2, frames[-1][3], "state after the last frame")


@pytest.mark.gpu
def test_cap_velocity_matches_oracle_bitwise(oracle):
    """Device RigidBody::cap_velocity (physics.rs:506-520, caller controllers.rs:307) for flagged bodies: bit-equal to
    the oracle's restatement, untouched bodies keep their bits, and the reference's own bounds hold
    (physics.rs:664-693: horizontal <= 20, vertical within +-40)."""
    n = 5000
    s = scenes.uniform(n, L=40.0, seed=71, rank_seed=72, floor=False)
    rng = np.random.default_rng(9)
    s["vel"] = ((rng.random((n, 3)) - 0.5) * 120.0).astype(np.float32)
    s["vel"][0] = (20.0, 40.0, 0.0)          # exactly at the caps: unchanged
    s["vel"][1] = (12.0, -41.0, 16.0)        # |h| = 20 exactly (not above), y clamped
    s["vel"][2] = (0.0, 0.0, 0.0)
    idx = np.arange(0, n, 3, dtype=np.uint32)
    with world() as w:
        w.upload(s)
        w.cap_velocity(idx)
        _, v = w.download()
        w.cap_velocity(None)
        _, v_all = w.download()
    flagged = np.zeros(n, bool)
    flagged[idx] = True
    want = np.stack([oracle.cap_velocity(x) for x in s["vel"]])
    assert_same_bits(v[flagged], want[flagged], "capped bodies")
    assert_same_bits(v[~flagged], s["vel"][~flagged], "bodies that were not flagged")
    # the second call caps the flagged bodies a SECOND time (a capped horizontal speed can round to just above 20)
    want2 = want.copy()
    want2[flagged] = np.stack([oracle.cap_velocity(x) for x in want[flagged]])
    assert_same_bits(v_all, want2, "all bodies")
    h = np.sqrt(v_all[:, 0].astype(np.float64) ** 2 + v_all[:, 2].astype(np.float64) ** 2)
    assert h.max() <= 20.0 + 1e-4 and np.abs(v_all[:, 1]).max() <= 40.0

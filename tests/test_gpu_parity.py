"""CUDA path vs the CPU oracle, through the C ABI.  Bit-exact: positions, velocities, S_snapshot, S_sweep."""
import json
import os

import numpy as np
import pytest

from gears_b200 import scenes

pytestmark = pytest.mark.gpu

DT = np.float32(1.0) / np.float32(60.0)
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "known_answers.json")))


def bits(x):
    return "0x%08x" % int(np.float32(x).view(np.uint32))


def world(**kw):
    from gears_b200.world import PhysicsWorld
    kw.setdefault("record_pairs", True)
    return PhysicsWorld(**kw)


def assert_same_bits(a, b, what):
    a = np.ascontiguousarray(a, np.float32).view(np.uint32)
    b = np.ascontiguousarray(b, np.float32).view(np.uint32)
    if not np.array_equal(a, b):
        bad = np.argwhere(a != b)
        raise AssertionError(f"{what}: {len(bad)} words differ, first at {bad[0]}: "
                             f"{a[tuple(bad[0])]:#x} vs {b[tuple(bad[0])]:#x}")


def run_and_compare(oracle, scene, steps, mode=0, check_every=1, oracle_pair_cap=None, **wkw):
    """Step the GPU and the oracle in lock step, comparing bits every `check_every` steps."""
    with world(**wkw) as w:
        w.upload(scene)
        s = dict(scene)
        done = 0
        stats = None
        while done < steps:
            k = min(check_every, steps - done)
            for _ in range(k):
                stats = w.step(DT)
            r = oracle.step(s, DT, k, mode=mode, want_snapshot=True, pair_cap=oracle_pair_cap)
            done += k
            pos, vel = w.download()
            assert_same_bits(pos, r["pos"], f"pos after step {done}")
            assert_same_bits(vel, r["vel"], f"vel after step {done}")
            assert np.array_equal(w.pairs(0), r["snapshot"]), f"S_snapshot at step {done}"
            assert np.array_equal(w.pairs(1), r["sweep"]), f"S_sweep at step {done}"
            assert stats["n_snapshot"] == len(r["snapshot"]) and stats["n_contacts"] == len(r["sweep"])
            s["pos"], s["vel"] = r["pos"], r["vel"]
        return stats


def one_body(pos, vel, acc, static=False):
    s = scenes.subset(scenes.uniform(1, L=1.0, floor=False, rank_seed=None), np.arange(1))
    s["pos"][0], s["vel"][0], s["acc"][0] = pos, vel, acc
    s["is_static"][0] = 1 if static else 0
    return s


def test_known_answers_update_pos():
    for k in ("KA1", "KA2", "KA3"):
        g = GOLD[k]
        with world() as w:
            w.upload(one_body(g["pos"], g["vel"], g["acc"]))
            w.step(np.float32(g["dt"]))
            p, v = w.download()
        if "vel_y_bits" in g:
            assert bits(v[0, 1]) == g["vel_y_bits"] and bits(p[0, 1]) == g["pos_y_bits"], k
        else:
            assert [bits(x) for x in v[0]] == g["vel_bits"] and [bits(x) for x in p[0]] == g["pos_bits"], k


def test_damping_factors_match_libm(oracle):
    from gears_b200.world import damping_factors
    for dt in (DT, np.float32(0.016), np.float32(0.1), np.float32(1e-3)):
        assert np.array_equal(damping_factors(dt).view(np.uint32), oracle.damping_factors(dt).view(np.uint32))
    assert [bits(x) for x in damping_factors(DT)] == GOLD["KD"]["damp_bits"]


def test_static_body_does_not_move():  # physics.rs:711-722
    with world() as w:
        w.upload(one_body([0, 5, 0], [0, 0, 0], [0, 0, 0], static=True))
        w.step(np.float32(0.016))
        p, v = w.download()
    assert np.array_equal(p[0], np.float32([0, 5, 0]))


def pair_scene(a, b):
    s = scenes.subset(scenes.uniform(2, L=1.0, floor=False, rank_seed=None), np.arange(2))
    for i, d in enumerate((a, b)):
        h = np.float32(d["half"])
        s["pos"][i], s["vel"][i] = d["pos"], d["vel"]
        s["box_min"][i], s["box_max"][i] = -h, h
        s["mass"][i], s["restitution"][i], s["is_static"][i] = d["mass"], d["restitution"], int(d["is_static"])
    return s


@pytest.mark.parametrize("k", ["KR1", "KR2", "KR3", "KR4", "KR5"])
def test_known_answers_resolve_via_tiny_dt(oracle, k):
    # The C ABI exposes the step, not resolve() alone; a step with dt = 0 leaves pass 1 an exact
    # no-op for these inputs only if velocities survive damping, so compare GPU vs oracle on the full
    # step and, separately, the oracle's resolve vs the golden bits (tests/test_oracle_known_answers.py).
    g = GOLD[k]
    s = pair_scene(g["a"], g["b"])
    run_and_compare(oracle, s, 3)


def test_touching_boxes_do_not_collide(oracle):  # physics.rs:632-642
    a = dict(pos=[0, 0, 0], vel=[0, 0, 0], mass=0, restitution=0.2, is_static=True, half=[1, 1, 1])
    b = dict(pos=[2, 0, 0], vel=[0, 0, 0], mass=1, restitution=0.2, is_static=False, half=[1, 1, 1])
    s = pair_scene(a, b)
    with world() as w:
        w.upload(s)
        st = w.step(np.float32(0.0))  # dt = 0: nothing moves, boxes exactly touch
        assert st["n_snapshot"] == 0 and st["n_contacts"] == 0
        assert st["n_candidates"] == 1  # but the margin-inflated broad phase sees the pair


@pytest.mark.parametrize("order", ["ascending", "shuffle", "reversed"])
def test_physics_example_1000_steps(oracle, order):
    stats = run_and_compare(oracle, scenes.physics_example(order), 1000, check_every=50)
    assert stats["n_contacts"] >= 40  # everything has settled on the ground plane


def test_uniform_100k_single_step(oracle):
    """BASELINE config 2: pair sets and state bit-exact vs the reference algorithm (brute force)."""
    s = scenes.uniform(100_000, L=126.0, seed=1001, rank_seed=2001)
    st = run_and_compare(oracle, s, 1, mode=0)
    assert st["n_contacts"] > 10_000 and st["n_big"] == 1


def test_uniform_identity_rank_multi_step(oracle):
    s = scenes.uniform(20_000, L=60.0, seed=11, rank_seed=None)
    run_and_compare(oracle, s, 10, mode=1, check_every=5)


def test_pile_settling(oracle):
    s = scenes.pile(16, 16, 16)
    st = run_and_compare(oracle, s, 120, mode=1, check_every=40)
    assert st["n_rounds"] > 1


def test_mixed_sizes_and_statics(oracle):
    s = scenes.mixed(20_000, L=110.0)
    run_and_compare(oracle, s, 6, mode=1, check_every=3)


@pytest.mark.parametrize("tiny", ["1", "0"])
@pytest.mark.parametrize("scene", ["uniform", "dense", "mixed", "pile"])
def test_both_repair_paths(oracle, monkeypatch, scene, tiny):
    """GP_TINY_REPAIR=1 (the default): the warp-per-component repair runs in front of the cooperative one; piles exceed
    its component limits and exercise the hand-over between the two.  =0: the cooperative path repairs everything."""
    monkeypatch.setenv("GP_TINY_REPAIR", tiny)
    s = {"uniform": lambda: scenes.uniform(20_000, L=60.0, seed=31, rank_seed=32),
         "dense": lambda: scenes.uniform(20_000, L=40.0, seed=33, rank_seed=34),
         "mixed": lambda: scenes.mixed(20_000, L=110.0),
         "pile": lambda: scenes.pile(12, 12, 12)}[scene]()
    retries = 0
    with world() as w:
        w.upload(s)
        cur = dict(s)
        for step in range(8):
            retries += w.step(DT)["n_retries"]
            r = oracle.step(cur, DT, 1, mode=1, want_snapshot=True)
            pos, vel = w.download()
            assert_same_bits(pos, r["pos"], f"pos after step {step + 1}")
            assert_same_bits(vel, r["vel"], f"vel after step {step + 1}")
            assert np.array_equal(w.pairs(1), r["sweep"]), f"S_sweep at step {step + 1}"
            cur["pos"], cur["vel"] = r["pos"], r["vel"]
    assert retries > 0  # verification appended pairs, so a repair ran


@pytest.mark.parametrize("scene", ["uniform", "dense", "mixed", "pile", "one_cell"])
def test_kernel_paths(oracle, scene):
    """Scenes chosen to reach every branch of K3/K4/K6: windows of the flattened test list (dense), multi-cell
    big-list tests (mixed), heavy chains (pile), and thousands of bodies in ONE grid cell (the body-at-a-time walk of
    k_find_pairs; 12.5 M pairs, every chain heavy)."""
    if scene == "one_cell":
        s = scenes.uniform(5000, L=1.5, seed=41, rank_seed=42)
        run_and_compare(oracle, s, 1, mode=1, max_pairs=40_000_000, oracle_pair_cap=14_000_000)
        return
    s = {"uniform": lambda: scenes.uniform(100_000, L=126.0, seed=1001, rank_seed=2001),
         "dense": lambda: scenes.uniform(20_000, L=40.0, seed=33, rank_seed=34),
         "mixed": lambda: scenes.mixed(20_000, L=110.0),
         "pile": lambda: scenes.pile(12, 12, 12)}[scene]()
    run_and_compare(oracle, s, 4, mode=0 if scene == "uniform" else 1)


def test_rotated_scaled_asymmetric_boxes(oracle):
    n = 4000
    s = scenes.uniform(n, L=30.0, seed=21, rank_seed=22)
    m = len(s["mass"])
    rng = np.random.default_rng(3)
    q = rng.normal(size=(m, 4)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True).astype(np.float32)
    q[-1] = (0, 0, 0, 1)  # keep the floor axis-aligned
    s["rot"] = q
    sc = (0.5 + rng.random((m, 3))).astype(np.float32)
    sc[-1] = 1
    s["scale"] = sc
    s["box_min"][:n] = -(0.2 + rng.random((n, 3))).astype(np.float32) * 0.6
    s["box_max"][:n] = (0.2 + rng.random((n, 3))).astype(np.float32) * 0.6
    run_and_compare(oracle, s, 4, mode=0, check_every=2)


def test_external_writers_set_state_and_dirty(oracle):
    s = scenes.uniform(3000, L=24.0, seed=31, rank_seed=32)
    with world() as w:
        w.upload(s)
        w.step(DT)
        r = oracle.step(s, DT, 1)
        s2 = dict(s, pos=r["pos"], vel=r["vel"])
        # a user system rewrites some velocities/accelerations, another teleports a body
        idx = np.arange(0, 3000, 7, dtype=np.uint32)
        s2["vel"] = s2["vel"].copy()
        s2["acc"] = s2["acc"].copy()
        s2["pos"] = s2["pos"].copy()
        s2["vel"][idx] = (1.0, 20.0, -2.0)
        s2["acc"][idx] = (3.0, 0.0, 1.0)
        w.update_dirty(idx, vel=s2["vel"][idx], acc=s2["acc"][idx])
        s2["pos"][5] = (12.0, 30.0, 12.0)
        w.set_state(pos=s2["pos"])
        w.step(DT)
        r2 = oracle.step(s2, DT, 1)
        p, v = w.download()
        assert_same_bits(p, r2["pos"], "pos")
        assert_same_bits(v, r2["vel"], "vel")


def test_dirty_sets_every_step_from_pinned_memory(oracle):
    """A host that tracks a dirty set: gp_update_dirty of a different subset before EVERY step, from
    pinned buffers (asynchronous copies, no stream synchronisation), synchronous and asynchronous steps mixed; K3
    keeps the id -> slot map current.  Bit-equal to the oracle fed with the same writes."""
    n = 4000
    s = scenes.uniform(n, L=27.0, seed=51, rank_seed=52)
    rng = np.random.default_rng(17)
    cur = dict(s, pos=s["pos"].copy(), vel=s["vel"].copy(), acc=s["acc"].copy())
    with world() as w:
        w.upload(s)
        k_max = n // 5
        hidx = w.host_alloc((k_max,), np.uint32)
        hvel, hacc = w.host_alloc((k_max, 3)), w.host_alloc((k_max, 3))
        for step in range(6):
            k = int(rng.integers(1, k_max))
            idx = rng.choice(n, size=k, replace=False).astype(np.uint32)
            hidx[:k] = idx
            hvel[:k] = (rng.random((k, 3)).astype(np.float32) - np.float32(0.5)) * np.float32(8.0)
            hacc[:k] = (rng.random((k, 3)).astype(np.float32) - np.float32(0.5)) * np.float32(2.0)
            cur["vel"][idx], cur["acc"][idx] = hvel[:k], hacc[:k]
            w.update_dirty(hidx[:k], vel=hvel[:k], acc=hacc[:k])
            if step % 2:
                w.step_async(DT, 1)
                w.sync()
            else:
                w.step(DT)
            r = oracle.step(cur, DT, 1, mode=1, want_snapshot=False)
            cur["pos"], cur["vel"] = r["pos"], r["vel"]
            p, v = w.download()
            assert_same_bits(p, r["pos"], f"pos after step {step + 1}")
            assert_same_bits(v, r["vel"], f"vel after step {step + 1}")


def test_shape_update_moves_body_to_big_list(oracle):
    s = scenes.uniform(2000, L=20.0, seed=41, rank_seed=42)
    with world() as w:
        w.upload(s)
        st0 = w.step(DT)
        r = oracle.step(s, DT, 1)
        s2 = dict(s, pos=r["pos"], vel=r["vel"])
        for k in ("box_min", "box_max", "mass", "restitution", "is_static"):
            s2[k] = s2[k].copy()
        idx = np.array([3, 17], np.uint32)
        s2["box_min"][idx] = (-6, -0.5, -6)
        s2["box_max"][idx] = (6, 0.5, 6)
        s2["mass"][idx] = 50
        w.update_shape(idx, s2["box_min"][idx], s2["box_max"][idx], s2["mass"][idx], s2["restitution"][idx],
                       s2["is_static"][idx])
        st1 = w.step(DT)
        r2 = oracle.step(s2, DT, 1)
        p, v = w.download()
        assert st1["n_big"] == st0["n_big"] + 2
        assert_same_bits(p, r2["pos"], "pos")
        assert_same_bits(v, r2["vel"], "vel")


def test_margin_retry_is_transparent(oracle):
    # deep initial overlaps force position corrections far beyond a tiny margin: the step must re-run
    s = scenes.uniform(5000, L=14.0, seed=51, rank_seed=52)
    with world(margin_init=1e-4, margin_max=2.0) as w:
        w.upload(s)
        st = w.step(DT)
        r = oracle.step(s, DT, 1)
        p, v = w.download()
        assert st["n_retries"] >= 1 and st["status"] == 1
        assert_same_bits(p, r["pos"], "pos")
        assert np.array_equal(w.pairs(1), r["sweep"])


def test_margin_violation_is_reported_not_hidden():
    from gears_b200.world import GpError
    s = scenes.uniform(5000, L=14.0, seed=51, rank_seed=52)
    with world(margin_init=1e-4, no_retry=True) as w:  # no repair sweeps allowed: must say so, not guess
        w.upload(s)
        with pytest.raises(GpError) as e:
            w.step(DT)
        assert e.value.status == -6


def test_pair_capacity_overflow_grows_or_reports(oracle):
    from gears_b200.world import GpError
    s = scenes.uniform(5000, L=14.0, seed=61, rank_seed=62)
    with world(max_pairs=1000) as w:  # far too small: the step grows the buffers and re-runs
        w.upload(s)
        st = w.step(DT)
        r = oracle.step(s, DT, 1)
        assert st["n_retries"] >= 1
        assert np.array_equal(w.pairs(1), r["sweep"])
    with world(max_pairs=1000, no_retry=True) as w:
        w.upload(s)
        with pytest.raises(GpError) as e:
            w.step(DT)
        assert e.value.status == -5


def test_async_steps_equal_sync_steps():
    s = scenes.uniform(30_000, L=70.0, seed=71, rank_seed=72)
    with world(margin_init=0.6) as a, world(margin_init=0.6) as b:  # no async step may need a retry
        a.upload(s)
        b.upload(s)
        for _ in range(8):
            a.step(DT)
        b.step_async(DT, 8)
        st = b.sync()
        pa, va = a.download()
        pb, vb = b.download()
        assert st["steps_inexact"] == 0
        assert_same_bits(pa, pb, "pos")
        assert_same_bits(va, vb, "vel")


def test_deterministic_across_runs():
    s = scenes.pile(12, 12, 12)
    outs = []
    for _ in range(2):
        with world() as w:
            w.upload(s)
            for _ in range(30):
                w.step(DT)
            outs.append(w.download())
    assert_same_bits(outs[0][0], outs[1][0], "pos")
    assert_same_bits(outs[0][1], outs[1][1], "vel")


@pytest.mark.parametrize("n", [0, 1, 2])
def test_tiny_worlds(oracle, n):
    s = scenes.subset(scenes.uniform(4, L=1.5, floor=False), np.arange(n))
    run_and_compare(oracle, s, 3)


def test_all_static_world(oracle):
    s = scenes.uniform(100, L=3.0, floor=True)
    s["is_static"][:] = 1
    with world() as w:
        w.upload(s)
        st = w.step(DT)
        p, v = w.download()
        assert st["n_candidates"] == 0 and st["n_contacts"] == 0
        assert_same_bits(p, s["pos"], "pos")


def test_heavy_chain_body(oracle):
    # one big DYNAMIC platform overlapping hundreds of small bodies: its chain needs the block-wide sorter
    s = scenes.uniform(3000, L=12.0, seed=81, rank_seed=82, floor=True)
    s["box_min"][0] = (-5, -0.3, -5)
    s["box_max"][0] = (5, 0.3, 5)
    s["pos"][0] = (6, 6, 6)
    s["mass"][0] = 500
    st = run_and_compare(oracle, s, 2, mode=0, margin_max=2.0)
    assert st["n_heavy"] >= 1


def test_step_before_upload_is_an_error():
    from gears_b200.world import GpError
    with world() as w:
        with pytest.raises(GpError) as e:
            w.step(DT)
        assert e.value.status == -8


def _scene_golden_cases():
    import importlib.util
    here = os.path.dirname(__file__)
    spec = importlib.util.spec_from_file_location("make_scene_goldens", os.path.join(here, "golden", "make_scene_goldens.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    return gen


@pytest.mark.parametrize("name", ["physics_example_ascending_1000", "physics_example_shuffle_1000",
                                  "uniform_5k_shuffled_20", "pile_8x8x8_60", "mixed_3k_10"])
def test_scene_goldens(name):
    """CUDA path vs the COMMITTED fixtures (no live oracle involved)."""
    gen = _scene_golden_cases()
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "scene_goldens.json")))["cases"][name]
    scene, steps = next((s, k) for n, s, k in gen.cases() if n == name)
    with world() as w:
        w.upload(scene)
        for _ in range(steps):
            w.step(DT)
        pos, vel = w.download()
        sweep, snap = w.pairs(1), w.pairs(0)
    assert gen.digest(pos) == gold["pos_sha256"] and gen.digest(vel) == gold["vel_sha256"]
    assert len(sweep) == gold["n_sweep"] and gen.digest(sweep) == gold["sweep_sha256"]
    assert len(snap) == gold["n_snapshot"] and gen.digest(snap) == gold["snapshot_sha256"]


def test_nan_body_does_not_poison_the_others(oracle):
    """A body with a NaN position (an external writer gone wrong) must neither crash the step nor change any other
    body: every comparison with NaN is false, in the reference as here."""
    s = scenes.uniform(3000, L=20.0, seed=91, rank_seed=92)
    s["pos"][7] = (np.nan, 5.0, 5.0)
    with world() as w:
        w.upload(s)
        for _ in range(3):
            w.step(DT)
        pos, vel = w.download()
    r = oracle.step(s, DT, 3, mode=0)
    ok = np.ones(len(pos), bool)
    ok[7] = False
    assert_same_bits(pos[ok], r["pos"][ok], "pos of the finite bodies")
    assert_same_bits(vel[ok], r["vel"][ok], "vel of the finite bodies")
    assert np.isnan(pos[7, 0]) and np.isnan(r["pos"][7, 0])

"""Multi-GPU slab decomposition, exercised on ONE GPU: n slab worlds in one process (gp_comm_init_local +
gp_step_group, peer copies instead of NCCL).  The acceptance bar: every owned body the verification does not flag is
bit-identical to the single-world result, and on these scenes nothing is flagged."""
import numpy as np
import pytest

from gears_b200 import scenes

pytestmark = pytest.mark.gpu

DT = np.float32(1.0) / np.float32(60.0)


def single_world_result(scene, steps, **kw):
    from gears_b200.world import PhysicsWorld
    with PhysicsWorld(**kw) as w:
        w.upload(scene)
        w.step_async(DT, steps)
        st = w.sync()
        assert st["steps_inexact"] == 0
        pos, vel = w.download()
    return pos, vel


def assert_same_bits(a, b, what):
    a = np.ascontiguousarray(a, np.float32).view(np.uint32)
    b = np.ascontiguousarray(b, np.float32).view(np.uint32)
    if not np.array_equal(a, b):
        bad = np.argwhere(a != b)
        raise AssertionError(f"{what}: {len(bad)} words differ, first at {bad[0]}")


@pytest.mark.parametrize("nranks", [2, 3, 4])
def test_local_slabs_bit_exact_vs_single_world(nranks):
    """BASELINE config 4 in miniature (same number density 0.05): slabs + halo + migration, checked every step."""
    from gears_b200.slabs import LocalSlabs
    from gears_b200.world import PhysicsWorld
    s = scenes.uniform(60_000, L=106.0, seed=101, rank_seed=102)
    n = s["pos"].shape[0]
    steps = 12
    with PhysicsWorld() as w, LocalSlabs(s, nranks, halo=6.0) as g:
        w.upload(s)
        mig = halo = unverified = 0
        for k in range(steps):
            w.step(DT)
            ref_pos, ref_vel = w.download()
            g.step(DT, 1)
            for st in g.sync():
                mig += st["n_migrated_out"]
                halo += st["n_halo_in"]
                unverified += st["n_unverified"]
                assert st["steps_inexact"] == 0
            ent, pos, vel = g.download()
            assert len(ent) == n and np.array_equal(ent, np.arange(n, dtype=np.uint32)), "every body owned exactly once"
            assert_same_bits(pos, ref_pos, f"pos after step {k + 1}")
            assert_same_bits(vel, ref_vel, f"vel after step {k + 1}")
    assert mig > 0 and halo > 0, "the scene must exercise migration and the halo"
    assert unverified == 0, "with a 6-extent halo every owned body is provably independent of unseen bodies"


def test_unflagged_bodies_are_exact_in_a_dense_scene():
    """Denser scene, default halo: a few bodies cannot be proven independent; every body that differs from the
    single-world result (if any) must be among the flagged ones."""
    from gears_b200.slabs import LocalSlabs
    from gears_b200.world import PhysicsWorld
    s = scenes.uniform(60_000, L=90.0, seed=101, rank_seed=102)
    with PhysicsWorld() as w, LocalSlabs(s, 3) as g:
        w.upload(s)
        flagged_total = 0
        for k in range(8):
            w.step(DT)
            ref_pos, ref_vel = w.download()
            g.step(DT, 1)
            g.sync()
            ent, pos, vel = g.download()
            u = g.last_unverified
            bad = (pos.view(np.uint32) != ref_pos.view(np.uint32)).any(axis=1) | \
                  (vel.view(np.uint32) != ref_vel.view(np.uint32)).any(axis=1)
            assert not (bad & (u == 0)).any(), f"step {k + 1}: an unflagged body differs from the single-world result"
            flagged_total += int(u.sum())
            if bad.any():
                break  # from here on the inputs of the two runs differ
        assert flagged_total < 8 * 600  # a small minority of the 60k bodies


def test_local_slabs_async_batches_and_equal_width_faces():
    from gears_b200.slabs import LocalSlabs
    s = scenes.uniform(30_000, L=84.0, seed=111, rank_seed=None)  # identity ranks, number density 0.05
    n = s["pos"].shape[0]
    ref_pos, ref_vel = single_world_result(s, 20)
    with LocalSlabs(s, 3, mode="width", halo=6.0) as g:
        g.step(DT, 20)  # twenty steps enqueued without any host synchronisation
        sts = g.sync()
        assert sum(st["unverified_total"] for st in sts) == 0
        assert all(st["steps_inexact"] == 0 for st in sts)
        ent, pos, vel = g.download()
    assert np.array_equal(ent, np.arange(n, dtype=np.uint32))
    assert_same_bits(pos, ref_pos, "pos")
    assert_same_bits(vel, ref_vel, "vel")


def test_too_small_halo_is_reported_not_hidden():
    from gears_b200.slabs import LocalSlabs
    s = scenes.uniform(40_000, L=60.0, seed=121, rank_seed=122)  # denser: contact chains cross the faces
    with LocalSlabs(s, 2, halo=0.05) as g:  # far less than one body extent
        g.step(DT, 5)
        sts = g.sync()
        assert sum(st["unverified_total"] for st in sts) > 0


def test_pile_across_a_face_counts_what_it_cannot_prove():
    # a settling pile is one big contact component: the slabs cannot prove independence, and must say so
    from gears_b200.slabs import LocalSlabs
    s = scenes.pile(16, 8, 8)
    with LocalSlabs(s, 2) as g:
        g.step(DT, 3)
        sts = g.sync()
        assert sum(st["unverified_total"] for st in sts) > 0
        ent, pos, vel = g.download()
        assert len(np.unique(ent)) == s["pos"].shape[0]


def test_mixed_sizes_and_statics_on_slabs():
    """BASELINE config 5 in miniature: 100:1 extent ratio, 10 % static bodies, a replicated floor; the grid of a slab
    world hosts every dynamic body (only static bodies may be on the brute-force list)."""
    from gears_b200.slabs import LocalSlabs
    from gears_b200.world import PhysicsWorld
    s = scenes.mixed(40_000, L=540.0, seed=141, rank_seed=142)  # the number density of config 5
    n = s["pos"].shape[0]
    with PhysicsWorld() as w, LocalSlabs(s, 2) as g:
        w.upload(s)
        flagged = 0
        for k in range(6):
            w.step(DT)
            ref_pos, ref_vel = w.download()
            g.step(DT, 1)
            sts = g.sync()
            assert all(st["steps_inexact"] == 0 for st in sts)
            ent, pos, vel = g.download()
            u = g.last_unverified
            dyn = s["is_static"][ent] == 0
            assert np.array_equal(np.sort(ent[dyn]), np.nonzero(s["is_static"] == 0)[0]), "dynamic bodies owned once"
            # static bodies near a face may be reported by both ranks (they are never written): compare by id
            bad = (pos.view(np.uint32) != ref_pos[ent].view(np.uint32)).any(axis=1) | \
                  (vel.view(np.uint32) != ref_vel[ent].view(np.uint32)).any(axis=1)
            assert not (bad & (u == 0)).any(), f"step {k + 1}: an unflagged body differs from the single-world result"
            flagged += int(u.sum())
            if bad.any():
                break
        assert flagged < 6 * 0.02 * n


def test_exchange_buffer_overflow_is_reported_not_hidden():
    """A face buffer that is too small must surface as an error at sync (GP_ERR_CAPACITY), never as silently
    missing halo copies."""
    from gears_b200.slabs import LocalSlabs
    from gears_b200.world import GpError
    s = scenes.uniform(40_000, L=60.0, seed=151, rank_seed=152)
    with LocalSlabs(s, 2, max_exchange=16) as g:   # a face needs thousands of records
        g.step(DT, 2)
        with pytest.raises(GpError) as e:
            g.sync()
        assert e.value.status == -4


def test_slab_width_below_twice_the_halo_is_rejected():
    from gears_b200.slabs import LocalSlabs
    from gears_b200.world import GpError
    s = scenes.uniform(20_000, L=40.0, seed=161, rank_seed=162)
    with pytest.raises(GpError) as e:
        LocalSlabs(s, 4, halo=8.0)   # inner slabs are ~10 wide: a body could reach a rank two slabs away
    assert e.value.status == -1


def test_slab_world_rejects_per_index_calls():
    from gears_b200.slabs import LocalSlabs
    from gears_b200.world import GpError
    s = scenes.uniform(5_000, L=40.0, seed=131)
    with LocalSlabs(s, 2) as g:
        with pytest.raises(GpError):
            g.worlds[0].download()
        with pytest.raises(GpError):
            g.worlds[0].step_async(DT, 1)  # in-process slabs advance together


def test_owned_rows_can_be_written_back():
    """gp_download_owned is in deterministic device order and gp_set_state_owned hands the same rows back (what
    external systems changed between two steps): slab worlds fed that way equal a single world fed through
    gp_update_dirty; a step in between invalidates the order (GP_ERR_STATE)."""
    from gears_b200.slabs import LocalSlabs
    from gears_b200.world import GpError, PhysicsWorld
    s = scenes.uniform(40_000, L=92.0, seed=111, rank_seed=112)
    n = s["pos"].shape[0]
    rng = np.random.default_rng(23)
    with PhysicsWorld() as w, LocalSlabs(s, 3, halo=6.0) as g:
        w.upload(s)
        for step in range(5):
            # the same external write on both sides: new velocities / accelerations for a random tenth of the bodies
            idx = np.sort(rng.choice(n - 1, size=n // 10, replace=False)).astype(np.uint32)
            nv = ((rng.random((len(idx), 3)) - 0.5) * 6.0).astype(np.float32)
            na = ((rng.random((len(idx), 3)) - 0.5) * 2.0).astype(np.float32)
            w.update_dirty(idx, vel=nv, acc=na)
            new_v, new_a = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
            has = np.zeros(n, bool)
            new_v[idx], new_a[idx], has[idx] = nv, na, True
            for sw in g.worlds:
                ent, pos, vel = sw.download_owned()
                ent2, _, _ = sw.download_owned()
                assert np.array_equal(ent, ent2), "the order of gp_download_owned is deterministic"
                vel = vel.copy()
                acc = np.zeros_like(vel)          # (the scene's accelerations are zero until written)
                m = has[ent]
                vel[m] = new_v[ent[m]]
                if step:                           # accelerations written at earlier steps persist
                    acc = prev_acc[ent]
                acc[m] = new_a[ent[m]]
                sw.set_state_owned(vel=vel, acc=acc)
            prev_acc = (prev_acc if step else np.zeros((n, 3), np.float32))
            prev_acc[idx] = na
            w.step(DT)
            g.step(DT, 1)
            g.sync()
            ref_pos, ref_vel = w.download()
            ent, pos, vel = g.download()
            assert np.array_equal(ent, np.arange(n, dtype=np.uint32))
            assert_same_bits(pos, ref_pos, f"pos after step {step + 1}")
            assert_same_bits(vel, ref_vel, f"vel after step {step + 1}")
        sw = g.worlds[0]
        ent, pos, vel = sw.download_owned()
        g.step(DT, 1)
        g.sync()
        with pytest.raises(GpError) as e:
            sw.set_state_owned(vel=vel)
        assert e.value.status == -8


def test_rebalancing_keeps_a_drifting_scene_even_and_exact():
    """SURVEY 8e "re-balanced every K steps": every body is pushed along +x, so with fixed faces the left slab
    empties; gp_slab_rebalance_local every 5 steps keeps the owned counts within 10 % of each other while the result
    stays bit-identical to the single world (bodies change owner through the ordinary migration path).  No floor: a
    layer of bodies resting on it is one connected component, which the exactness proof flags wholesale."""
    from gears_b200.slabs import LocalSlabs
    from gears_b200.world import PhysicsWorld
    n = 30_000
    s = scenes.uniform(n, L=84.0, seed=121, rank_seed=122, floor=False)
    s["acc"][:] = (30.0, 0.0, 0.0)            # steady drift of ~19 units/s (physics.rs:405-462: damped, then + a*dt)
    nranks, steps, every = 3, 60, 5

    def spread(counts):
        c = np.asarray(counts, np.float64)
        return (c.max() - c.min()) / c.mean()

    with LocalSlabs(s, nranks, halo=6.0) as fixed:
        fixed.step(DT, steps)
        fixed.sync()
        drift_spread = spread(fixed.owned_counts())
    with PhysicsWorld() as w, LocalSlabs(s, nranks, halo=6.0) as g:
        w.upload(s)
        moved_any, unverified = False, 0
        for k in range(steps // every):
            w.step_async(DT, every)
            g.step(DT, every)
            for st in g.sync():
                assert st["steps_inexact"] == 0
                unverified += st["unverified_total"]
            if k % 4 == 3:
                ref_pos, ref_vel = w.download()
                ent, pos, vel = g.download()
                assert np.array_equal(ent, np.arange(n, dtype=np.uint32)), "every body owned exactly once"
                assert_same_bits(pos, ref_pos, f"pos after step {every * (k + 1)}")
                assert_same_bits(vel, ref_vel, f"vel after step {every * (k + 1)}")
            moved_any |= g.rebalance()
        assert w.sync()["steps_inexact"] == 0
        balanced_spread = spread(g.owned_counts())
    assert moved_any
    assert unverified == 0
    assert drift_spread > 0.3, f"the scene must drift enough to unbalance fixed faces (spread {drift_spread:.2f})"
    assert balanced_spread < 0.10, f"owned counts differ by {balanced_spread:.2%} of the mean after re-balancing"

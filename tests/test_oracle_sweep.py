"""The grid-accelerated oracle must reproduce the brute-force (reference-algorithm) oracle bit for bit,
and the sweep must honour the reference's ordering semantics."""
import numpy as np
import pytest

from gears_b200 import scenes

DT = np.float32(1.0) / np.float32(60.0)
KEYS = ("pos", "vel", "snapshot", "sweep", "static_static")


def same(a, b):
    return all(np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)) for k in KEYS)


@pytest.mark.parametrize("name,scene,steps", [
    ("example_asc", lambda: scenes.physics_example(), 200),
    ("example_shuffle", lambda: scenes.physics_example("shuffle"), 200),
    ("example_rev", lambda: scenes.physics_example("reversed"), 200),
    ("uniform_dense", lambda: scenes.uniform(4000, L=28.0, seed=5), 10),
    ("uniform_identity_rank", lambda: scenes.uniform(3000, L=26.0, seed=6, rank_seed=None), 5),
    ("pile", lambda: scenes.pile(8, 8, 8), 60),
    ("mixed", lambda: scenes.mixed(3000, L=55.0), 8),
])
def test_grid_oracle_matches_brute_force(oracle, name, scene, steps):
    s = scene()
    a = oracle.step(s, DT, steps, mode=0)
    b = oracle.step(s, DT, steps, mode=1)
    assert same(a, b), name
    assert len(a["sweep"]) > 0


def test_order_matters_and_is_an_input(oracle):
    s = scenes.uniform(2000, L=18.0, seed=9, rank_seed=None)
    a = oracle.step(s, DT, 3)
    s2 = dict(s)
    s2["rank"] = scenes.shuffled_rank(len(s["mass"]), 77)
    b = oracle.step(s2, DT, 3)
    assert not np.array_equal(a["pos"], b["pos"])  # Gauss-Seidel: order changes the result


def test_pairs_are_in_iteration_order(oracle):
    s = scenes.uniform(1500, L=16.0, seed=3)
    r = oracle.step(s, DT, 1)
    rank = s["rank"]
    for key in ("snapshot", "sweep"):
        p = r[key]
        k = rank[p[:, 0]].astype(np.int64) * (1 << 32) + rank[p[:, 1]]
        assert np.all(rank[p[:, 0]] < rank[p[:, 1]])
        assert np.all(np.diff(k) > 0)


def test_static_bodies_never_move(oracle):
    s = scenes.mixed(2000, L=40.0)
    r = oracle.step(s, DT, 5)
    st = s["is_static"] != 0
    assert np.array_equal(r["pos"][st], s["pos"][st]) and np.array_equal(r["vel"][st], s["vel"][st])


def test_rank_must_be_permutation(oracle):
    s = scenes.uniform(10, L=5.0)
    s["rank"] = np.zeros(11, np.uint32)
    with pytest.raises(ValueError):
        oracle.step(s, DT, 1)


def test_empty_scene(oracle):
    s = scenes.subset(scenes.uniform(4, L=3.0), np.zeros(0, np.int64))
    r = oracle.step(s, DT, 2)
    assert r["pos"].shape == (0, 3) and len(r["sweep"]) == 0

"""BASELINE.json's FULL-SIZE configurations on the CUDA path, bit for bit against the CPU oracle.

The oracle (grid mode, bit-identical to the reference's brute-force sweep: tests/test_oracle_sweep.py) needs minutes per
case at these sizes, so its results are committed as digests (tests/golden/large_goldens.json, written by
tests/golden/make_large_goldens.py) and the GPU is compared with those: positions, velocities and the S_sweep pair list
(update_systems.rs:408-487 at its turn) of pile_1M after 1 / 10 / 100 steps, mixed_4M and uniform_16M after one step.
"""
import importlib.util
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
DT = np.float32(1.0) / np.float32(60.0)


def _gen():
    spec = importlib.util.spec_from_file_location("make_large_goldens", os.path.join(HERE, "golden", "make_large_goldens.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _gold():
    p = os.path.join(HERE, "golden", "large_goldens.json")
    return json.load(open(p))["cases"] if os.path.exists(p) else {}


@pytest.mark.parametrize("name", ["pile_1M", "mixed_4M", "uniform_16M"])
def test_full_size_config_matches_the_oracle_digests(name):
    gold = _gold().get(name)
    if gold is None:
        pytest.skip(f"no committed digests for {name} (run tests/golden/make_large_goldens.py {name})")
    from gears_b200.world import PhysicsWorld
    gen = _gen()
    scene = gen.make_scene(gold["spec"])
    assert int(len(scene["mass"])) == gold["n"]
    n = gold["n"]
    with PhysicsWorld(record_pairs=True, max_pairs=max(1 << 22, 4 * n)) as w:
        w.upload(scene)
        done = 0
        for cp in sorted(int(k) for k in gold["checkpoints"]):
            g = gold["checkpoints"][str(cp)]
            st = None
            for _ in range(cp - done):
                st = w.step(DT)
            done = cp
            pos, vel = w.download()
            sweep = w.pairs(1)
            first = ["0x%08x" % int(x) for x in pos.view(np.uint32).reshape(-1)[:6]]
            assert first == g["pos_first_bits"], f"{name} step {cp}: first position words {first}"
            assert gen.digest(pos) == g["pos_sha256"], f"{name} step {cp}: positions differ from the oracle"
            assert gen.digest(vel) == g["vel_sha256"], f"{name} step {cp}: velocities differ from the oracle"
            assert len(sweep) == g["n_sweep"] == st["n_contacts"], f"{name} step {cp}: |S_sweep| {len(sweep)} vs {g['n_sweep']}"
            assert gen.digest(sweep) == g["sweep_sha256"], f"{name} step {cp}: S_sweep differs from the oracle"

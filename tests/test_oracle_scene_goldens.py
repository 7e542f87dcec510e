"""The oracle still produces the committed whole-step fixtures (tests/golden/scene_goldens.json)."""
import importlib.util
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_scene_goldens", os.path.join(HERE, "golden", "make_scene_goldens.py"))
gen = importlib.util.module_from_spec(spec)
spec.loader.exec_module(gen)
GOLD = json.load(open(os.path.join(HERE, "golden", "scene_goldens.json")))["cases"]


@pytest.mark.parametrize("name,scene,steps", list(gen.cases()), ids=[c[0] for c in gen.cases()])
def test_oracle_reproduces_scene_golden(oracle, name, scene, steps):
    g = GOLD[name]
    r = oracle.step(scene, gen.DT, steps, mode=0)
    assert gen.digest(r["pos"]) == g["pos_sha256"] and gen.digest(r["vel"]) == g["vel_sha256"]
    assert gen.digest(r["sweep"]) == g["sweep_sha256"] and len(r["sweep"]) == g["n_sweep"]
    assert gen.digest(r["snapshot"]) == g["snapshot_sha256"]
    # the grid-accelerated oracle mode visits the same pairs in the same order
    r1 = oracle.step(scene, gen.DT, steps, mode=1)
    assert gen.digest(r1["pos"]) == g["pos_sha256"] and gen.digest(r1["sweep"]) == g["sweep_sha256"]

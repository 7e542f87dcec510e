"""Probe the slab verification: python tools/slab_probe.py n L nranks halo steps"""
import sys, numpy as np
sys.path.insert(0, '.')
from gears_b200 import scenes
from gears_b200.slabs import LocalSlabs
from gears_b200.world import PhysicsWorld
n, L, nr, halo, steps = int(sys.argv[1]), float(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]), int(sys.argv[5])
DT = np.float32(1) / np.float32(60)
s = scenes.uniform(n, L=L, seed=101, rank_seed=102)
with PhysicsWorld() as w:
    w.upload(s)
    ref = []
    for k in range(steps):
        st = w.step(DT); p, v = w.download(); ref.append((p.copy(), v.copy()))
    print("single: margin", st["margin"], "cand", st["n_candidates"], "contacts", st["n_contacts"])
with LocalSlabs(s, nr, halo=halo) as g:
    for k in range(steps):
        g.step(DT, 1)
        sts = g.sync()
        e, p, v = g.download()
        u = g.last_unverified
        bad = (p.view(np.uint32) != ref[k][0].view(np.uint32)).any(axis=1) | (v.view(np.uint32) != ref[k][1].view(np.uint32)).any(axis=1)
        print(k, "unverified", [st["n_unverified"] for st in sts], "flagged", int(u.sum()), "mismatch", int(bad.sum()),
              "mismatch&~flagged", int((bad & (u == 0)).sum()), "margin", [round(st["margin"], 3) for st in sts],
              "halo_in", [st["n_halo_in"] for st in sts], "mig", [st["n_migrated_out"] for st in sts], "inexact", [st["steps_inexact"] for st in sts])

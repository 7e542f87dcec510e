"""Two slab worlds on ONE GPU (in-process transport) for profiling: python tools/slab_local_run.py n L steps"""
import sys, numpy as np
sys.path.insert(0, '.')
from gears_b200 import scenes
from gears_b200.slabs import LocalSlabs
n, L, steps = int(sys.argv[1]), float(sys.argv[2]), int(sys.argv[3])
DT = np.float32(1) / np.float32(60)
s = scenes.uniform(n, L=L, seed=1003, rank_seed=2003)
with LocalSlabs(s, 2, halo=4.0, max_pairs=4 * n) as g:
    g.step(DT, steps)
    sts = g.sync()
    print([ (st["n_bodies"], st["n_halo_in"], st["n_unverified"], st["steps_inexact"]) for st in sts])

#!/usr/bin/env python
"""How violent is the settled pile?  Distribution of per-step displacement at steps ~300 of pile_1M."""
import sys, numpy as np
sys.path.insert(0, '.')
import bench
from gears_b200.world import PhysicsWorld
gen, wkw = bench.WORKLOADS["pile_1M"]
s = bench.make_scene(gen); n = len(s["mass"]); DT = bench.DT
with PhysicsWorld(**wkw) as w:
    w.upload(s)
    for k in range(300):
        st = w.step(DT)
    p0, v0 = w.download()
    st = w.step(DT)
    p1, v1 = w.download()
print({k: st[k] for k in ("n_candidates", "n_contacts", "n_rounds", "n_retries", "margin", "max_displacement")})
dyn = s["is_static"] == 0
d = np.abs(p1 - p0).max(axis=1)[dyn]          # per-axis displacement over the step (integration + corrections)
free = np.abs(v1 * DT).max(axis=1)[dyn]       # what the final velocity alone would move it
sp = np.abs(v1).max(axis=1)[dyn]
q = [0.5, 0.9, 0.99, 0.999, 1.0]
print("displacement per step, quantiles", q, np.quantile(d, q))
print("|v|*dt quantiles", np.quantile(free, q), " speed quantiles", np.quantile(sp, q))
for thr in (0.05, 0.1, 0.2, 0.5, 1.0):
    print(f"bodies moving more than {thr}: {(d > thr).mean():.4f}")
print("height quantiles", np.quantile(p1[dyn, 1], [0.01, 0.5, 0.99]))

"""Find the first step at which the GPU and the oracle disagree, and how."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gears_b200 import scenes
from gears_b200.world import PhysicsWorld
from oracle import pyoracle as po

dt = np.float32(1) / np.float32(60)
name = sys.argv[1] if len(sys.argv) > 1 else "pile"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
s = {"pile": lambda: scenes.pile(16, 16, 16), "pile8": lambda: scenes.pile(8, 8, 8)}[name]()
rank = s["rank"] if s["rank"] is not None else np.arange(len(s["mass"]), dtype=np.uint32)
with PhysicsWorld(record_pairs=True) as w:
    w.upload(s)
    cur = dict(s)
    for k in range(1, steps + 1):
        st = w.step(dt)
        r = po.step(cur, dt, 1, mode=1)
        p, v = w.download()
        snap, sweep = w.pairs(0), w.pairs(1)
        okp = np.array_equal(p.view(np.uint32), r["pos"].view(np.uint32))
        okv = np.array_equal(v.view(np.uint32), r["vel"].view(np.uint32))
        oks = np.array_equal(snap, r["snapshot"])
        okw = np.array_equal(sweep, r["sweep"])
        print(k, "pos", okp, "vel", okv, "snap", oks, len(snap), len(r["snapshot"]), "sweep", okw, len(sweep),
              len(r["sweep"]), {x: st[x] for x in ("n_candidates", "n_rounds", "n_retries", "margin", "max_displacement", "cell_size", "n_heavy", "grid_dim")})
        if not (okp and okv and oks and okw):
            A = set(map(tuple, sweep.tolist())); B = set(map(tuple, r["sweep"].tolist()))
            print("sweep only gpu:", sorted(A - B)[:10], "only oracle:", sorted(B - A)[:10])
            A = set(map(tuple, snap.tolist())); B = set(map(tuple, r["snapshot"].tolist()))
            print("snap only gpu:", sorted(A - B)[:10], "only oracle:", sorted(B - A)[:10])
            bad = np.argwhere((p.view(np.uint32) != r["pos"].view(np.uint32)).any(axis=1)).ravel()
            print("bodies with pos diff:", len(bad), bad[:10], "ranks", rank[bad[:10]])
            for bidx in bad[:3]:
                print(bidx, p[bidx], r["pos"][bidx], v[bidx], r["vel"][bidx])
            break
        cur["pos"], cur["vel"] = r["pos"], r["vel"]

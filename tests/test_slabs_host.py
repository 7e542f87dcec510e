"""Host side of the multi-GPU slab decomposition, on the CPU: faces, partition, and (world_size 2, gloo) the
rendezvous plumbing bench.py uses.  No GPU compute here; the device side is covered by tests/test_gpu_slabs.py."""
import os
import subprocess
import sys
import textwrap

import numpy as np

from gears_b200 import scenes, slabs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_faces_are_sorted_and_open_ended():
    s = scenes.uniform(20_000, L=100.0, seed=5)
    rep = slabs.big_static_mask(s, 4.0)
    for mode in ("count", "balanced", "width"):
        f = slabs.bounds(s, 4, mode, replicate=rep)
        assert len(f) == 5 and f[0] == -np.inf and f[-1] == np.inf
        assert np.all(np.diff(f[1:-1]) > 0)
    assert list(slabs.bounds(s, 1)) == [-np.inf, np.inf]


def test_balanced_faces_give_inner_slabs_less_to_own():
    s = scenes.uniform(200_000, L=100.0, seed=6, floor=False)
    f = slabs.bounds(s, 4, "balanced", halo=4.0)
    cx = slabs.centres_x(s)
    own = np.histogram(cx, bins=np.concatenate([[-1e30], f[1:-1], [1e30]]))[0]
    assert own.sum() == 200_000
    assert own[1] < own[0] and own[2] < own[3]          # inner slabs mirror two faces
    total = own + np.array([1, 2, 2, 1]) * 200_000 * (4.0 + 1.5) / 100.0
    assert total.max() / total.min() < 1.05             # owned + mirrored is level


def test_partition_owns_every_body_once_and_replicates_big_statics():
    s = scenes.uniform(30_000, L=80.0, seed=7, rank_seed=8)
    n = s["pos"].shape[0]
    rep = slabs.big_static_mask(s, 4.0)
    assert rep.sum() == 1 and rep[-1]                    # the floor
    faces = slabs.bounds(s, 3, "count", replicate=rep)
    parts = slabs.partition(s, faces, rep)
    owned = np.concatenate([p["entity"][p["is_static"] == 0] for p in parts])
    assert len(owned) == n - 1 and len(np.unique(owned)) == n - 1
    for r, p in enumerate(parts):
        assert (p["entity"] == n - 1).sum() == 1         # every rank has the floor
        dyn = p["is_static"] == 0
        cx = slabs.centres_x(p)[dyn]
        assert np.all(cx >= faces[r]) and np.all(cx < faces[r + 1])
        # global iteration ranks travel with the bodies
        assert np.array_equal(p["rank"], s["rank"][p["entity"]])


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r})
    import numpy as np, torch, torch.distributed as dist
    from gears_b200 import scenes, slabs
    rank, ws = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    s = scenes.uniform(20_000, L=60.0, seed=9, rank_seed=10)
    rep = slabs.big_static_mask(s, 4.0)
    faces = slabs.bounds(s, ws, "balanced", replicate=rep)
    part = slabs.partition(s, faces, rep)[rank]
    # every rank derives the same faces from the same scene; the owned sets tile the scene
    f = torch.tensor(faces[1:-1].astype(np.float64)); g = [torch.zeros_like(f) for _ in range(ws)]
    dist.all_gather(g, f)
    assert all(torch.equal(x, f) for x in g)
    cnt = torch.tensor([int((part["is_static"] == 0).sum())]); dist.all_reduce(cnt)
    assert int(cnt.item()) == 20_000
    # the unique-id broadcast used by slabs.comm_init (128 opaque bytes from rank 0)
    uid = torch.arange(128, dtype=torch.uint8) if rank == 0 else torch.zeros(128, dtype=torch.uint8)
    dist.broadcast(uid, src=0)
    assert bytes(uid.tolist()) == bytes(range(128))
    cfg = slabs.slab_config(rank, ws, faces, halo=4.0)
    assert cfg.rank == rank and cfg.nranks == ws and cfg.x_lo < cfg.x_hi
    dist.destroy_process_group()
    print("ok", rank)
""")


def test_world_size_2_gloo_rendezvous(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29641", WORLD_SIZE="2", OMP_NUM_THREADS="1")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"ok {r}" in o, o


def test_slab_ownership_uses_the_rotated_scaled_aabb_centre():
    """centres_x must give the very bits of the world AABB centre the device's ownership test computes
    (k_pack corner offsets + k_keys), rotation and scale included: here checked against the oracle's world AABB
    (physics.rs:84-135)."""
    from oracle import pyoracle
    s = scenes.uniform(300, L=12.0, seed=5, rank_seed=6)
    m = len(s["mass"])
    rng = np.random.default_rng(11)
    q = rng.normal(size=(m, 4)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True).astype(np.float32)
    s["rot"] = q
    s["scale"] = (0.5 + rng.random((m, 3))).astype(np.float32)
    s["box_min"][:] = -(0.2 + rng.random((m, 3))).astype(np.float32)
    cx = slabs.centres_x(s)
    for i in range(m):
        lo, hi = pyoracle.world_aabb(s["box_min"][i], s["box_max"][i], s["pos"][i], s["rot"][i], s["scale"][i])
        c = np.float32((lo[0] + hi[0]) * np.float32(0.5))
        assert c.view(np.uint32) == cx[i].view(np.uint32), i

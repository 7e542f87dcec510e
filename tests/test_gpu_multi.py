"""Two processes, two GPUs (one rank per GPU): the slab worlds exchange halo copies and migrating bodies with the
peer-to-peer transport (and, second case, with ncclSend/ncclRecv) and must reproduce the single-GPU result bit for
bit.  Skipped on boxes with fewer than two GPUs (the single-GPU slab tests cover the same logic in one process)."""
import os
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r})
    import numpy as np, torch, torch.distributed as dist
    from gears_b200 import scenes, slabs
    from gears_b200.world import PhysicsWorld
    rank, ws, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    DT = np.float32(1) / np.float32(60)
    steps = 10
    s = scenes.uniform(200_000, L=158.7, seed=301, rank_seed=302)
    rep = slabs.big_static_mask(s, 4.0)
    faces = slabs.bounds(s, ws, "balanced", replicate=rep, halo=6.0)
    part = slabs.partition(s, faces, rep)[rank]
    w = PhysicsWorld(device=lr, max_bodies=int(part["pos"].shape[0] * 1.5) + 65536, big_static_only=True)
    w.upload(part)
    slabs.comm_init(w, dist, rank, ws, faces, halo=6.0, device=torch.device("cuda", lr))
    w.step_async(DT, steps)
    st = w.sync()
    ent, pos, vel = w.download_owned()
    out = [None] * ws if rank == 0 else None
    dist.gather_object((ent, pos, vel, st["unverified_total"], st["steps_inexact"], st["n_migrated_out"]), out, dst=0)
    if rank == 0:
        e = np.concatenate([o[0] for o in out]); p = np.concatenate([o[1] for o in out]); v = np.concatenate([o[2] for o in out])
        order = np.argsort(e, kind="stable")
        n = s["pos"].shape[0]
        assert np.array_equal(e[order], np.arange(n, dtype=np.uint32)), "every body owned exactly once"
        assert sum(o[3] for o in out) == 0 and sum(o[4] for o in out) == 0
        with PhysicsWorld(device=lr) as ref:
            ref.upload(s)
            ref.step_async(DT, steps)
            assert ref.sync()["steps_inexact"] == 0
            rp, rv = ref.download()
        assert np.array_equal(p[order].view(np.uint32), rp.view(np.uint32)), "positions differ from the single-GPU run"
        assert np.array_equal(v[order].view(np.uint32), rv.view(np.uint32)), "velocities differ from the single-GPU run"
        print("MULTI_OK transport=" + os.environ.get("GP_SLAB_TRANSPORT", "p2p"))
    w.close()
    dist.destroy_process_group()
""")


def _gpus() -> int:
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("transport", ["p2p", "nccl"])
def test_two_ranks_match_single_gpu(tmp_path, transport):
    if _gpus() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ)
    if transport == "nccl":
        env["GP_SLAB_TRANSPORT"] = "nccl"
    else:
        env.pop("GP_SLAB_TRANSPORT", None)
    port = "29731" if transport == "p2p" else "29732"
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", port, str(script)], env=env,
                       capture_output=True, text=True, timeout=240)
    assert r.returncode == 0 and "MULTI_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


REBALANCE_WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r})
    import numpy as np, torch, torch.distributed as dist
    from gears_b200 import scenes, slabs
    from gears_b200.world import PhysicsWorld
    rank, ws, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    DT = np.float32(1) / np.float32(60)
    n = 60_000
    s = scenes.uniform(n, L=106.0, seed=321, rank_seed=322, floor=False)
    s["acc"][:] = (30.0, 0.0, 0.0)     # everything drifts along +x: fixed faces would empty rank 0
    faces = slabs.bounds(s, ws, "count", halo=6.0)
    part = slabs.partition(s, faces, np.zeros(n, bool))[rank]
    w = PhysicsWorld(device=lr, max_bodies=n + 65536, big_static_only=True)
    w.upload(part)
    slabs.comm_init(w, dist, rank, ws, faces, halo=6.0, device=torch.device("cuda", lr))
    moved = False
    for k in range(12):
        w.step_async(DT, 5)
        st = w.sync()
        moved |= slabs.rebalance(w)        # collective: NCCL all-reduces of the range, the faces and the histogram
    ent, pos, vel = w.download_owned()
    out = [None] * ws if rank == 0 else None
    dist.gather_object((ent, pos, vel, st["steps_inexact"], moved), out, dst=0)
    if rank == 0:
        e = np.concatenate([o[0] for o in out]); p = np.concatenate([o[1] for o in out]); v = np.concatenate([o[2] for o in out])
        order = np.argsort(e, kind="stable")
        assert np.array_equal(e[order], np.arange(n, dtype=np.uint32)), "every body owned exactly once"
        assert sum(o[3] for o in out) == 0 and any(o[4] for o in out)
        counts = np.array([len(o[0]) for o in out], np.float64)
        assert (counts.max() - counts.min()) / counts.mean() < 0.10, counts
        with PhysicsWorld(device=lr) as ref:
            ref.upload(s)
            ref.step_async(DT, 60)
            assert ref.sync()["steps_inexact"] == 0
            rp, rv = ref.download()
        assert np.array_equal(p[order].view(np.uint32), rp.view(np.uint32)), "positions differ from the single-GPU run"
        assert np.array_equal(v[order].view(np.uint32), rv.view(np.uint32)), "velocities differ from the single-GPU run"
        print("REBALANCE_OK", counts)
    w.close()
    dist.destroy_process_group()
""")


def test_two_ranks_rebalance_a_drifting_scene(tmp_path):
    if _gpus() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker_rebalance.py"
    script.write_text(REBALANCE_WORKER.format(root=ROOT))
    env = dict(os.environ)
    env.pop("GP_SLAB_TRANSPORT", None)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29734", str(script)], env=env,
                       capture_output=True, text=True, timeout=240)
    assert r.returncode == 0 and "REBALANCE_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


TIMEOUT_WORKER = textwrap.dedent("""
    import os, sys, time
    sys.path.insert(0, {root!r})
    import numpy as np, torch, torch.distributed as dist
    from gears_b200 import scenes, slabs
    from gears_b200.world import PhysicsWorld, GpError
    rank, ws, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    DT = np.float32(1) / np.float32(60)
    s = scenes.uniform(50_000, L=100.0, seed=311, rank_seed=312)
    rep = slabs.big_static_mask(s, 4.0)
    faces = slabs.bounds(s, ws, "balanced", replicate=rep, halo=6.0)
    part = slabs.partition(s, faces, rep)[rank]
    w = PhysicsWorld(device=lr, max_bodies=int(part["pos"].shape[0] * 1.5) + 65536, big_static_only=True)
    w.upload(part)
    slabs.comm_init(w, dist, rank, ws, faces, halo=6.0, device=torch.device("cuda", lr))
    if rank == 0:
        # the neighbour never steps: the bounded wait expires, the world is dead until it is uploaded again
        w.step_async(DT, 1)
        try:
            w.sync()
            raise SystemExit("the expired wait went unnoticed")
        except GpError as e:
            assert e.status == -9, e
        try:
            w.step_async(DT, 1)
            raise SystemExit("a dead world accepted another step")
        except GpError as e:
            assert e.status == -9, e
        print("TIMEOUT_OK", flush=True)
    else:
        time.sleep(1.5)
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)   # (no orderly tear-down of a half-dead slab pair: the point of the test is the error, not the exit)
""")


def test_missing_neighbour_is_a_fatal_error_not_a_hang(tmp_path):
    """an expired peer-to-peer wait must not append a stale buffer and carry on: GP_ERR_PEER_TIMEOUT,
    no further steps until gp_upload."""
    if _gpus() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker_timeout.py"
    script.write_text(TIMEOUT_WORKER.format(root=ROOT))
    env = dict(os.environ, GP_PEER_TIMEOUT_MS="300")
    env.pop("GP_SLAB_TRANSPORT", None)
    # own process group + bounded wait: a worker that hangs must not survive this test and squat on a GPU
    import signal
    log = tmp_path / "out.log"
    with open(log, "w") as f:
        p = subprocess.Popen([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                              "--master-addr", "127.0.0.1", "--master-port", "29733", str(script)], env=env, stdout=f,
                             stderr=subprocess.STDOUT, start_new_session=True)
        try:
            rc = p.wait(timeout=120)
        except subprocess.TimeoutExpired:
            os.killpg(p.pid, signal.SIGKILL)
            rc = -999
    out = log.read_text()
    assert rc == 0 and "TIMEOUT_OK" in out, f"rc {rc}\n" + out[-4000:]

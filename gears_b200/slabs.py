"""Spatial slab decomposition along x (BASELINE config 4): host-side helpers over the C ABI.

partition()   - split a scene (dict of numpy arrays, see scenes.py) into per-rank scenes: a rank gets the bodies
                whose AABB centre x lies in its slab; big static bodies (a floor) are replicated on every rank.
bounds()      - slab faces: equal-width, or equal-count quantiles of the body centres.
LocalSlabs    - n worlds in ONE process (gp_comm_init_local + gp_step_group): peer copies instead of NCCL; the
                worlds may share a device, which is how the exchange logic is tested on one GPU.
comm_init()   - one process per GPU: rank 0 makes the NCCL unique id, torch.distributed broadcasts it, every rank
                attaches its world (gp_comm_init).  torch.distributed is plumbing only; the halo/migration traffic
                goes through ncclSend/ncclRecv inside the library, on the step stream.
The reference has no distributed runtime; nothing here mirrors reference code.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .world import GpError, PhysicsWorld

INF = float("inf")


def aabb_offsets_x(scene: dict) -> tuple[np.ndarray, np.ndarray]:
    """(min, max) x offset of every body's world AABB from its position: the x component of the component-wise min / max
    over the 8 corners rot * (corner * scale), with the operation order of the upload kernel (k_pack / corner_offsets,
    reference physics.rs:93-119), in float32 -- so the host assigns a rotated or scaled body to the slab the device's
    ownership test (k_keys<true>) will agree with."""
    f = np.float32
    bmin, bmax = scene["box_min"].astype(f), scene["box_max"].astype(f)
    n = bmin.shape[0]
    rot, sc = scene.get("rot"), scene.get("scale")
    if rot is None and sc is None:
        return bmin[:, 0].copy(), bmax[:, 0].copy()
    q = np.tile(np.array([0, 0, 0, 1], f), (n, 1)) if rot is None else rot.astype(f)
    sc = np.ones((n, 3), f) if sc is None else sc.astype(f)
    qv, qs = q[:, :3], q[:, 3:4]

    def cross(a, b):
        return np.stack([a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1], a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2],
                         a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]], axis=1).astype(f)
    lo = hi = None
    for i in range(8):
        c = np.stack([bmax[:, 0] if i & 1 else bmin[:, 0], bmax[:, 1] if i & 2 else bmin[:, 1],
                      bmax[:, 2] if i & 4 else bmin[:, 2]], axis=1)
        v = (c * sc).astype(f)
        t = (cross(qv, v) + v * qs).astype(f)
        r = (cross(qv, t) * f(2.0) + v).astype(f)
        lo = r[:, 0] if lo is None else np.minimum(lo, r[:, 0])
        hi = r[:, 0] if hi is None else np.maximum(hi, r[:, 0])
    return lo.astype(f), hi.astype(f)


def centres_x(scene: dict) -> np.ndarray:
    """AABB centre x as the kernels compute it: ((off_min + pos) + (off_max + pos)) * 0.5 in float32, rotation and
    scale included."""
    olo, ohi = aabb_offsets_x(scene)
    lo = (olo + scene["pos"][:, 0]).astype(np.float32)
    hi = (ohi + scene["pos"][:, 0]).astype(np.float32)
    return ((lo + hi) * np.float32(0.5)).astype(np.float32)


def bounds(scene: dict, nranks: int, mode: str = "balanced", lo: float | None = None, hi: float | None = None,
           replicate: np.ndarray | None = None, halo: float = 4.0) -> np.ndarray:
    """nranks+1 faces; the outer two are -inf / +inf.
    mode 'count': equal owned-body counts; 'balanced': equal owned + halo-copy counts (inner slabs mirror two faces,
    the end slabs one, so inner slabs get a little less to own); 'width': equal widths."""
    cx = centres_x(scene)
    if replicate is not None:
        cx = cx[~replicate]
    if nranks == 1:
        return np.array([-INF, INF], np.float32)
    if mode == "width":
        a = float(cx.min()) if lo is None else lo
        b = float(cx.max()) if hi is None else hi
        inner = a + (b - a) * np.arange(1, nranks) / nranks
    else:
        n = len(cx)
        faces_per_rank = np.array([1.0] + [2.0] * (nranks - 2) + [1.0]) if nranks > 1 else np.array([0.0])
        g = 0.0
        if mode == "balanced" and n:
            ext = float((scene["box_max"] - scene["box_min"])[:, 0].max()) if replicate is None else \
                float((scene["box_max"] - scene["box_min"])[~replicate, 0].max())
            span = max(1e-6, float(cx.max() - cx.min()))
            g = n * min(0.25, (halo + 1.5 * ext) / span)  # halo copies per face at the average density
        total = (n + faces_per_rank.sum() * g) / nranks   # owned + mirrored per rank
        own = np.maximum(total - faces_per_rank * g, 0.05 * n / nranks)
        q = np.cumsum(own)[:-1] / own.sum()
        inner = np.quantile(cx.astype(np.float64), q)
    return np.concatenate([[-INF], inner, [INF]]).astype(np.float32)


def big_static_mask(scene: dict, threshold: float) -> np.ndarray:
    ext = (scene["box_max"] - scene["box_min"]).max(axis=1)
    return (scene["is_static"] != 0) & (ext > threshold)


def partition(scene: dict, faces: np.ndarray, replicate: np.ndarray) -> list[dict]:
    """Per-rank scenes.  entity = global body index (or the scene's entity ids), rank = global iteration rank."""
    n = scene["pos"].shape[0]
    cx = centres_x(scene)
    ent = scene.get("entity")
    ent = np.arange(n, dtype=np.uint32) if ent is None else np.asarray(ent, np.uint32)
    rank = scene.get("rank")
    rank = np.arange(n, dtype=np.uint32) if rank is None else np.asarray(rank, np.uint32)
    owner = np.clip(np.searchsorted(faces[1:-1], cx, side="right"), 0, len(faces) - 2)
    out = []
    for r in range(len(faces) - 1):
        idx = np.nonzero((owner == r) | replicate)[0]
        s = {}
        for k, v in scene.items():
            s[k] = None if v is None else np.ascontiguousarray(v[idx])
        s["entity"] = np.ascontiguousarray(ent[idx])
        s["rank"] = np.ascontiguousarray(rank[idx])  # GLOBAL ranks (distinct, not dense: only the order matters)
        out.append(s)
    return out


def slab_config(r: int, nranks: int, faces: np.ndarray, halo: float = 0.0, max_exchange: int = 0,
                verify: bool = True) -> capi.GpSlabConfig:
    c = capi.GpSlabConfig()
    c.struct_size = C.sizeof(capi.GpSlabConfig)
    c.rank, c.nranks = r, nranks
    c.x_lo, c.x_hi = float(faces[r]), float(faces[r + 1])
    c.halo = halo
    c.max_exchange = max_exchange
    c.flags = 0 if verify else capi.GP_SLAB_NO_VERIFY
    return c


def world_kwargs(n_local: int, faces: np.ndarray, r: int, halo_guess: float, scene: dict, headroom: float = 1.5) -> dict:
    """Capacity and grid bounds for a slab world: room for halo copies / immigrants, grid covering slab + halo."""
    kw = dict(max_bodies=int(n_local * headroom) + 65536)
    return kw


class LocalSlabs:
    """n slab worlds in one process (devices[i] may repeat)."""

    def __init__(self, scene: dict, nranks: int, devices=None, halo: float = 0.0, max_exchange: int = 0,
                 big_threshold: float | None = None, mode: str = "balanced", verify: bool = True, **world_kw):
        self.lib = capi.lib()
        devices = devices or [0] * nranks
        ext = (scene["box_max"] - scene["box_min"]).max(axis=1)
        if big_threshold is None:
            dyn = ext[scene["is_static"] == 0]
            big_threshold = 4.0 * float(dyn.max()) if len(dyn) else 0.0
        rep = big_static_mask(scene, big_threshold)
        self.faces = bounds(scene, nranks, mode, replicate=rep)
        parts = partition(scene, self.faces, rep)
        self.worlds = []
        n_total = scene["pos"].shape[0]
        for r, part in enumerate(parts):
            kw = dict(world_kw)
            kw.setdefault("max_bodies", int(part["pos"].shape[0] * 1.5) + max(65536, n_total // 8))
            kw.setdefault("big_static_only", True)
            w = PhysicsWorld(device=devices[r], **kw)
            w.upload(part)
            self.worlds.append(w)
        cfgs = (capi.GpSlabConfig * nranks)(*[slab_config(r, nranks, self.faces, halo, max_exchange, verify)
                                               for r in range(nranks)])
        hs = (C.c_void_p * nranks)(*[w._h for w in self.worlds])
        st = self.lib.gp_comm_init_local(hs, nranks, cfgs)
        if st < 0:
            msg = "; ".join(m for m in (self.lib.gp_last_error(w._h).decode() for w in self.worlds) if m)
            for w in self.worlds:
                w.close()
            raise GpError(st, msg)
        self._hs = hs
        self.n = nranks

    def step(self, dt: float, nsteps: int = 1, damp=None):
        d = None if damp is None else np.ascontiguousarray(damp, np.float32)
        st = self.lib.gp_step_group(self._hs, self.n, C.c_float(dt), None if d is None else d.ctypes.data_as(C.c_void_p),
                                    nsteps)
        if st < 0:
            raise GpError(st, "; ".join(self.lib.gp_last_error(w._h).decode() for w in self.worlds))

    def sync(self) -> list[dict]:
        return [w.sync() for w in self.worlds]

    def rebalance(self, max_shift: float = 0.0) -> bool:
        """Move the faces towards equal owned counts (between two steps); True if a face moved."""
        moved = C.c_uint32()
        st = self.lib.gp_slab_rebalance_local(self._hs, self.n, C.c_float(max_shift), C.byref(moved))
        if st < 0:
            raise GpError(st, "; ".join(self.lib.gp_last_error(w._h).decode() for w in self.worlds))
        return bool(moved.value)

    def owned_counts(self) -> list[int]:
        out = []
        for w in self.worlds:
            n = C.c_uint32()
            w._check(self.lib.gp_owned_count(w._h, C.byref(n)))
            out.append(int(n.value))
        return out

    def download(self):
        """(entity, pos, vel) of all owned bodies, sorted by entity id."""
        es, ps, vs, us = [], [], [], []
        for w in self.worlds:
            e, p, v = w.download_owned()
            es.append(e), ps.append(p), vs.append(v), us.append(w.last_unverified)
        e, p, v, u = np.concatenate(es), np.concatenate(ps), np.concatenate(vs), np.concatenate(us)
        o = np.argsort(e, kind="stable")
        self.last_unverified = u[o]  # per body: the last step could not prove it independent of unseen bodies
        return e[o], p[o], v[o]

    def close(self):
        for w in self.worlds:
            w.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def rebalance(world: PhysicsWorld, max_shift: float = 0.0) -> bool:
    """One process per GPU: COLLECTIVE re-balancing of the slab faces (every rank calls it at the same step)."""
    moved = C.c_uint32()
    st = capi.lib().gp_slab_rebalance(world._h, C.c_float(max_shift), C.byref(moved))
    if st < 0:
        raise GpError(st, capi.lib().gp_last_error(world._h).decode())
    return bool(moved.value)


def comm_init(world: PhysicsWorld, dist, rank: int, nranks: int, faces: np.ndarray, halo: float = 0.0,
              max_exchange: int = 0, verify: bool = True, device=None):
    """One process per GPU: broadcast the NCCL unique id through torch.distributed, attach this rank's world."""
    import torch
    lib = capi.lib()
    uid = (C.c_uint8 * capi.GP_UNIQUE_ID_BYTES)()
    if rank == 0:
        st = lib.gp_comm_unique_id(uid)
        if st < 0:
            raise GpError(st, lib.gp_last_error(None).decode())
    t = torch.tensor(list(bytes(uid)), dtype=torch.uint8)
    if device is not None:
        t = t.to(device)
    dist.broadcast(t, src=0)
    raw = bytes(t.cpu().tolist())
    uid = (C.c_uint8 * capi.GP_UNIQUE_ID_BYTES)(*raw)
    cfg = slab_config(rank, nranks, faces, halo, max_exchange, verify)
    st = lib.gp_comm_init(world._h, uid, C.byref(cfg))
    if st < 0:
        raise GpError(st, lib.gp_last_error(world._h).decode())

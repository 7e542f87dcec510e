"""PhysicsWorld — thin Python host mirror over the C ABI, used by tests and bench.py.

It plays the role of the `gears-physics-cuda` Rust crate's safe wrapper (rust/gears-physics-cuda):
own a gp_world handle, mirror the ECS component arrays into it, call the step, read results back.
No physics is computed here and there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


class GpError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"gp_status {status}: {msg}")
        self.status = status


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a, cols=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float32)
    if cols is not None and a.size:
        assert a.shape[-1] == cols
    return a


class PhysicsWorld:
    def __init__(self, device: int = 0, max_bodies: int = 0, max_pairs: int = 0, margin_init: float = 0.0,
                 margin_max: float = 0.0, cell_size: float = 0.0, max_big: int = 0, record_pairs: bool = False,
                 no_retry: bool = False, profile: bool = False, grid_min=None, grid_max=None,
                 big_static_only: bool = False):
        self._lib = capi.lib()
        cfg = capi.GpConfig()
        cfg.struct_size = C.sizeof(capi.GpConfig)
        cfg.device = device
        cfg.max_bodies = max_bodies
        cfg.max_pairs = max_pairs
        cfg.margin_init = margin_init
        cfg.margin_max = margin_max
        cfg.cell_size = cell_size
        cfg.max_big = max_big
        cfg.flags = (capi.GP_FLAG_RECORD_PAIRS if record_pairs else 0) | (capi.GP_FLAG_NO_RETRY if no_retry else 0) | \
            (capi.GP_FLAG_PROFILE if profile else 0) | (capi.GP_FLAG_BIG_STATIC_ONLY if big_static_only else 0)
        if grid_min is not None:
            cfg.grid_min[:] = [float(x) for x in grid_min]
            cfg.grid_max[:] = [float(x) for x in grid_max]
        h = C.c_void_p()
        st = self._lib.gp_create(C.byref(cfg), C.byref(h))
        if st != 0:
            raise GpError(st, self._lib.gp_last_error(None).decode())
        self._h = h
        self.n = 0

    # -- helpers --
    def _check(self, st: int) -> int:
        if st < 0:
            raise GpError(st, self._lib.gp_last_error(self._h).decode())
        return st

    def close(self):
        if getattr(self, "_h", None):
            self._lib.gp_destroy(self._h)
            self._h = None
            for p in getattr(self, "_pinned", []):
                self._lib.gp_host_free(p)
            self._pinned = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- mirror --
    def upload(self, scene: dict):
        n = int(scene["pos"].shape[0])
        keep = dict(
            pos=_f32(scene["pos"], 3), rot=_f32(scene.get("rot"), 4), scale=_f32(scene.get("scale"), 3),
            vel=_f32(scene["vel"], 3), acc=_f32(scene.get("acc"), 3), bmin=_f32(scene["box_min"], 3),
            bmax=_f32(scene["box_max"], 3), mass=_f32(scene["mass"]), rest=_f32(scene["restitution"]),
            stat=np.ascontiguousarray(scene["is_static"], dtype=np.uint8),
            ent=None if scene.get("entity") is None else np.ascontiguousarray(scene["entity"], dtype=np.uint32),
            rank=None if scene.get("rank") is None else np.ascontiguousarray(scene["rank"], dtype=np.uint32))
        self._check(self._lib.gp_upload(self._h, n, _ptr(keep["pos"]), _ptr(keep["rot"]), _ptr(keep["scale"]),
                                        _ptr(keep["vel"]), _ptr(keep["acc"]), _ptr(keep["bmin"]), _ptr(keep["bmax"]),
                                        _ptr(keep["mass"]), _ptr(keep["rest"]), _ptr(keep["stat"]), _ptr(keep["ent"]),
                                        _ptr(keep["rank"])))
        self.n = n

    def set_state(self, pos=None, vel=None, acc=None):
        pos, vel, acc = _f32(pos, 3), _f32(vel, 3), _f32(acc, 3)
        self._check(self._lib.gp_set_state(self._h, _ptr(pos), _ptr(vel), _ptr(acc)))
        self._check(self._lib.gp_sync(self._h, None))

    def update_dirty(self, idx, pos=None, vel=None, acc=None):
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        pos, vel, acc = _f32(pos, 3), _f32(vel, 3), _f32(acc, 3)
        self._check(self._lib.gp_update_dirty(self._h, len(idx), _ptr(idx), _ptr(pos), _ptr(vel), _ptr(acc)))

    def update_shape(self, idx, box_min, box_max, mass, restitution, is_static, rot=None, scale=None):
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        a = [_f32(box_min, 3), _f32(box_max, 3), _f32(rot, 4), _f32(scale, 3), _f32(mass), _f32(restitution),
             np.ascontiguousarray(is_static, dtype=np.uint8)]
        self._check(self._lib.gp_update_shape(self._h, len(idx), _ptr(idx), *[_ptr(x) for x in a]))

    # -- step --
    def step(self, dt: float, damp=None) -> dict:
        st = capi.GpStepStats()
        d = None if damp is None else _f32(damp)
        rc = self._check(self._lib.gp_step(self._h, C.c_float(dt), _ptr(d), C.byref(st)))
        out = st.as_dict()
        out["status"] = rc
        return out

    def step_async(self, dt: float, nsteps: int = 1, damp=None):
        d = None if damp is None else _f32(damp)
        self._check(self._lib.gp_step_async(self._h, C.c_float(dt), _ptr(d), nsteps))

    def sync(self) -> dict:
        st = capi.GpStepStats()
        self._check(self._lib.gp_sync(self._h, C.byref(st)))
        return st.as_dict()

    def last_step_ms(self) -> float:
        ms = C.c_float()
        self._check(self._lib.gp_last_step_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def kernel_times(self) -> dict:
        """Per-kernel-group device ms summed over the steps since the last call (needs profile=True)."""
        kt = capi.GpKernelTimes()
        self._check(self._lib.gp_kernel_times_read(self._h, C.byref(kt)))
        return dict(steps=int(kt.steps), ms={g: float(kt.ms[i]) for i, g in enumerate(capi.KERNEL_GROUPS)},
                    launches={g: int(kt.launches[i]) for i, g in enumerate(capi.KERNEL_GROUPS)})

    def kernel_launches(self) -> int:
        return int(self._lib.gp_kernel_launches(self._h))

    def cap_velocity(self, idx=None):
        """RigidBody::cap_velocity (physics.rs:506-520) on the device for bodies idx (None: all)."""
        ix = None if idx is None else np.ascontiguousarray(idx, dtype=np.uint32)
        self._check(self._lib.gp_cap_velocity(self._h, 0 if ix is None else len(ix), _ptr(ix)))

    # -- pipelined frames (gp_frame_submit / gp_frame_wait) --
    def host_alloc(self, shape, dtype=np.float32) -> np.ndarray:
        """A numpy array over PINNED host memory (gp_host_alloc); freed when the world closes."""
        count = int(np.prod(shape))
        nbytes = count * np.dtype(dtype).itemsize
        p = self._lib.gp_host_alloc(max(1, nbytes))
        if not p:
            raise MemoryError("gp_host_alloc failed")
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append(p)
        buf = (C.c_char * max(1, nbytes)).from_address(p)
        return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)

    def frame_submit(self, pos, vel, acc, dt: float, state6_out, damp=None):
        """Enqueue one frame: H2D of pos / vel / acc (n x 3 each, upload order) -> state refresh + step -> D2H of
        (pos.xyz, vel.xyz) rows into state6_out (n x 6).  Returns at once; buffers should be pinned."""
        for a in (pos, vel, acc, state6_out):
            assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
        d = None if damp is None else _f32(damp)
        self._check(self._lib.gp_frame_submit(self._h, _ptr(pos), _ptr(vel), _ptr(acc), C.c_float(dt), _ptr(d),
                                              _ptr(state6_out)))

    def frame_wait(self, max_in_flight: int = 0) -> dict:
        st = capi.GpStepStats()
        self._check(self._lib.gp_frame_wait(self._h, max_in_flight, C.byref(st)))
        return st.as_dict()

    # -- results --
    def download(self, pos_out=None, vel_out=None):
        pos = np.empty((self.n, 3), np.float32) if pos_out is None else pos_out
        vel = np.empty((self.n, 3), np.float32) if vel_out is None else vel_out
        self._check(self._lib.gp_download(self._h, _ptr(pos), _ptr(vel)))
        return pos, vel

    def download_owned(self, out=None):
        """(entity, pos, vel) of the bodies this world owns now (slab worlds: after migration; no halo copies), in
        device order.  out = (ent, pos, vel, unverified) buffers to fill (e.g. pinned, sized for max_bodies)."""
        n = C.c_uint32()
        self._check(self._lib.gp_owned_count(self._h, C.byref(n)))
        k = max(1, n.value)
        if out is None:
            ent, pos, vel = np.zeros(k, np.uint32), np.zeros((k, 3), np.float32), np.zeros((k, 3), np.float32)
            unv = np.zeros(k, np.uint8)
        else:
            ent, pos, vel, unv = out
            assert len(ent) >= k
        self._check(self._lib.gp_download_owned(self._h, _ptr(ent), _ptr(pos), _ptr(vel), _ptr(unv)))
        self.last_unverified = unv[: n.value]
        return ent[: n.value], pos[: n.value], vel[: n.value]

    def set_state_owned(self, pos=None, vel=None, acc=None):
        """Rows of the last download_owned written back (no step in between)."""
        k = next(len(a) for a in (pos, vel, acc) if a is not None)
        pos, vel, acc = _f32(pos, 3), _f32(vel, 3), _f32(acc, 3)
        self._check(self._lib.gp_set_state_owned(self._h, k, _ptr(pos), _ptr(vel), _ptr(acc)))

    # -- next rows --
    def instances(self, rot=None, scale=None) -> np.ndarray:
        """InstanceRaw of every body (n, 25): 4x4 model + 3x3 normal, column-major (instance.rs:22-42)."""
        rot, scale = _f32(rot, 4), _f32(scale, 3)
        out = np.zeros((max(1, self.n), 25), np.float32)
        self._check(self._lib.gp_instances(self._h, _ptr(rot), _ptr(scale), _ptr(out)))
        return out[: self.n]

    def raycast(self, origins, dirs, targets=None) -> np.ndarray:
        """hits[r, t]: Weapon::ray_intersects_aabb (interactive.rs:94-158) of ray r against body targets[t]."""
        origins, dirs = _f32(origins, 3), _f32(dirs, 3)
        t = None if targets is None else np.ascontiguousarray(targets, dtype=np.uint32)
        nt = self.n if t is None else len(t)
        hits = np.zeros((len(origins), max(1, nt)), np.uint8)
        self._check(self._lib.gp_raycast(self._h, len(origins), _ptr(origins), _ptr(dirs), nt, _ptr(t), _ptr(hits)))
        return hits[:, :nt].astype(bool)

    def obstacle_grid(self, cell_size: float, origin, dims, idx=None):
        """(bitmap as bool array [z, y, x], bounds(6)): PathfindingGrid::add_obstacle (pathfinding.rs:296-338)."""
        ix = None if idx is None else np.ascontiguousarray(idx, dtype=np.uint32)
        o = np.ascontiguousarray(origin, dtype=np.int32)
        d = np.ascontiguousarray(dims, dtype=np.uint32)
        cells = int(d[0]) * int(d[1]) * int(d[2])
        bits = np.zeros((cells + 31) // 32, np.uint32)
        bounds = np.zeros(6, np.int32)
        self._check(self._lib.gp_obstacle_grid(self._h, 0 if ix is None else len(ix), _ptr(ix), C.c_float(cell_size),
                                               _ptr(o), _ptr(d), _ptr(bits), _ptr(bounds)))
        flat = np.unpackbits(bits.view(np.uint8), bitorder="little")[:cells].astype(bool)
        return flat.reshape(int(d[2]), int(d[1]), int(d[0])), bounds

    def pairs(self, which: int) -> np.ndarray:
        n = C.c_uint64()
        self._check(self._lib.gp_pairs(self._h, which, None, 0, C.byref(n)))
        out = np.zeros((n.value, 2), np.uint32)
        if n.value:
            self._check(self._lib.gp_pairs(self._h, which, _ptr(out), n.value, C.byref(n)))
        return out


def damping_factors(dt: float) -> np.ndarray:
    out = np.zeros(3, np.float32)
    capi.lib().gp_damping_factors(C.c_float(dt), out.ctypes.data_as(capi.f32p))
    return out

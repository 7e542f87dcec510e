"""ctypes loader for the CPU oracle (oracle/gears_oracle.cpp).

TEST INFRASTRUCTURE ONLY.  May be imported from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Nothing under gears_b200/ imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libgears_oracle.so")
_lib = None

f32p = C.POINTER(C.c_float)
u32p = C.POINTER(C.c_uint32)
u8p = C.POINTER(C.c_uint8)


class StepOut(C.Structure):
    _fields_ = [("n_snapshot", C.c_uint64), ("n_sweep", C.c_uint64), ("n_static_static", C.c_uint64)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "gears_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.go_step.restype = C.c_int
        _lib.go_intersects.restype = C.c_int
        _lib.go_check_and_resolve.restype = C.c_int
        _lib.go_max_threads.restype = C.c_int
    return _lib


def _f(a, cols=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float32)
    if cols is not None:
        assert a.shape[-1] == cols, (a.shape, cols)
    return a


def _p(a, typ):
    return None if a is None else a.ctypes.data_as(typ)


def step(scene: dict, dt: float, steps: int = 1, mode: int = 0, threads: int = 1, want_snapshot: bool = True,
         pair_cap: int | None = None, want_pairs: bool = True):
    """Run `steps` physics steps on a copy of `scene` (dict of numpy arrays, see gears_b200.scenes).

    Returns dict(pos, vel, snapshot, sweep, static_static); pair arrays are (m,2) uint32 of caller
    indices in the reference's iteration order (pairs of the LAST step).
    mode 0 = the reference's brute-force algorithm, 1 = grid-accelerated (bit-identical).
    """
    n = int(scene["pos"].shape[0])
    pos = _f(scene["pos"], 3).copy()
    vel = _f(scene["vel"], 3).copy()
    rot = _f(scene.get("rot"), 4)
    scale = _f(scene.get("scale"), 3)
    acc = _f(scene["acc"], 3)
    bmin = _f(scene["box_min"], 3)
    bmax = _f(scene["box_max"], 3)
    mass = _f(scene["mass"])
    rest = _f(scene["restitution"])
    stat = np.ascontiguousarray(scene["is_static"], dtype=np.uint8)
    rank = scene.get("rank")
    rank = None if rank is None else np.ascontiguousarray(rank, dtype=np.uint32)
    cap = pair_cap if pair_cap is not None else max(1024, 64 * n)
    snap = np.zeros((cap, 2), np.uint32)
    sweep = np.zeros((cap, 2), np.uint32)
    ss = np.zeros((cap, 2), np.uint32)
    out = StepOut()
    rc = lib().go_step(C.c_uint32(n), _p(pos, f32p), _p(rot, f32p), _p(scale, f32p), _p(vel, f32p), _p(acc, f32p),
                       _p(bmin, f32p), _p(bmax, f32p), _p(mass, f32p), _p(rest, f32p), _p(stat, u8p),
                       _p(rank, u32p), C.c_float(dt), C.c_int(steps), C.c_int(mode), C.c_int(threads),
                       C.c_int(1 if want_snapshot else 0), _p(snap, u32p), C.c_uint64(cap), _p(sweep, u32p),
                       C.c_uint64(cap), _p(ss, u32p), C.c_uint64(cap), C.byref(out))
    if rc != 0:
        raise ValueError("oracle: rank is not a permutation of 0..n-1")
    if want_pairs and max(out.n_snapshot, out.n_sweep, out.n_static_static) > cap:
        raise OverflowError("oracle pair capacity too small")
    if not want_pairs:  # (timing runs: the pair lists were counted, not stored)
        return dict(pos=pos, vel=vel, n_sweep=int(out.n_sweep))
    return dict(pos=pos, vel=vel, snapshot=snap[: out.n_snapshot].copy(), sweep=sweep[: out.n_sweep].copy(),
                static_static=ss[: out.n_static_static].copy())


def update_pos(pos, vel, acc, is_static: bool, dt: float):
    pos = _f(pos, 3).copy()
    vel = _f(vel, 3).copy()
    acc = _f(acc, 3)
    lib().go_update_pos(_p(pos, f32p), _p(vel, f32p), _p(acc, f32p), C.c_int(int(is_static)), C.c_float(dt))
    return pos, vel


def world_aabb(bmin, bmax, pos, rot=None, scale=None):
    lo = np.zeros(3, np.float32)
    hi = np.zeros(3, np.float32)
    bmin, bmax, pos, rot, scale = _f(bmin, 3), _f(bmax, 3), _f(pos, 3), _f(rot, 4), _f(scale, 3)
    lib().go_world_aabb(_p(bmin, f32p), _p(bmax, f32p), _p(pos, f32p), _p(rot, f32p), _p(scale, f32p),
                        _p(lo, f32p), _p(hi, f32p))
    return lo, hi


def intersects(a_min, a_max, a_pos, b_min, b_max, b_pos, a_rot=None, a_scale=None, b_rot=None, b_scale=None) -> bool:
    arrs = [_f(x) for x in (a_min, a_max, a_pos, a_rot, a_scale, b_min, b_max, b_pos, b_rot, b_scale)]
    return bool(lib().go_intersects(*[_p(x, f32p) for x in arrs]))


def check_and_resolve(a: dict, b: dict):
    """a, b: dict(pos, vel, box_min, box_max, mass, restitution, is_static[, rot, scale]).
    Returns (hit, a_pos, a_vel, b_pos, b_vel)."""
    ap, av = _f(a["pos"], 3).copy(), _f(a["vel"], 3).copy()
    bp, bv = _f(b["pos"], 3).copy(), _f(b["vel"], 3).copy()
    keep = [_f(a.get("rot")), _f(a.get("scale")), _f(a["box_min"]), _f(a["box_max"]), _f(b.get("rot")),
            _f(b.get("scale")), _f(b["box_min"]), _f(b["box_max"])]
    hit = lib().go_check_and_resolve(
        _p(ap, f32p), _p(av, f32p), _p(keep[0], f32p), _p(keep[1], f32p), _p(keep[2], f32p), _p(keep[3], f32p),
        C.c_float(a["mass"]), C.c_float(a["restitution"]), C.c_int(int(a["is_static"])),
        _p(bp, f32p), _p(bv, f32p), _p(keep[4], f32p), _p(keep[5], f32p), _p(keep[6], f32p), _p(keep[7], f32p),
        C.c_float(b["mass"]), C.c_float(b["restitution"]), C.c_int(int(b["is_static"])))
    return bool(hit), ap, av, bp, bv


def cap_velocity(vel):
    v = _f(vel, 3).copy()
    lib().go_cap_velocity(_p(v, f32p))
    return v


def damping_factors(dt: float):
    out = np.zeros(3, np.float32)
    lib().go_damping_factors(C.c_float(dt), _p(out, f32p))
    return out


def instance_raw(pos, rot, scale) -> np.ndarray:
    """Instance::to_raw (instance.rs:22-31): 16 model + 9 normal floats, column-major."""
    out = np.zeros(25, np.float32)
    pos, rot, scale = _f(pos, 3), _f(rot, 4), _f(scale, 3)
    lib().go_instance_raw(_p(pos, f32p), _p(rot, f32p), _p(scale, f32p), _p(out, f32p))
    return out


def ray_intersects_aabb(origin, direction, pos, bmin, bmax) -> bool:
    a = [_f(x, 3) for x in (origin, direction, pos, bmin, bmax)]
    lib().go_ray_intersects_aabb.restype = C.c_int
    return bool(lib().go_ray_intersects_aabb(*[_p(x, f32p) for x in a]))


def obstacle_cells(pos, bmin, bmax, cell_size: float):
    lo, hi = np.zeros(3, np.int32), np.zeros(3, np.int32)
    a = [_f(x, 3) for x in (pos, bmin, bmax)]
    lib().go_obstacle_cells(*[_p(x, f32p) for x in a], C.c_float(cell_size), lo.ctypes.data_as(C.c_void_p),
                            hi.ctypes.data_as(C.c_void_p))
    return lo, hi


def max_threads() -> int:
    return int(lib().go_max_threads())

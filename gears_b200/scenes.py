"""Deterministic synthetic scenes for the physics step.

A scene is a dict of numpy arrays laid out exactly as the C ABI's gp_upload takes them:
  pos (n,3) f32, rot (n,4) f32 or None [cgmath order v.x v.y v.z s], scale (n,3) f32 or None,
  vel (n,3), acc (n,3), box_min (n,3), box_max (n,3), mass (n,), restitution (n,),
  is_static (n,) u8, entity (n,) u32, rank (n,) u32 or None.
`rank[i]` is the slot of body i in the reference's iteration order, i.e. the index of its entity
in the Vec that World::get_entities_with_component returns (reference
crates/gears-ecs/src/lib.rs:422-433); None means identity.

Random numbers come from a counter-based splitmix64 so C++, CUDA and Python can reproduce them:
  h(seed,i,c) = mix(seed + 0x9E3779B97F4A7C15*(8i+c+1)),  u = (h>>40) * 2^-24  (exact in f32).
"""
from __future__ import annotations

import numpy as np

_GOLD = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def _mix(z: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = z ^ (z >> np.uint64(30))
        z = z * _M1
        z = z ^ (z >> np.uint64(27))
        z = z * _M2
        z = z ^ (z >> np.uint64(31))
    return z


def hash64(seed: int, i: np.ndarray, c: int) -> np.ndarray:
    i = np.asarray(i, dtype=np.uint64)
    with np.errstate(over="ignore"):
        return _mix(np.uint64(seed) + _GOLD * (np.uint64(8) * i + np.uint64(c + 1)))


def uniform01(seed: int, i: np.ndarray, c: int) -> np.ndarray:
    """u in [0,1) with 24 random bits, exactly representable in float32."""
    return ((hash64(seed, i, c) >> np.uint64(40)).astype(np.float32)) * np.float32(2.0 ** -24)


def shuffled_rank(n: int, seed: int) -> np.ndarray:
    """A seeded permutation standing in for DashMap's per-process hash order:
    slot order = bodies sorted by h(seed,i,7) (ties by index); returns rank[i] = slot of body i."""
    order = np.argsort(hash64(seed, np.arange(n), 7), kind="stable")
    rank = np.empty(n, np.uint32)
    rank[order] = np.arange(n, dtype=np.uint32)
    return rank


def _empty(n: int) -> dict:
    return dict(
        pos=np.zeros((n, 3), np.float32), rot=None, scale=None,
        vel=np.zeros((n, 3), np.float32), acc=np.zeros((n, 3), np.float32),
        box_min=np.full((n, 3), -0.5, np.float32), box_max=np.full((n, 3), 0.5, np.float32),
        mass=np.ones(n, np.float32), restitution=np.full(n, 0.2, np.float32),
        is_static=np.zeros(n, np.uint8), entity=np.arange(n, dtype=np.uint32), rank=None)


def _set_static(s: dict, i: int, pos, half) -> None:
    s["pos"][i] = pos
    s["box_min"][i] = -np.asarray(half, np.float32)
    s["box_max"][i] = np.asarray(half, np.float32)
    s["mass"][i] = 0.0  # RigidBody::new_static, physics.rs:379-388
    s["is_static"][i] = 1
    s["vel"][i] = 0
    s["acc"][i] = 0


def physics_example(order: str = "ascending", seed: int = 2000) -> dict:
    """C1: the 52 rigid bodies of reference examples/src/physics.rs:68-226 (entity ids 2..53)."""
    crates = [(-20, 15, -20), (-20, 18, -10), (-20, 21, 0), (-20, 24, 10), (-20, 27, 20),
              (-10, 12, -20), (-10, 16, -10), (-10, 20, 0), (-10, 24, 10), (-10, 28, 20),
              (0, 10, -20), (0, 15, -10), (0, 20, 0), (0, 25, 10), (0, 30, 20),
              (10, 14, -20), (10, 17, -10), (10, 21, 0), (10, 25, 10), (10, 29, 20),
              (20, 13, -20), (20, 19, -10), (20, 22, 0), (20, 26, 10), (20, 31, 20)]
    spheres = [(-25, 35, -15, 2.0), (-18, 40, 5, 1.5), (-12, 28, -25, 3.0), (-5, 45, 15, 1.0), (0, 50, -5, 2.5),
               (5, 32, -18, 1.8), (12, 38, 8, 2.2), (18, 42, -12, 1.3), (25, 36, 18, 2.8), (-22, 48, 22, 1.6),
               (22, 44, -22, 1.4), (-8, 52, -8, 1.2), (8, 34, 8, 2.4), (15, 46, -2, 1.7), (-15, 30, 12, 2.1)]
    proj = [(-40, 10, 0, 35, 8, 0), (40, 12, 0, -30, 5, 5), (0, 15, -40, 0, 6, 32), (0, 18, 40, 5, 4, -28),
            (-30, 14, -30, 25, 7, 25), (30, 16, 30, -22, 6, -22)]
    high = [(5, 60, 5), (-5, 65, -5), (3, 70, -3), (-3, 75, 3), (0, 80, 0)]
    n = 1 + len(crates) + len(spheres) + len(proj) + len(high)
    s = _empty(n)
    s["entity"] = np.arange(2, 2 + n, dtype=np.uint32)
    _set_static(s, 0, (0.0, -1.0, 0.0), (50.0, 0.1, 50.0))
    k = 1
    for (x, y, z) in crates:
        s["pos"][k] = (x, y, z); s["mass"][k] = 60.0
        s["box_min"][k] = -1.5; s["box_max"][k] = 1.5
        k += 1
    for (x, y, z, m) in spheres:
        s["pos"][k] = (x, y, z); s["mass"][k] = m; s["acc"][k] = (0, -10, 0)
        s["box_min"][k] = -1.0; s["box_max"][k] = 1.0
        k += 1
    for (x, y, z, vx, vy, vz) in proj:
        s["pos"][k] = (x, y, z); s["vel"][k] = (vx, vy, vz); s["mass"][k] = 1.5; s["acc"][k] = (0, -10, 0)
        s["box_min"][k] = -1.0; s["box_max"][k] = 1.0
        k += 1
    for (x, y, z) in high:
        s["pos"][k] = (x, y, z); s["mass"][k] = 55.0
        s["box_min"][k] = -1.5; s["box_max"][k] = 1.5
        k += 1
    assert k == n == 52
    if order == "shuffle":
        s["rank"] = shuffled_rank(n, seed)
    elif order == "reversed":
        s["rank"] = np.arange(n - 1, -1, -1, dtype=np.uint32)
    return s


def uniform(n: int = 100_000, L: float = 126.0, seed: int = 1001, rank_seed: int | None = 2001,
            floor: bool = True, speed: float = 5.0) -> dict:
    """C2 / C4: n dynamic unit boxes uniform in [0,L)^3 (+ one static floor slab).
    C2: n=100_000, L=126, seed 1001.   C4: n=16_777_216, L=695, seed 1003."""
    m = n + (1 if floor else 0)
    s = _empty(m)
    i = np.arange(n)
    for ax in range(3):
        s["pos"][:n, ax] = uniform01(seed, i, ax) * np.float32(L)
        s["vel"][:n, ax] = (np.float32(2.0) * uniform01(seed, i, 3 + ax) - np.float32(1.0)) * np.float32(speed)
    s["mass"][:n] = np.float32(1.0) + np.float32(9.0) * uniform01(seed, i, 6)
    if floor:
        _set_static(s, n, (L / 2, -0.5, L / 2), (L / 2 + 1, 0.5, L / 2 + 1))
    if rank_seed is not None:
        s["rank"] = shuffled_rank(m, rank_seed)
    return s


def pile(nx: int = 100, ny: int = 100, nz: int = 100, seed: int = 1002, rank_seed: int | None = 2002,
         spacing: float = 1.05, jitter: float = 0.02) -> dict:
    """C3: jittered lattice of unit boxes (mass 1, at rest) over a static floor, inside 4 static walls."""
    n = nx * ny * nz
    s = _empty(n + 5)
    i = np.arange(n)
    ix, iy, iz = i % nx, (i // nx) % ny, i // (nx * ny)
    sp, jt = np.float32(spacing), np.float32(jitter)
    for ax, g in enumerate((ix, iy, iz)):
        j = jt * (np.float32(2.0) * uniform01(seed, i, ax) - np.float32(1.0))
        base = g.astype(np.float32) * sp + (np.float32(0.6) if ax == 1 else np.float32(0.5))
        s["pos"][:n, ax] = base + j
    wx, wz = nx * spacing, nz * spacing
    h = ny * spacing + 4.0
    _set_static(s, n + 0, (wx / 2, -0.5, wz / 2), (wx / 2 + 2, 0.5, wz / 2 + 2))          # floor (top at y=0)
    _set_static(s, n + 1, (-0.5 - 0.05, h / 2, wz / 2), (0.5, h / 2 + 1, wz / 2 + 2))      # wall x-
    _set_static(s, n + 2, (wx + 0.5 + 0.05, h / 2, wz / 2), (0.5, h / 2 + 1, wz / 2 + 2))  # wall x+
    _set_static(s, n + 3, (wx / 2, h / 2, -0.5 - 0.05), (wx / 2 + 2, h / 2 + 1, 0.5))      # wall z-
    _set_static(s, n + 4, (wx / 2, h / 2, wz + 0.5 + 0.05), (wx / 2 + 2, h / 2 + 1, 0.5))  # wall z+
    if rank_seed is not None:
        s["rank"] = shuffled_rank(n + 5, rank_seed)
    return s


def mixed(n: int = 4_194_304, L: float = 2500.0, seed: int = 1004, rank_seed: int | None = 2004) -> dict:
    """C5: cubic boxes with half-extent log-uniform in [0.05, 5] (100:1), 10 % static, + static floor."""
    s = _empty(n + 1)
    i = np.arange(n)
    u = uniform01(seed, i, 6)
    h = (np.float32(0.05) * np.power(np.float32(100.0), u, dtype=np.float32)).astype(np.float32)
    st = uniform01(seed, i, 7) < np.float32(0.1)
    for ax in range(3):
        s["pos"][:n, ax] = uniform01(seed, i, ax) * np.float32(L)
        s["vel"][:n, ax] = (np.float32(2.0) * uniform01(seed, i, 3 + ax) - np.float32(1.0)) * np.float32(2.0)
        s["box_min"][:n, ax] = -h
        s["box_max"][:n, ax] = h
    s["mass"][:n] = np.clip(h * h * h * np.float32(8.0), np.float32(0.1), np.float32(1e4))
    s["is_static"][:n] = st
    s["vel"][:n][st] = 0
    s["mass"][:n][st] = 0
    _set_static(s, n, (L / 2, -0.5, L / 2), (L / 2 + 6, 0.5, L / 2 + 6))
    if rank_seed is not None:
        s["rank"] = shuffled_rank(n + 1, rank_seed)
    return s


def subset(s: dict, idx: np.ndarray) -> dict:
    """Bodies `idx` of scene s, iteration order preserved (ranks compacted)."""
    idx = np.asarray(idx)
    out = {}
    for k, v in s.items():
        out[k] = None if v is None else np.ascontiguousarray(v[idx])
    if s.get("rank") is not None:
        r = s["rank"][idx]
        order = np.argsort(r, kind="stable")
        nr = np.empty(len(idx), np.uint32)
        nr[order] = np.arange(len(idx), dtype=np.uint32)
        out["rank"] = nr
    return out

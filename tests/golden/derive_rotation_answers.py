"""Derives, WITHOUT the oracle and without any float hardware, float32 known answers for the two places of the hot
path where cgmath's quaternion arithmetic matters, and writes them to tests/golden/rotation_answers.json together with
a step-by-step trace (tests/golden/ROTATION_DERIVATION.md):

  R1  AABBCollisionBox::get_rotated_aabb (reference crates/gears-ecs/src/components/physics.rs:84-135) for an
      ASYMMETRIC box under a quarter turn about y and Scale::Uniform(3.0) -- the shapes the API is used with in
      examples/src/interactive/map.rs:325-346 (walls / trees: Rotation3::from_axis_angle(unit_y / unit_x, Deg(+-90)))
      and :599-613 (concrete blocks: Scale::Uniform(3.0));
  R2  Instance::to_raw (crates/gears-renderer/src/instance.rs:22-31) for the same quaternion, position and scale.

Arithmetic: every operation is carried out EXACTLY on rationals (fractions.Fraction) and then rounded to the nearest
binary32 value, ties to even (fl() below) -- i.e. IEEE-754 single precision without fused multiply-add, which is what
rustc emits for cgmath's scalar code.  Operation order (cgmath 0.18.0, Cargo.lock:393-399, source not vendored; as
published in its quaternion.rs / matrix.rs):
  Quaternion * Vector3:   tmp = q.v x v + v * q.s ;  result = (q.v x tmp) * 2 + v
  a x b               :   (a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x)
  Matrix3::from(q)    :   x2 = x+x, y2 = y+y, z2 = z+z; xx2 = x2*x, xy2 = x2*y, xz2 = x2*z, yy2 = y2*y, yz2 = y2*z,
                          zz2 = z2*z, sy2 = y2*s, sz2 = z2*s, sx2 = x2*s;
                          columns (1-yy2-zz2, xy2+sz2, xz2-sy2), (xy2-sz2, 1-xx2-zz2, yz2+sx2), (xz2+sy2, yz2-sx2, 1-xx2-yy2)
  Matrix4 * Matrix4   :   column j = ((a*r0 + b*r1) + c*r2) + d*r3   (a..d = columns of the left operand)
The quaternion of a quarter turn about y is (v = (0, h, 0), s = h) with h = fl(sin(pi/4)) = fl(cos(pi/4)) = 0x3f3504f3
(Rotation3::from_axis_angle: (s, c) = sin_cos(angle/2), Quaternion::from_sv(c, axis * s)); the fixture pins the
quaternion BITS, so no trigonometry is involved in the check.

    python tests/golden/derive_rotation_answers.py
"""
import json
import os
import struct
from fractions import Fraction as F

HERE = os.path.dirname(os.path.abspath(__file__))
TRACE = []


def fl(x: F) -> F:
    """Nearest binary32 to the rational x, ties to even (normal range only, which is all that occurs here)."""
    if x == 0:
        return F(0)
    s = -1 if x < 0 else 1
    a = abs(x)
    e = 0
    while a >= 2:
        a /= 2
        e += 1
    while a < 1:
        a *= 2
        e -= 1
    assert -126 <= e <= 127
    m = a * (1 << 23)            # in [2^23, 2^24)
    lo = m.numerator // m.denominator
    rem = m - lo
    if rem > F(1, 2) or (rem == F(1, 2) and (lo & 1)):
        lo += 1
    return s * F(lo, 1 << 23) * (F(2) ** e)


def bits(x: F) -> str:
    return "0x%08x" % struct.unpack("<I", struct.pack("<f", float(x)))[0]   # x is exactly representable


def from_bits(b: int) -> F:
    return F(struct.unpack("<f", struct.pack("<I", b))[0])


def mul(a, b): return fl(a * b)
def add(a, b): return fl(a + b)
def sub(a, b): return fl(a - b)


def cross(a, b):
    return (sub(mul(a[1], b[2]), mul(a[2], b[1])), sub(mul(a[2], b[0]), mul(a[0], b[2])), sub(mul(a[0], b[1]), mul(a[1], b[0])))


def quat_mul_vec(qv, qs, v):
    c = cross(qv, v)
    tmp = tuple(add(c[k], mul(v[k], qs)) for k in range(3))
    c2 = cross(qv, tmp)
    return tuple(add(mul(c2[k], F(2)), v[k]) for k in range(3))


def say(line):
    TRACE.append(line)


def vec_s(v):
    return "(" + ", ".join(f"{float(x):.9g} {bits(x)}" for x in v) + ")"


def derive(name, qbits, title):
    qv, qs = tuple(from_bits(b) for b in qbits[:3]), from_bits(qbits[3])
    h = qs
    bmin, bmax = (F(-3, 2), F(-1), F(-3, 2)), (F(3, 2), F(1), F(10))
    scale = (F(3), F(3), F(3))
    pos = (F(14), F(1), F(38))
    say(f"# {name}: {title}\n")
    say(f"q.v = {vec_s(qv)}, q.s = {vec_s((qs,))}; box min (-1.5, -1, -1.5) max (1.5, 1, 10); "
        f"Scale::Uniform(3.0); pos (14, 1, 38)\n")
    say("## R1 get_rotated_aabb (physics.rs:84-135): corner * scale, q * v, + pos, component-wise min / max\n")
    lo = hi = None
    for i in range(8):
        c = tuple(bmax[k] if (i >> k) & 1 else bmin[k] for k in range(3))
        s = tuple(mul(c[k], scale[k]) for k in range(3))          # physics.rs:107-113
        r = quat_mul_vec(qv, qs, s)                               # physics.rs:115
        w = tuple(add(r[k], pos[k]) for k in range(3))            # physics.rs:117
        say(f"corner {i}: c = {tuple(float(x) for x in c)}, c*scale = {tuple(float(x) for x in s)}, "
            f"q*v = {vec_s(r)}, world = {vec_s(w)}")
        lo = w if lo is None else tuple(min(lo[k], w[k]) for k in range(3))
        hi = w if hi is None else tuple(max(hi[k], w[k]) for k in range(3))
    say(f"\nworld AABB min = {vec_s(lo)}\n\nworld AABB max = {vec_s(hi)}\n")
    # ---- R2 Instance::to_raw ----
    x, y, z, s_ = qv[0], qv[1], qv[2], qs
    x2, y2, z2 = add(x, x), add(y, y), add(z, z)
    xx2, xy2, xz2 = mul(x2, x), mul(x2, y), mul(x2, z)
    yy2, yz2, zz2 = mul(y2, y), mul(y2, z), mul(z2, z)
    sy2, sz2, sx2 = mul(y2, s_), mul(z2, s_), mul(x2, s_)
    one = F(1)
    r3 = [sub(sub(one, yy2), zz2), add(xy2, sz2), sub(xz2, sy2),
          sub(xy2, sz2), sub(sub(one, xx2), zz2), add(yz2, sx2),
          add(xz2, sy2), sub(yz2, sx2), sub(sub(one, xx2), yy2)]
    say("## R2 Instance::to_raw (instance.rs:22-31)\n")
    say(f"y2 = {vec_s((y2,))}, yy2 = {vec_s((yy2,))}, sy2 = {vec_s((sy2,))}; Matrix3::from(q), column-major = "
        + vec_s(r3) + "\n")
    Z, O = F(0), F(1)
    T = [O, Z, Z, Z, Z, O, Z, Z, Z, Z, O, Z, pos[0], pos[1], pos[2], O]
    R = [r3[0], r3[1], r3[2], Z, r3[3], r3[4], r3[5], Z, r3[6], r3[7], r3[8], Z, Z, Z, Z, O]
    S = [scale[0], Z, Z, Z, Z, scale[1], Z, Z, Z, Z, scale[2], Z, Z, Z, Z, O]

    def mat4(A, B):
        out = [Z] * 16
        for j in range(4):
            for i in range(4):
                acc = mul(A[0 * 4 + i], B[j * 4 + 0])
                acc = add(acc, mul(A[1 * 4 + i], B[j * 4 + 1]))
                acc = add(acc, mul(A[2 * 4 + i], B[j * 4 + 2]))
                acc = add(acc, mul(A[3 * 4 + i], B[j * 4 + 3]))
                out[j * 4 + i] = acc
        return out
    model = mat4(mat4(T, R), S)
    say("model (column-major) = " + vec_s(model) + "\n")
    return {
        "title": title,
        "quat_bits_v_xyz_s": [bits(v) for v in qv] + [bits(qs)],
        "box_min": [-1.5, -1.0, -1.5], "box_max": [1.5, 1.0, 10.0], "scale": [3.0, 3.0, 3.0], "pos": [14.0, 1.0, 38.0],
        "world_aabb_min_bits": [bits(v) for v in lo], "world_aabb_max_bits": [bits(v) for v in hi],
        "model_bits": [bits(v) for v in model], "normal_bits": [bits(v) for v in r3],
    }


def main():
    out = {
        "derivation": "tests/golden/derive_rotation_answers.py (exact rationals + round-to-nearest-even binary32 per "
                      "operation; no oracle, no float hardware); trace in tests/golden/ROTATION_DERIVATION.md",
        "cases": {
            "quarter_turn_y": derive("quarter_turn_y", [0x00000000, 0x3F3504F3, 0x00000000, 0x3F3504F3],
                                     "Rotation3::from_axis_angle(unit_y, Deg(90)) as in examples/src/interactive/map.rs:325"),
            # fl(1/sqrt(30)) * (1, 2, 3, 4): a generic unit quaternion, every product rounds
            "generic": derive("generic", [0x3E3AF4BA, 0x3EBAF4BA, 0x3F0C378B, 0x3F3AF4BA],
                              "a generic rotation, (1, 2, 3; 4) / sqrt(30) rounded to binary32"),
        },
    }
    with open(os.path.join(HERE, "rotation_answers.json"), "w") as f:
        json.dump(out, f, indent=1)
    with open(os.path.join(HERE, "ROTATION_DERIVATION.md"), "w") as f:
        f.write("\n".join(TRACE) + "\n")
    print("wrote rotation_answers.json:", {k: v["world_aabb_max_bits"] for k, v in out["cases"].items()})


if __name__ == "__main__":
    main()

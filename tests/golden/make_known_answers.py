"""Writes tests/golden/known_answers.json.

The reference (Rust) cannot run in this container, so these are NOT outputs of the reference
binary.  They are the float32 known-answer vectors of SURVEY.md Appendix B: inputs are the
reference's own unit-test inputs (crates/gears-ecs/src/components/physics.rs:696-892), results were
derived by hand-following Appendix A's operation order in IEEE binary32 with glibc 2.39 expf, and the
bit patterns below were typed in from that appendix — the oracle is then CHECKED against this
file (tests/test_oracle_known_answers.py), it does not generate it.
"""
import json
import os

KA = {
    "dt_1_60_bits": "0x3c888889",
    "KA1": {"ref": "physics.rs:696-708", "pos": [0, 10, 0], "vel": [0, 0, 0], "acc": [0, 0, 0], "dt": 0.016,
            "vel_y_bits": "0xbed8e24d", "pos_y_bits": "0x411fe43d"},
    "KA2": {"ref": "physics.rs:955-974", "pos": [0, 10, 0], "vel": [0, -5, 0], "acc": [0, 0, 0], "dt": 0.016,
            "vel_y_bits": "0xc0b91687", "pos_y_bits": "0x411e84f1"},
    "KA3": {"ref": "physics.rs:725-742", "pos": [0, 0, 0], "vel": [0, 0, 0], "acc": [10, 0, 0], "dt": 0.1,
            "vel_bits": ["0x3f800000", "0xc018b444", "0x00000000"],
            "pos_bits": ["0x3dcccccd", "0xbe7453a0", "0x00000000"]},
    "KD": {"damp_bits": ["0x3f794378", "0x3f786f1e", "0x3f717e6f"]},
    "KR1": {"ref": "physics.rs:803-829",
            "a": {"pos": [1.5, 0, 0], "vel": [-5, 0, 0], "mass": 1, "restitution": 0.2, "is_static": False, "half": [1, 1, 1]},
            "b": {"pos": [0, 0, 0], "vel": [0, 0, 0], "mass": 0, "restitution": 0.2, "is_static": True, "half": [1, 1, 1]},
            "a_pos_bits": ["0x3fd91687", None, None], "a_vel_bits": ["0x3f800000", None, None]},
    "KR2": {"ref": "physics.rs:832-857",
            "a": {"pos": [-0.2, 0, 0], "vel": [1, 0, 0], "mass": 1, "restitution": 0.2, "is_static": False, "half": [0.5, 0.5, 0.5]},
            "b": {"pos": [0.2, 0, 0], "vel": [-1, 0, 0], "mass": 1, "restitution": 0.2, "is_static": False, "half": [0.5, 0.5, 0.5]},
            "a_pos_bits": ["0xbea2d0e6", None, None], "a_vel_bits": ["0xbe4cccd0", None, None],
            "b_pos_bits": ["0x3ea2d0e6", None, None], "b_vel_bits": ["0x3e4cccd0", None, None]},
    "KR3": {"ref": "physics.rs:860-892",
            "a": {"pos": [0, 0.8, 0], "vel": [0, -10, 0], "mass": 1, "restitution": 0.8, "is_static": False, "half": [0.5, 0.5, 0.5]},
            "b": {"pos": [0, 0, 0], "vel": [0, 0, 0], "mass": 0, "restitution": 0.8, "is_static": True, "half": [10, 0.5, 10]},
            "a_pos_bits": [None, "0x3f604189", None], "a_vel_bits": [None, "0x41000000", None]},
    "KR4": {"ref": "physics.rs:767-800",
            "a": {"pos": [0.8, 0, 0], "vel": [-1, 0, 0], "mass": 1, "restitution": 0.2, "is_static": False, "half": [1, 1, 1]},
            "b": {"pos": [0, 0, 0], "vel": [0, 0, 0], "mass": 0, "restitution": 0.2, "is_static": True, "half": [1, 1, 1]},
            "a_pos_bits": ["0x3fa353f8", None, None], "a_vel_bits": ["0x3e4cccd0", None, None]},
    "KR5": {"ref": "friction quirk, physics.rs:273-290",
            "a": {"pos": [0, 0.9, 0], "vel": [3, -2, 0], "mass": 2, "restitution": 0.2, "is_static": False, "half": [0.5, 0.5, 0.5]},
            "b": {"pos": [0, 0, 0], "vel": [0, 0, 0], "mass": 0, "restitution": 0.2, "is_static": True, "half": [10, 0.5, 10]},
            "a_pos_bits": [None, "0x3f6f9db2", None], "a_vel_bits": ["0x40700000", "0x3eccccd0", "0x00000000"]},
}

if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "known_answers.json")
    with open(out, "w") as f:
        json.dump(KA, f, indent=1)
    print("wrote", out)

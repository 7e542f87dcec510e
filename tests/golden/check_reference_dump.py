#!/usr/bin/env python
"""Compares the output of rust/golden-dump (run against the REAL reference crates on a machine with cargo) with the
fixtures this repository's oracle and CUDA path are pinned to.  A clean run turns the parity pin from "hand-derived"
into "produced by the reference".

    python tests/golden/check_reference_dump.py reference_dump.json
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def main(path):
    ref = json.load(open(path))
    ka = json.load(open(os.path.join(HERE, "known_answers.json")))
    rot = json.load(open(os.path.join(HERE, "rotation_answers.json")))["cases"]
    bad = []

    def cmp(what, got, want):
        for i, (g, w) in enumerate(zip(got, want)):
            if w is not None and g.lower() != w.lower():
                bad.append(f"{what}[{i}]: reference {g} vs fixture {w}")
    cmp("KA1.vel_y", [ref["KA1"]["vel_bits"][1]], [ka["KA1"]["vel_y_bits"]])
    cmp("KA1.pos_y", [ref["KA1"]["pos_bits"][1]], [ka["KA1"]["pos_y_bits"]])
    cmp("KA2.vel_y", [ref["KA2"]["vel_bits"][1]], [ka["KA2"]["vel_y_bits"]])
    cmp("KA2.pos_y", [ref["KA2"]["pos_bits"][1]], [ka["KA2"]["pos_y_bits"]])
    cmp("KA3.vel", ref["KA3"]["vel_bits"], ka["KA3"]["vel_bits"])
    cmp("KA3.pos", ref["KA3"]["pos_bits"], ka["KA3"]["pos_bits"])
    cmp("KD", ref["KD"]["damp_bits"], ka["KD"]["damp_bits"])
    for k in ("KR1", "KR2", "KR3", "KR4", "KR5"):
        for f in ("a_pos_bits", "a_vel_bits", "b_pos_bits", "b_vel_bits"):
            if f in ka[k]:
                cmp(f"{k}.{f}", ref[k][f], ka[k][f])
    for name, g in rot.items():
        r = ref["rotation"][name]
        for f in ("world_aabb_min_bits", "world_aabb_max_bits", "model_bits", "normal_bits"):
            cmp(f"rotation.{name}.{f}", r[f], g[f])
    if bad:
        print("\n".join(bad))
        print(f"{len(bad)} value(s) differ: the oracle's restatement is NOT what the reference computes")
        return 1
    print("every known answer equals the reference's own output: parity is pinned by the reference")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1]))

#!/usr/bin/env python3
"""Writes tests/golden/kat_vectors.json: the hand-derived known-answer vectors KAT1-5 of SURVEY.md §4.

The reference repository has no tests or fixtures, so these vectors were derived by following
/root/reference/cuking.cu:216-307 by hand in IEEE fp32; the expected values below are literals, NOT outputs of the
oracle or of the product, so both can be checked against them.
"""
import json
import os

KATS = [
    # id, genotypes i, genotypes j ('.' = missing), het_i, het_j, het-het, opp-hom, conc-hom, shared, ibs0/1/2, kin hex, emitted
    ("KAT1", "0 1 2 1 0 . 1 2 0 1", "0 1 0 1 2 1 . 2 1 1", 3, 4, 3, 2, 2, 8, (2, 1, 5), "be800000", True),
    ("KAT2", "1 1 1 0 2 1 0 0 1 2 1 1", "1 0 1 1 2 1 2 0 . 2 1 0", 6, 5, 4, 1, 3, 11, (1, 3, 7), "3e19999a", True),
    ("KAT3", "0 0 2 2 . 0", "0 2 2 0 1 0", 0, 0, 0, 2, 3, 5, (2, 0, 3), "ff800000", False),  # -inf: never emitted
    ("KAT4", "0 0 2", "0 0 2", 0, 0, 0, 0, 3, 3, (0, 0, 3), "nan", False),                   # 0/0: never emitted
    ("KAT5", "1 1 0 2 1 0 1", "1 1 0 2 1 0 1", 4, 4, 4, 0, 3, 7, (0, 0, 7), "3f000000", True),
]

out = []
for kid, gi, gj, het_i, het_j, bh, opp, conc, shared, ibs, kin_hex, emitted in KATS:
    out.append({
        "id": kid,
        "genotypes_i": [-1 if t == "." else int(t) for t in gi.split()],
        "genotypes_j": [-1 if t == "." else int(t) for t in gj.split()],
        "het_i": het_i, "het_j": het_j, "both_het": bh, "opposing_hom": opp, "concordant_hom": conc,
        "shared_sites": shared, "ibs0": ibs[0], "ibs1": ibs[1], "ibs2": ibs[2],
        "kin_f32_hex": kin_hex, "emitted_at_threshold_minus_1": emitted,
    })
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat_vectors.json")
with open(path, "w") as f:
    json.dump(out, f, indent=1)
print("wrote", path)

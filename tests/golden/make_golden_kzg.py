#!/usr/bin/env python3
"""Generate tests/golden/kzg_golden.json.

Sources of the vectors:
  * "reference_kats": the Lagrange-basis known answers the reference's own tests pin (multilinear_kzg/src/trusted_setup.rs:101-126),
    transcribed by hand, and the inputs of its three KZG tests (multilinear_kzg.rs:223-303), which only assert verify() == true;
  * "generated": for those three inputs (and one seeded 2^5 polynomial) the commitment, the evaluation and every opening proof as
    computed by the independent Python big-int model oracle/pykzg.py (affine chord-and-tangent arithmetic, plain integers) --
    each checked there with the pairing form of the reference's verify() before it is written.  NOT pinned by the reference:
    they pin our reading of `mul_bigint` sums, quotients and the blow-up, for the C oracle, the host verifier and the GPU path alike.
Points are affine (x, y) hex strings, null = the point at infinity; scalars are canonical integers (decimal strings).

Run from the repo root:  python tests/golden/make_golden_kzg.py
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pykzg as pk  # noqa: E402

pk.self_check()
R = pk.R


def pt(p):
    return None if p is None else ["%096x" % p[0], "%096x" % p[1]]


reference_kats = {
    "lagrange_basis": [
        {"src": "trusted_setup.rs:101-118", "taus": [5, 2, 3], "out": [-8, 12, 16, -24, 10, -15, -20, 30]},
        {"src": "trusted_setup.rs:120-126", "taus": [5, 2], "out": [4, -8, -5, 10]},
    ],
    "kzg_tests": [
        {"src": "multilinear_kzg.rs:223-246", "taus": [5, 2, 3], "values": [0, 4, 0, 4, 0, 4, 3, 7], "opening": [6, 4, 0]},
        {"src": "multilinear_kzg.rs:248-272", "taus": [2, 3, 4], "values": [0, 7, 0, 5, 0, 7, 4, 9], "opening": [5, 9, 6]},
        {"src": "multilinear_kzg.rs:274-303", "taus": [12, 9, 28, 40],
         "values": [0, 0, 0, 2, 0, 0, 10, 12, 0, -12, 4, -6, 0, -12, 14, 4], "opening": [54, 90, 76, 160]},
    ],
}

generated = []
rnd = random.Random(0xB200)
cases = [(c["taus"], c["values"], c["opening"], c["src"]) for c in reference_kats["kzg_tests"]]
cases.append(([rnd.randrange(R) for _ in range(5)], [rnd.randrange(R) for _ in range(32)], [rnd.randrange(R) for _ in range(5)], "seeded 2^5"))
for taus, values, opening, src in cases:
    taus, values, opening = ([x % R for x in v] for v in (taus, values, opening))
    setup = pk.TrustedSetup.initialize(taus)
    c = pk.commit(values, setup)
    v, proofs = pk.open_and_prove(values, setup, opening)
    assert pk.verify(setup, c, opening, v, proofs) and pk.verify_with_trapdoor(taus, c, opening, v, proofs)
    generated.append({"src": src, "taus": [str(x) for x in taus], "values": [str(x) for x in values], "opening": [str(x) for x in opening],
                      "g1_powers_of_tau_first": pt(setup.g1_powers_of_tau[0]), "g1_powers_of_tau_last": pt(setup.g1_powers_of_tau[-1]),
                      "commitment": pt(c), "evaluation": str(v), "proofs": [pt(p) for p in proofs]})

out = {"curve": {"g1_generator": pt(pk.G1_GEN), "g2_generator": [["%096x" % c for c in pk.G2_GEN[0]], ["%096x" % c for c in pk.G2_GEN[1]]],
                 "q": "%096x" % pk.Q, "r": "%064x" % pk.R},
       "reference_kats": reference_kats, "generated": generated}
with open(os.path.join(ROOT, "tests", "golden", "kzg_golden.json"), "w") as f:
    json.dump(out, f, indent=1)
    f.write("\n")
print("wrote kzg_golden.json:", len(generated), "generated cases")

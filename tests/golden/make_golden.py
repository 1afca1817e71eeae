#!/usr/bin/env python3
"""Generate tests/golden/golden.json.

Sources of the vectors:
  * "reference_kats": the small-integer known-answer tests the reference's own unit tests pin
    (file:line given per entry) -- transcribed by hand from /root/reference, NOT computed here;
  * "appendix_b": the survey-derived transcript / proof vectors of SURVEY.md appendix B (parity is
    unpinned by the reference for everything that flows through Fiat-Shamir; these pin our reading);
  * "generated": seeded proofs produced by the independent Python big-int model oracle/pyoracle.py.

Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle as po  # noqa: E402

po.self_check()

reference_kats = {
    # polynomials/src/multilinear/evaluation_form.rs:177-220 (BN254 Fq)
    "new_panics": {"src": "evaluation_form.rs:171-176", "field": "BN254_FQ", "table": [0, 0, 3, 8, 0, 0],
                   "message": "Evaluated values must be a power of 2"},
    "partial_evaluate": [
        {"src": "evaluation_form.rs:179-185", "field": "BN254_FQ", "table": [0, 0, 3, 8], "var": 0, "r": 6, "out": [18, 48]},
        {"src": "evaluation_form.rs:186-189", "field": "BN254_FQ", "table": [0, 0, 3, 8], "var": 1, "r": 2, "out": [0, 13]},
        {"src": "evaluation_form.rs:191-195", "field": "BN254_FQ", "table": [18, 48], "var": 0, "r": 2, "out": [78]},
        {"src": "evaluation_form.rs:197-211", "field": "BN254_FQ", "table": [0, 0, 0, 3, 0, 0, 2, 5], "var": 2, "r": 3, "out": [0, 9, 0, 11]},
    ],
    "evaluate": [
        {"src": "evaluation_form.rs:214-220", "field": "BN254_FQ", "table": [0, 0, 3, 8], "values": [6, 2], "out": 78},
    ],
    "tensor": [
        {"src": "evaluation_form.rs:223-241", "field": "BN254_FQ", "op": "add", "wb": [1, 2], "wc": [3, 4], "out": [4, 5, 5, 6]},
        {"src": "evaluation_form.rs:244-267", "field": "BN254_FQ", "op": "mul", "wb": [2, 3], "wc": [4, 5], "out": [8, 10, 12, 15]},
    ],
    "tensor_panics": {"src": "evaluation_form.rs:270-277", "field": "BN254_FQ", "wb": [2, 3], "wc": [4],
                      "message": "Different polynomial length"},
    # polynomials/src/composed/product_polynomial.rs:94-173
    "product_polynomial": {"src": "product_polynomial.rs:107-173", "field": "BN254_FQ",
                            "polys": [[0, 0, 0, 2], [0, 0, 0, 3]],
                            "evaluate_at": [1, 2], "evaluate_out": 24, "fold_var": 0, "fold_r": 2,
                            "fold_out": [[0, 4], [0, 6]], "elementwise": [0, 0, 0, 6], "degree": 2},
    "product_panics": {"src": "product_polynomial.rs:94-104", "field": "BN254_FQ", "polys": [[0, 2], [0, 0, 0, 3]],
                       "message": "different number of variables"},
    # polynomials/src/composed/sum_polynomial.rs:102-245
    "sum_polynomial": {"src": "sum_polynomial.rs:117-245", "field": "BN254_FQ",
                       "products": [[[0, 0, 0, 2], [0, 0, 0, 3]], [[0, 0, 0, 1], [0, 0, 0, 2]]],
                       "evaluate_at": [1, 2], "evaluate_out": 32, "fold_var": 0, "fold_r": 2,
                       "fold_out": [[[0, 4], [0, 6]], [[0, 2], [0, 4]]],
                       "elementwise": [0, 0, 0, 8], "degree": 2, "number_of_variables": 2},
    # polynomials/src/univariate/dense_univariate.rs:186-261
    "univariate_evaluate": {"src": "dense_univariate.rs:186-222", "field": "BN254_FQ",
                            "coeffs": [0, 0, 2, 0, 0, 0, 0, 3], "x": 2, "out": 392, "degree": 7},
    "lagrange": {"src": "dense_univariate.rs:251-260", "field": "BN254_FQ", "xs": [0, 1, 2], "ys": [2, 4, 10], "coeffs": [2, 0, 2]},
    # sumcheck_protocol/src/gkr_sumcheck/sumcheck_gkr_protocol.rs:163-212
    "round_univariate": {"src": "sumcheck_gkr_protocol.rs:163-186", "field": "BN254_FQ",
                         "products": [[[0, 0, 0, 2], [0, 0, 0, 3]], [[0, 0, 0, 2], [0, 0, 0, 3]]], "evals": [0, 12, 48]},
    "product_round_trip": {"src": "sumcheck_gkr_protocol.rs:188-212", "field": "BN254_FQ",
                           "products": [[[0, 0, 0, 2], [0, 0, 0, 3]], [[0, 0, 0, 2], [0, 0, 0, 3]]], "claimed_sum": 12},
    # sumcheck_protocol/src/basic_sumcheck/protocol.rs:9-116, prover.rs:95-107
    "basic_claimed_sum": [
        {"src": "protocol.rs:9-26", "field": "BLS12_381_FR", "table": [0, 0, 2, 7, 3, 3, 6, 11], "sum": 32},
        {"src": "prover.rs:100-107", "field": "BN254_FQ", "table": [0, 0, 3, 8], "sum": 11},
    ],
    "basic_round_trips": [
        {"src": "protocol.rs:28-56", "field": "BLS12_381_FR", "constant": 3, "log2": 20},
        {"src": "protocol.rs:58-86", "field": "BLS12_381_FR", "table": [0, 0, 0, 0, 0, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0]},
        {"src": "protocol.rs:88-116", "field": "BLS12_381_FR", "table": [1, 3, 5, 7, 2, 4, 6, 8, 3, 5, 7, 9, 4, 6, 8, 10]},
    ],
    # circuit/src/arithmetic_circuit.rs:216-384; gates are (left, right, out, op) with op 0 = Add, 1 = Mul
    "circuit_evaluate": [
        {"src": "arithmetic_circuit.rs:216-238", "field": "BN254_FQ", "inputs": [2, 3, 4, 5],
         "layers": [[[0, 1, 0, 1]], [[0, 1, 0, 0], [2, 3, 1, 1]]], "layer_evaluations": [[100], [5, 20], [2, 3, 4, 5]]},
        {"src": "arithmetic_circuit.rs:240-262", "field": "BN254_FQ", "inputs": [1, 2, 3, 4],
         "layers": [[[0, 1, 0, 0]], [[0, 1, 1, 0], [2, 3, 0, 1]]], "layer_evaluations": [[15], [12, 3], [1, 2, 3, 4]]},
        {"src": "arithmetic_circuit.rs:264-303", "field": "BN254_FQ", "inputs": [1, 2, 3, 4, 5, 6, 7, 8],
         "layers": [[[0, 1, 0, 0]], [[0, 1, 0, 0], [2, 3, 1, 1]], [[0, 1, 0, 0], [2, 3, 1, 1], [4, 5, 2, 1], [6, 7, 3, 1]]],
         "output": [1695]},
    ],
    "num_of_layer_variables": {"src": "arithmetic_circuit.rs:305-317", "values": [[0, 3], [1, 5], [2, 8], [3, 11], [4, 14]]},
    "add_i_mul_i": [
        {"src": "arithmetic_circuit.rs:319-356", "layers": [[[0, 1, 0, 0]], [[0, 1, 1, 0], [2, 3, 0, 1]]], "layer": 0,
         "add_ones": [1], "mul_ones": [], "size": 8},
        {"src": "arithmetic_circuit.rs:358-384", "layers": [[[0, 1, 0, 0]], [[0, 1, 1, 0], [2, 3, 0, 1]]], "layer": 1,
         "add_ones": [17], "mul_ones": [11], "size": 32},
    ],
    # gkr/src/gkr_protocol.rs:246-299 (round trips only)
    "gkr_round_trips": [
        {"src": "gkr_protocol.rs:246-263", "field": "BN254_FQ", "inputs": [2, 3, 4, 5],
         "layers": [[[0, 1, 0, 1]], [[0, 1, 0, 0], [2, 3, 1, 1]]]},
        {"src": "gkr_protocol.rs:265-299", "field": "BN254_FQ", "inputs": [1, 2, 3, 4, 5, 6, 7, 8],
         "layers": [[[0, 1, 0, 0]], [[0, 1, 0, 1], [2, 3, 1, 0]], [[0, 1, 0, 0], [2, 3, 1, 0], [4, 5, 2, 0], [6, 7, 3, 0]]]},
    ],
}


def _check_reference_kats():
    """the hand-transcribed KATs must hold in the Python model (guards against transcription slips)"""
    for k in reference_kats["partial_evaluate"]:
        assert po.partial_evaluate(k["table"], k["var"], k["r"], po.P[k["field"]]) == k["out"], k
    for k in reference_kats["evaluate"]:
        assert po.mle_evaluate(k["table"], k["values"], po.P[k["field"]]) == k["out"]
    for k in reference_kats["tensor"]:
        f = po.tensor_add if k["op"] == "add" else po.tensor_mul
        assert f(k["wb"], k["wc"], po.P[k["field"]]) == k["out"]
    pp = reference_kats["product_polynomial"]
    p = po.P[pp["field"]]
    assert po.product_elementwise(pp["polys"], p) == pp["elementwise"]
    assert [po.partial_evaluate(t, 0, pp["fold_r"], p) for t in pp["polys"]] == pp["fold_out"]
    acc = 1
    for t in pp["polys"]:
        acc = acc * po.mle_evaluate(t, pp["evaluate_at"], p) % p
    assert acc == pp["evaluate_out"]
    sp = reference_kats["sum_polynomial"]
    assert po.sumpoly_elementwise(sp["products"], p) == sp["elementwise"]
    assert po.sumpoly_evaluate(sp["products"], sp["evaluate_at"], p) == sp["evaluate_out"]
    assert po.sumpoly_partial_evaluate(sp["products"], 0, sp["fold_r"], p) == sp["fold_out"]
    ue = reference_kats["univariate_evaluate"]
    assert po.uni_evaluate(ue["coeffs"], ue["x"], p) == ue["out"]
    lg = reference_kats["lagrange"]
    assert po.lagrange_interpolate(lg["xs"], lg["ys"], p) == lg["coeffs"]
    ru = reference_kats["round_univariate"]
    assert po.generate_round_univariate(ru["products"], p) == ru["evals"]
    rt = reference_kats["product_round_trip"]
    polys, _, _ = po.product_prove(rt["products"], rt["claimed_sum"], po.Transcript(), p)
    assert po.product_verify(rt["claimed_sum"], polys, po.Transcript(), p)[0]
    for k in reference_kats["basic_claimed_sum"]:
        assert sum(k["table"]) % po.P[k["field"]] == k["sum"]
    for k in reference_kats["circuit_evaluate"]:
        ev = po.circuit_evaluate([[po.Gate(*g) for g in l] for l in k["layers"]], k["inputs"], po.P[k["field"]])
        if "layer_evaluations" in k:
            assert ev == k["layer_evaluations"]
        else:
            assert ev[0] == k["output"]
    for i, v in reference_kats["num_of_layer_variables"]["values"]:
        assert po.num_of_layer_variables(i) == v
    for k in reference_kats["add_i_mul_i"]:
        add, mul = po.add_i_mul_i([[po.Gate(*g) for g in l] for l in k["layers"]], k["layer"])
        assert len(add) == k["size"] and [i for i, v in enumerate(add) if v] == k["add_ones"]
        assert [i for i, v in enumerate(mul) if v] == k["mul_ones"]
    for k in reference_kats["gkr_round_trips"]:
        layers = [[po.Gate(*g) for g in l] for l in k["layers"]]
        pf = po.gkr_prove(layers, k["inputs"], po.P[k["field"]])
        assert po.gkr_verify(layers, pf, k["inputs"], po.P[k["field"]])


appendix_b = {
    "transcript": {"append": "boy", "sample": "acbbcf124e93ff6725437257164b2a8fc97dffbf62361af983ab933488b89bb4",
                   "field": "BN254_FQ",
                   "challenge": 0x0c807d31d48ab74fae7fc85ede4c3053530443438876552c324715986d1097a3},
    "keccak256_empty": "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470",
    "basic": {"field": "BLS12_381_FR", "table": [0, 0, 2, 7, 3, 3, 6, 11], "claimed_sum": 32,
              "round_polys": [[9, 23],
                              [30511417544605799419573844462630429561927407198359212793726032291763660055754,
                               5724639942723605573133298944716595524109507930793858509898937255725825951339],
                              [35073895304701777423868366994412653894656436401285037996038924113754839832535,
                               49148553107713759394493141382131185562200392126865752775946655507094044636920]],
              "challenges": [0x58873e7f7abcfb11efb0ee871c80fc431c9557a4d2c803ed511aa20cef8ed0cd,
                             0x063931de95609de7c20105593a4d7f003792d936b2c9cc42675d4a60781cb96d,
                             0x6001856e68be1e4603603f01617b77285f8e1d4bbaecd13a419bc14c2f6b44cb],
              "final": 0x13a02aa38a8fd8c226e4bda30fadcb4b3055955bf8d4b769995c5b4645841253},
    "gkr": {"field": "BN254_FQ", "src": "gkr_protocol.rs:248-256",
            "layers": [[[0, 1, 0, 1]], [[0, 1, 0, 0], [2, 3, 1, 1]]],   # (left, right, out, op) op: 0 add, 1 mul
            "inputs": [2, 3, 4, 5], "output": [100],
            "layer0_coeffs": [[0x1c5a44b9fff1cdd1f6508815743b3fbe730541abdc9859d23e39522af242c1cb,
                               0x08503b011eb1fb7a3450ca7466f5271f4e8918c650bee9174052183f0c08864f,
                               0x0bb9ceb7c28dd6dd8daef32ca650f17fd5f3101f3b1a87a3bd9521acda31b52d],
                              [0x0, 0x08208c1e18c86c06691bf02c74fcf3e00921af341147c504af91352e15f63827,
                               0x1861a45a4a5944133b53d0855ef6dba01b650d9c33d74f0e0eb39f8a41e2a875]],
            "wb": [0x0724c23c1fb22b9c7a586efc9202a12cc9213a8cad6a956520d723efac8abbe2],
            "wc": [0x17a4997c1d397f8b7c7f9e340ca87b9cb1f91d9445fb40b437e21c82e9058f20],
            "claimed_sum": 0x2cfcb93579f7951067f7ee334a925a935d5087694e8d541572a8dd8ab6187307},
}


def _check_appendix_b():
    t = po.Transcript()
    t.append(b"boy")
    assert t.sample_random_challenge().hex() == appendix_b["transcript"]["sample"]
    assert t.random_challenge_as_field_element(po.P["BN254_FQ"]) == appendix_b["transcript"]["challenge"]
    b = appendix_b["basic"]
    c, r, ch, f = po.basic_prove(b["table"], po.P[b["field"]])
    assert (c, r, ch, f) == (b["claimed_sum"], b["round_polys"], b["challenges"], b["final"])
    g = appendix_b["gkr"]
    layers = [[po.Gate(*x) for x in l] for l in g["layers"]]
    pf = po.gkr_prove(layers, g["inputs"], po.P[g["field"]])
    assert pf.circuit_output == g["output"] and pf.sumcheck_proofs[0][1] == g["layer0_coeffs"]
    assert pf.wb_evaluations == g["wb"] and pf.wc_evaluations == g["wc"] and pf.claimed_sum == g["claimed_sum"]
    assert po.gkr_verify(layers, pf, g["inputs"], po.P[g["field"]])


def _generated():
    rng = random.Random(0xB200)
    out = {"basic": [], "product": [], "gkr": []}
    for field in ("BN254_FQ", "BLS12_381_FR"):
        p = po.P[field]
        for n in (0, 1, 3, 6):
            table = [rng.randrange(p) for _ in range(1 << n)]
            if n == 3:
                table[0], table[1], table[5] = 0, p - 1, 1      # edge values
            c, r, ch, f = po.basic_prove(table, p)
            assert po.basic_verify(table, c, r, p)
            out["basic"].append({"field": field, "table": table, "claimed_sum": c, "round_polys": r,
                                 "challenges": ch, "final": f})
        for (P_, D_, n) in ((2, 2, 1), (2, 2, 4), (2, 3, 3), (3, 2, 2)):
            sp = [[[rng.randrange(p) for _ in range(1 << n)] for _ in range(D_)] for _ in range(P_)]
            claimed = sum(po.sumpoly_elementwise(sp, p)) % p
            t = po.Transcript()
            polys, chals, fin = po.product_prove(sp, claimed, t, p)
            ok, _, last = po.product_verify(claimed, polys, po.Transcript(), p)
            assert ok and last == po.sumpoly_evaluate(sp, chals, p)
            out["product"].append({"field": field, "P": P_, "D": D_, "tables": sp, "claimed_sum": claimed,
                                   "coeffs": polys, "challenges": chals,
                                   "final_tables": [[t_[0] for t_ in prod] for prod in fin]})
        # reference-shaped random circuit, depth 3 (layer i: 2^i gates reading 2^(i+1) wires)
        depth = 3
        layers = []
        for i in range(depth):
            gates, seen = [], set()
            for o in range(1 << i):
                while True:
                    g = (rng.randrange(1 << (i + 1)), rng.randrange(1 << (i + 1)), o, rng.randrange(2))
                    if g not in seen:
                        seen.add(g)
                        gates.append(g)
                        break
            layers.append(gates)
        inputs = [rng.randrange(p) for _ in range(1 << depth)]
        gl = [[po.Gate(*x) for x in l] for l in layers]
        pf = po.gkr_prove(gl, inputs, p)
        assert po.gkr_verify(gl, pf, inputs, p)
        out["gkr"].append({"field": field, "layers": [[list(g) for g in l] for l in layers], "inputs": inputs,
                           "output": pf.circuit_output, "claimed_sum": pf.claimed_sum,
                           "sumcheck": [{"claimed_sum": c, "coeffs": polys, "challenges": ch}
                                        for (c, polys, ch) in pf.sumcheck_proofs],
                           "wb": pf.wb_evaluations, "wc": pf.wc_evaluations})
    return out


if __name__ == "__main__":
    _check_reference_kats()
    _check_appendix_b()
    doc = {"reference_kats": reference_kats, "appendix_b": appendix_b, "generated": _generated()}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.json")
    with open(path, "w") as f:
        json.dump(doc, f, indent=0, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes")

"""The oracle against the reference's own known-answer tests, the survey's appendix-B vectors and the
independent Python model.  CPU only."""
import hashlib
import random

import numpy as np
import pytest

import pyoracle as po
from conftest import FIELDS


def ints(co, fid, a):
    return co.to_ints(fid, a)


def test_keccak_public_vectors(co):
    po.self_check()
    assert co.keccak256(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    assert co.keccak256(b"abc").hex() == "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45"
    rng = random.Random(1)
    for n in (1, 55, 135, 136, 137, 271, 272, 273, 1000):
        m = bytes(rng.randrange(256) for _ in range(n))
        assert co.keccak256(m) == po.keccak256(m)


def test_field_against_python_ints(co):
    rng = random.Random(2)
    for name, fid in FIELDS.items():
        p = po.P[name]
        edge = [0, 1, 2, p - 1, p - 2, (p + 1) // 2, (1 << 64) - 1, 1 << 64, (1 << 128) + 12345]
        vals = edge + [rng.randrange(p) for _ in range(60)]
        A = co.from_ints(fid, vals)
        assert co.to_ints(fid, A) == vals
        for i in range(len(vals)):
            a, b = vals[i], vals[(i * 7 + 3) % len(vals)]
            x, y = A[i], A[(i * 7 + 3) % len(vals)]
            assert ints(co, fid, co.fe_op("add", fid, x, y))[0] == (a + b) % p
            assert ints(co, fid, co.fe_op("sub", fid, x, y))[0] == (a - b) % p
            assert ints(co, fid, co.fe_op("mul", fid, x, y))[0] == (a * b) % p
        one = co.from_ints(fid, [1])[0]
        assert list(one) == [int(v) for v in np.array(one)]


def test_reference_kats(co, golden):
    k = golden["reference_kats"]
    for e in k["partial_evaluate"]:
        fid = FIELDS[e["field"]]
        out = co.mle_partial_evaluate(fid, co.from_ints(fid, e["table"]), e["var"], co.from_ints(fid, [e["r"]])[0])
        assert ints(co, fid, out) == e["out"], e["src"]
    for e in k["evaluate"]:
        fid = FIELDS[e["field"]]
        out = co.mle_evaluate(fid, co.from_ints(fid, e["table"]), co.from_ints(fid, e["values"]))
        assert ints(co, fid, out) == [e["out"]]
    for e in k["tensor"]:
        fid = FIELDS[e["field"]]
        out = co.tensor(fid, e["op"], co.from_ints(fid, e["wb"]), co.from_ints(fid, e["wc"]))
        assert ints(co, fid, out) == e["out"]
    with pytest.raises(AssertionError, match=k["tensor_panics"]["message"]):
        co.tensor(0, "mul", co.from_ints(0, k["tensor_panics"]["wb"]), co.from_ints(0, k["tensor_panics"]["wc"]))
    sp = k["sum_polynomial"]
    fid = FIELDS[sp["field"]]
    tabs = np.stack([np.stack([co.from_ints(fid, t) for t in prod]) for prod in sp["products"]])
    assert ints(co, fid, co.sumpoly_reduce(fid, tabs)) == sp["elementwise"]
    pp = k["product_polynomial"]
    tabs = np.stack([np.stack([co.from_ints(fid, t) for t in pp["polys"]])] * 2)   # (a*b) + (a*b)
    assert ints(co, fid, co.sumpoly_reduce(fid, tabs)) == [2 * v for v in pp["elementwise"]]
    for t, want in zip(pp["polys"], pp["fold_out"]):
        out = co.mle_partial_evaluate(fid, co.from_ints(fid, t), pp["fold_var"], co.from_ints(fid, [pp["fold_r"]])[0])
        assert ints(co, fid, out) == want
    ue = k["univariate_evaluate"]
    assert ints(co, fid, co.univariate_evaluate(fid, co.from_ints(fid, ue["coeffs"]), co.from_ints(fid, [ue["x"]])[0])) == [ue["out"]]
    lg = k["lagrange"]
    assert ints(co, fid, co.lagrange_interpolate(fid, co.from_ints(fid, lg["xs"]), co.from_ints(fid, lg["ys"]))) == lg["coeffs"]
    ru = k["round_univariate"]
    tabs = np.stack([np.stack([co.from_ints(fid, t) for t in prod]) for prod in ru["products"]])
    assert ints(co, fid, co.generate_round_univariate(fid, tabs)) == ru["evals"]
    rt = k["product_round_trip"]
    claimed = co.from_ints(fid, [rt["claimed_sum"]])[0]
    coeffs, chal, _ = co.product_prove(fid, tabs, claimed, co.Transcript())
    ok, chal2, _ = co.product_verify(fid, claimed, coeffs, co.Transcript())
    assert ok and np.array_equal(chal, chal2)
    for e in k["basic_claimed_sum"]:
        fid = FIELDS[e["field"]]
        claimed, _, _, _ = co.basic_prove(fid, co.from_ints(fid, e["table"]))
        assert ints(co, fid, claimed) == [e["sum"]]
    for e in k["basic_round_trips"]:
        fid = FIELDS[e["field"]]
        table = e.get("table") or [e["constant"]] * (1 << 12)   # the reference uses 2^20 of the same value; 2^12 here, 2^20 on the GPU test
        T = co.from_ints(fid, table)
        claimed, rp, _, _ = co.basic_prove(fid, T)
        assert co.basic_verify(fid, T, claimed, rp)
    with pytest.raises(AssertionError, match=k["new_panics"]["message"]):
        co.basic_prove(0, co.from_ints(0, k["new_panics"]["table"]))


def test_reference_circuit_kats(co, golden):
    k = golden["reference_kats"]
    for e in k["circuit_evaluate"]:
        fid = FIELDS[e["field"]]
        c = co.Circuit(e["layers"])
        ev = c.evaluate(fid, co.from_ints(fid, e["inputs"]))
        if "layer_evaluations" in e:
            assert [ints(co, fid, x) for x in ev] == e["layer_evaluations"]
        else:
            assert ints(co, fid, ev[0]) == e["output"]
    for i, v in k["num_of_layer_variables"]["values"]:
        assert co.lib().zko_num_of_layer_variables(i) == v
    for e in k["add_i_mul_i"]:
        c = co.Circuit(e["layers"])
        a, m = c.add_i_mul_i(0, e["layer"])
        assert a.shape[0] == e["size"]
        assert [i for i, v in enumerate(ints(co, 0, a)) if v] == e["add_ones"]
        assert [i for i, v in enumerate(ints(co, 0, m)) if v] == e["mul_ones"]
        assert all(v in (0, 1) for v in ints(co, 0, a))
    for e in k["gkr_round_trips"]:
        fid = FIELDS[e["field"]]
        c = co.Circuit(e["layers"])
        I = co.from_ints(fid, e["inputs"])
        pf = co.gkr_prove(fid, c, I)
        assert co.gkr_verify(fid, c, pf, I)
        # a tampered proof must be rejected
        pf.coeffs[0, 1, 0] ^= np.uint64(1)
        assert not co.gkr_verify(fid, c, pf, I)


def test_appendix_b_vectors(co, golden):
    b = golden["appendix_b"]
    t = co.Transcript()
    t.append(b["transcript"]["append"].encode())
    assert t.sample_random_challenge().hex() == b["transcript"]["sample"]
    assert ints(co, 0, t.random_challenge_as_field_element(0)) == [b["transcript"]["challenge"]]
    e = b["basic"]
    fid = FIELDS[e["field"]]
    claimed, rp, ch, fin = co.basic_prove(fid, co.from_ints(fid, e["table"]))
    assert ints(co, fid, claimed) == [e["claimed_sum"]]
    assert [ints(co, fid, r) for r in rp] == e["round_polys"]
    assert ints(co, fid, ch) == e["challenges"]
    assert ints(co, fid, fin) == [e["final"]]
    g = b["gkr"]
    fid = FIELDS[g["field"]]
    c = co.Circuit(g["layers"])
    pf = co.gkr_prove(fid, c, co.from_ints(fid, g["inputs"]))
    assert ints(co, fid, pf.circuit_output) == g["output"]
    assert [ints(co, fid, x) for x in pf.coeffs[:2]] == g["layer0_coeffs"]
    assert ints(co, fid, pf.wb[:1]) == g["wb"] and ints(co, fid, pf.wc[:1]) == g["wc"]
    assert ints(co, fid, pf.claimed_sum) == [g["claimed_sum"]]


def test_generated_vectors_match_c_oracle(co, golden):
    gen = golden["generated"]
    for e in gen["basic"]:
        fid = FIELDS[e["field"]]
        T = co.from_ints(fid, e["table"])
        claimed, rp, ch, fin = co.basic_prove(fid, T)
        assert ints(co, fid, claimed) == [e["claimed_sum"]]
        assert [ints(co, fid, r) for r in rp] == e["round_polys"]
        assert ints(co, fid, ch) == e["challenges"] and ints(co, fid, fin) == [e["final"]]
        assert co.basic_verify(fid, T, claimed, rp)
    for e in gen["product"]:
        fid = FIELDS[e["field"]]
        tabs = np.stack([np.stack([co.from_ints(fid, t) for t in prod]) for prod in e["tables"]])
        coeffs, ch, fin = co.product_prove(fid, tabs, co.from_ints(fid, [e["claimed_sum"]])[0], co.Transcript())
        assert [ints(co, fid, c) for c in coeffs] == e["coeffs"]
        assert ints(co, fid, ch) == e["challenges"]
        assert [ints(co, fid, f) for f in fin] == e["final_tables"]
    for e in gen["gkr"]:
        fid = FIELDS[e["field"]]
        c = co.Circuit(e["layers"])
        I = co.from_ints(fid, e["inputs"])
        pf = co.gkr_prove(fid, c, I)
        assert co.gkr_verify(fid, c, pf, I)
        assert ints(co, fid, pf.circuit_output) == e["output"]
        assert ints(co, fid, pf.claimed_sum) == [e["claimed_sum"]]
        flat = [cf for sc in e["sumcheck"] for poly in sc["coeffs"] for cf in poly]
        assert ints(co, fid, pf.coeffs) == flat
        assert ints(co, fid, pf.layer_claims) == [sc["claimed_sum"] for sc in e["sumcheck"]]
        assert ints(co, fid, pf.wb[:len(e["wb"])]) == e["wb"] and ints(co, fid, pf.wc[:len(e["wc"])]) == e["wc"]


def test_random_cross_check_python_vs_c(co):
    rng = random.Random(7)
    for name, fid in FIELDS.items():
        p = po.P[name]
        for n in (2, 5):
            for var in range(n):
                tab = [rng.randrange(p) for _ in range(1 << n)]
                r = rng.randrange(p)
                out = co.mle_partial_evaluate(fid, co.from_ints(fid, tab), var, co.from_ints(fid, [r])[0])
                assert ints(co, fid, out) == po.partial_evaluate(tab, var, r, p)
        sp = [[[rng.randrange(p) for _ in range(16)] for _ in range(2)] for _ in range(2)]
        claimed = sum(po.sumpoly_elementwise(sp, p)) % p
        polys, chals, _ = po.product_prove(sp, claimed, po.Transcript(), p)
        tabs = np.stack([np.stack([co.from_ints(fid, t) for t in prod]) for prod in sp])
        coeffs, ch, _ = co.product_prove(fid, tabs, co.from_ints(fid, [claimed])[0], co.Transcript())
        assert [ints(co, fid, c) for c in coeffs] == polys and ints(co, fid, ch) == chals
        # a wrong claimed sum is caught by the verifier in round 0
        ok, _, _ = co.product_verify(fid, co.from_ints(fid, [(claimed + 1) % p])[0], coeffs, co.Transcript())
        assert not ok


def test_generalised_python_gkr_model_reduces_to_the_reference_shape():
    """pyoracle.gkr_prove_general (extension oracle for wide layers) == pyoracle.gkr_prove on reference shapes"""
    rng = random.Random(11)
    p = po.P["BN254_FQ"]
    for depth in (1, 2, 3, 4):
        layers = []
        for i in range(depth):
            layers.append([po.Gate(rng.randrange(2 << i), rng.randrange(2 << i), o, rng.randrange(2)) for o in range(1 << i)])
        inputs = [rng.randrange(p) for _ in range(1 << depth)]
        a = po.gkr_prove(layers, inputs, p)
        b = po.gkr_prove_general(layers, [1] + list(range(1, depth + 1)), inputs, p)
        assert a.sumcheck_proofs == b.sumcheck_proofs and a.claimed_sum == b.claimed_sum
        assert a.wb_evaluations == b.wb_evaluations and a.wc_evaluations == b.wc_evaluations
        assert po.gkr_verify(layers, a, inputs, p)


def test_all_core_mode_gives_the_same_proofs(co):
    """zko_set_threads(n > 1) only changes who runs the data-parallel loops (the labelled all-core CPU baseline of
    bench.py); every output is the same field element as in the single-threaded, reference-shaped mode"""
    import numpy as np
    fid = 0
    rng = np.random.default_rng(3)
    n = 1 << 13
    raw = rng.integers(0, 1 << 62, size=(2, 2, n, 4), dtype=np.uint64)
    raw[..., 3] &= np.uint64((1 << 58) - 1)
    claimed = np.zeros(4, dtype=np.uint64)
    co.lib().zko_fe_sum(fid, co._p(co.sumpoly_reduce(fid, raw)), n, co._p(claimed))
    try:
        outs = []
        for threads in (1, 4):
            co.set_threads(threads)
            assert co.get_threads() == threads
            c2 = np.zeros(4, dtype=np.uint64)
            co.lib().zko_fe_sum(fid, co._p(co.sumpoly_reduce(fid, raw)), n, co._p(c2))
            assert np.array_equal(c2, claimed)
            outs.append((co.product_prove(fid, raw, claimed, co.Transcript()), co.basic_prove(fid, raw[0, 0]),
                         co.mle_evaluate(fid, raw[0, 1], raw[1, 0][:13]), co.mle_to_bytes(fid, raw[1, 1])))
    finally:
        co.set_threads(1)
    (p1, b1, e1, y1), (p4, b4, e4, y4) = outs
    assert all(np.array_equal(x, y) for x, y in zip(p1, p4)) and all(np.array_equal(x, y) for x, y in zip(b1, b4))
    assert np.array_equal(e1, e4) and y1 == y4


@pytest.mark.parametrize("fid", [0, 2])
def test_gate_list_gkr_oracle_equals_the_dense_restatement_on_reference_shapes(co, fid):
    """zko_gkr_prove_sparse / zko_gkr_verify_sparse (gkr_protocol.rs:26-236 with add_i / mul_i evaluated from the gate list,
    O(gates) per round) are the checkers of the wide CUDA prover; here they are pinned to the dense line-by-line restatement
    zko_gkr_prove / zko_gkr_verify on circuits the reference's storage can hold, and to the Python model beyond."""
    import random
    import pyoracle as po
    name = {0: "BN254_FQ", 2: "BLS12_381_FR"}[fid]
    p = po.P[name]
    for depth in (1, 2, 3, 5, 6):
        rng = random.Random(7 * fid + depth)
        layers = []
        for i in range(depth):
            gates, seen = [], set()
            for o in range(1 if i == 0 else 1 << i):
                for _ in range(rng.choice([1, 1, 2, 3])):
                    g = (rng.randrange(1 << (i + 1)), rng.randrange(1 << (i + 1)), o, rng.randrange(2))
                    if g not in seen:
                        seen.add(g)
                        gates.append(g)
            layers.append(gates)
        I = co.from_ints(fid, [rng.randrange(p) for _ in range(1 << depth)])
        want = co.gkr_prove(fid, co.Circuit(layers), I)
        sc = co.SparseCircuit([0] + list(range(1, depth + 1)), layers)       # one output: the reference's padded layer 0
        got = co.gkr_prove_sparse(fid, sc, I)
        for f in ("coeffs", "challenges", "layer_claims", "wb", "wc", "claimed_sum", "circuit_output"):
            assert np.array_equal(getattr(got, f), getattr(want, f)), (depth, f)
        assert co.gkr_verify_sparse(fid, sc, got, I) and co.gkr_verify_sparse(fid, sc, want, I)
        got.coeffs[0, 0, 0] ^= np.uint64(1)
        assert not co.gkr_verify_sparse(fid, sc, got, I)
    if fid == 0:
        for bits in ([2, 3, 3, 2], [1, 4, 2, 5], [4, 1, 3]):
            rng = random.Random(sum(b * 7 ** i for i, b in enumerate(bits)))
            layers = []
            for li in range(len(bits) - 1):
                gates, seen = [], set()
                for o in range(1 << bits[li]):
                    for _ in range(rng.choice((1, 1, 2))):
                        g = (rng.randrange(1 << bits[li + 1]), rng.randrange(1 << bits[li + 1]), o, rng.randrange(2))
                        if g not in seen:
                            seen.add(g)
                            gates.append(g)
                layers.append(gates)
            inputs = [rng.randrange(p) for _ in range(1 << bits[-1])]
            want = po.gkr_prove_general([[po.Gate(*g) for g in l] for l in layers], bits, inputs, p)
            sc = co.SparseCircuit(bits, layers)
            got = co.gkr_prove_sparse(fid, sc, co.from_ints(fid, inputs))
            assert co.to_ints(fid, got.coeffs) == [c for (_, polys, _) in want.sumcheck_proofs for poly in polys for c in poly]
            assert co.to_ints(fid, got.challenges) == [c for (_, _, ch) in want.sumcheck_proofs for c in ch]
            assert co.to_ints(fid, got.claimed_sum) == [want.claimed_sum]
            assert co.gkr_verify_sparse(fid, sc, got, co.from_ints(fid, inputs))


def test_seeded_table_generator_matches_the_python_restatement(co):
    """zko_table_generate (the CPU arm's inputs in bench.py) against the pure-Python statement of SURVEY.md 8d that the
    GPU tests use to check zk_table_generate: all three agree, so both bench arms prove the same tables."""
    from zk_cryptography_research_implementations_b200.core import synthetic_table_ints
    for fid in (0, 1, 2):
        for first, step in ((0, 1), (5, 8)):
            assert co.to_ints(fid, co.table_generate(fid, 0xB200, 3, 64, first, step)) == synthetic_table_ints(fid, 0xB200, 3, 64, first, step)
    co.set_threads(4)
    try:
        a = co.table_generate(0, 7, 1, 1 << 13)
    finally:
        co.set_threads(1)
    assert np.array_equal(a, co.table_generate(0, 7, 1, 1 << 13))

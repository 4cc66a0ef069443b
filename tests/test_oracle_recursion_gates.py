"""Recursion gate set (the gates `verify_proof` instantiates at
/root/reference/wormhole/aggregator/src/circuits/tree.rs:119) in the oracle: CommonCircuitData round trip with the
new gate payloads, witness satisfaction, proofs accepted by the restated verifier, and every new gate's constraints
actually bite (a tampered wire of each gate kind gives a proof the verifier rejects).

PARITY UNPINNED for these eight gates (oracle/gates.hpp header): the reference tree holds no recursion-circuit
fixture, so this file proves self-consistency of prover and verifier, not agreement with qp-plonky2. CPU only."""
import numpy as np
import pytest

UNUSED = 0xFFFFFFFF
# gate order of the recursion-shaped circuit (sorted by degree, then id) and a wire of each new gate that only the
# gate's own constraints see (no copy constraint on it)
GATE_ORDER = ["Noop", "Constant", "PoseidonMds", "PublicInput", "BaseSum", "ReducingExtension", "Reducing",
              "ArithmeticExtension", "Arithmetic", "MulExtension", "Exponentiation", "RandomAccess", "CosetInterpolation",
              "Poseidon"]
PRIVATE_WIRE = {"ReducingExtension": 70, "Reducing": 50, "Exponentiation": 70, "RandomAccess": 75, "CosetInterpolation": 45,
                "PoseidonMds": 30, "ArithmeticExtension": 7, "MulExtension": 5}


@pytest.fixture(scope="module")
def rec(oracle):
    s = oracle.Synth(seed=3, **oracle.Synth.RECURSION_TINY)
    return s, oracle.Circuit(s.common, s.const_sigma_values)


def gate_of_row(s):
    sel = s.const_sigma_values[:4]
    g = np.full(s.n, -1, dtype=np.int64)
    for k in range(4):
        used = sel[k] != UNUSED
        g[used] = sel[k][used].astype(np.int64)
    return g


def test_recursion_circuit_shape(oracle, rec):
    s, _ = rec
    assert s.info["num_gates"] == 14 and s.info["num_constants"] == 6 and s.info["zero_knowledge"] == 0
    assert s.check() == ""
    g = gate_of_row(s)
    assert set(range(14)) <= set(int(x) for x in g), "every gate of the set has at least one row"
    # upstream's greedy selector grouping for these degrees: [0,7) [7,11) [11,13) [13,14)
    sel = s.const_sigma_values[:4]
    for gi, grp in zip(range(14), [0] * 7 + [1] * 4 + [2] * 2 + [3]):
        rows = np.nonzero(g == gi)[0]
        assert np.all(sel[grp][rows] == gi)
        for other in range(4):
            if other != grp:
                assert np.all(sel[other][rows] == UNUSED)


def test_recursion_proof_accepted(oracle, rec):
    s, c = rec
    proof = c.prove(s.wires, s.public_inputs)
    assert c.verify(proof) == ""
    assert oracle.proof_roundtrip(s.common, proof) == proof
    assert c.prove(s.wires, s.public_inputs) == proof


@pytest.mark.parametrize("gate", sorted(PRIVATE_WIRE))
def test_each_recursion_gate_rejects_a_tampered_wire(oracle, rec, gate):
    s, c = rec
    rows = np.nonzero(gate_of_row(s) == GATE_ORDER.index(gate))[0]
    assert len(rows) > 0
    w = s.wires.copy()
    w[PRIVATE_WIRE[gate], rows[0]] ^= np.uint64(1)
    assert c.verify(c.prove(w, s.public_inputs)) != ""


def test_zero_knowledge_recursion_circuit(oracle):
    """The aggregator builds its chunk circuits with the leaf circuit's zk config (aggregator.rs:21, tree.rs:111): blinding
    rows push the degree to 2^14 and every batch is salted."""
    s = oracle.Synth(zk=True, seed=3, **oracle.Synth.RECURSION_TINY)
    assert s.check() == "" and s.info["degree_bits"] == 14 and s.info["zero_knowledge"] == 1
    c = oracle.Circuit(s.common, s.const_sigma_values)
    proof = c.prove(s.wires, s.public_inputs, salt_seed=9)
    assert c.verify(proof) == "" and len(proof) == 149324
    assert oracle.proof_roundtrip(s.common, proof) == proof


def test_larger_recursion_circuit(oracle):
    s = oracle.Synth(seed=11, n_poseidon=60, n_base_sum=10, n_arith=20, n_const=6, num_public_inputs=16, n_arith_ext=40,
                     n_mul_ext=10, n_reducing=8, n_reducing_ext=8, n_random_access=12, n_exp=6, n_coset=8, n_mds=3)
    assert s.check() == "" and s.info["degree_bits"] == 8
    c = oracle.Circuit(s.common, s.const_sigma_values)
    assert c.verify(c.prove(s.wires, s.public_inputs)) == ""


# ---------------------------------------------------------------------------------------------------------------------
# Layout-independent SEMANTIC pins. The eight formulas cannot be compared with qp-plonky2 here, but WHAT each gate computes is
# public knowledge: on rows the constraint checker accepts, the gate's output wires must equal the mathematical function of
# its input wires, evaluated below by independent big-integer arithmetic (plain Lagrange interpolation, Horner's rule,
# square-and-multiply, list indexing, the MDS matrix of SURVEY A.2). A wrong barycentric weight, accumulator order, bit order
# or matrix index in oracle/gates.hpp would satisfy its own constraints and still fail here. What stays unpinned is only the
# wire ORDER inside a row (which wire is "alpha", which is "old_acc"), which needs the L4 fixture.
# ---------------------------------------------------------------------------------------------------------------------
P = 0xFFFFFFFF00000001
W = 7
MDS_C = [17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20]
MDS_D = [8] + [0] * 11


def e_add(x, y): return ((x[0] + y[0]) % P, (x[1] + y[1]) % P)
def e_sub(x, y): return ((x[0] - y[0]) % P, (x[1] - y[1]) % P)
def e_mul(x, y): return ((x[0] * y[0] + W * x[1] * y[1]) % P, (x[0] * y[1] + x[1] * y[0]) % P)
def e_scale(x, s): return (x[0] * s % P, x[1] * s % P)
def e_inv(x):
    ninv = pow((x[0] * x[0] - W * x[1] * x[1]) % P, P - 2, P)
    return (x[0] * ninv % P, (-x[1]) * ninv % P)


def rows_of(s, gate):
    return np.nonzero(gate_of_row(s) == GATE_ORDER.index(gate))[0]


def row_wires(s, row):
    return [int(x) for x in s.wires[:, row]]


def row_consts(s, row):
    nsel = 4
    return [int(x) for x in s.const_sigma_values[nsel:s.info["num_constants"], row]]


@pytest.fixture(scope="module")
def big(oracle):
    s = oracle.Synth(seed=17, n_poseidon=20, n_base_sum=4, n_arith=6, n_const=4, num_public_inputs=8, n_arith_ext=12,
                     n_mul_ext=9, n_reducing=9, n_reducing_ext=9, n_random_access=10, n_exp=9, n_coset=9, n_mds=5)
    assert s.check() == ""
    return s


def test_semantics_arithmetic_and_mul_extension(big):
    for gate, per, nops in (("ArithmeticExtension", 8, 10), ("MulExtension", 6, 13)):
        rows = rows_of(big, gate)
        assert len(rows) >= 5
        for r in rows:
            w, k = row_wires(big, r), row_consts(big, r)
            for i in range(nops):
                a, b = (w[per * i], w[per * i + 1]), (w[per * i + 2], w[per * i + 3])
                want = e_scale(e_mul(a, b), k[0])
                if gate == "ArithmeticExtension":
                    want = e_add(want, e_scale((w[per * i + 4], w[per * i + 5]), k[1]))
                assert (w[per * i + per - 2], w[per * i + per - 1]) == want


@pytest.mark.parametrize("gate,nc,ext", [("Reducing", 43, False), ("ReducingExtension", 32, True)])
def test_semantics_reducing_is_horner(big, gate, nc, ext):
    rows = rows_of(big, gate)
    assert len(rows) >= 5
    for r in rows:
        w = row_wires(big, r)
        out, alpha, acc = (w[0], w[1]), (w[2], w[3]), (w[4], w[5])
        coeffs = [(w[6 + 2 * i], w[7 + 2 * i]) for i in range(nc)] if ext else [(w[6 + i], 0) for i in range(nc)]
        assert any(c != (0, 0) for c in coeffs) and alpha != (0, 0)
        for c in coeffs:                                   # Horner: ((old * a + c0) * a + c1) * a + ...
            acc = e_add(e_mul(acc, alpha), c)
        assert out == acc


def test_semantics_exponentiation(big):
    rows = rows_of(big, "Exponentiation")
    assert len(rows) >= 5
    for r in rows:
        w = row_wires(big, r)
        nb = 66
        bits = w[1:1 + nb]
        assert set(bits) <= {0, 1} and any(bits)
        exponent = sum(b << i for i, b in enumerate(bits))          # little-endian power bits
        assert w[nb + 1] == pow(w[0], exponent, P)


def test_semantics_random_access(big):
    rows = rows_of(big, "RandomAccess")
    assert len(rows) >= 5
    seen = set()
    for r in rows:
        w, k = row_wires(big, r), row_consts(big, r)
        for cpy in range(4):
            base = 18 * cpy
            index, claimed, items = w[base], w[base + 1], w[base + 2: base + 18]
            assert index < 16 and claimed == items[index]
            seen.add(index)
        for j in range(2):                                         # the two extra constant slots carry the gate constants
            assert w[72 + j] == k[j]
    assert len(seen) >= 6


def test_semantics_coset_interpolation_is_lagrange(oracle, big):
    rows = rows_of(big, "CosetInterpolation")
    assert len(rows) >= 5
    w16 = oracle.root_of_unity(4)
    for r in rows:
        w = row_wires(big, r)
        shift = w[0]
        vals = [(w[1 + 2 * k], w[2 + 2 * k]) for k in range(16)]
        point, value = (w[33], w[34]), (w[35], w[36])
        xs = [shift * pow(w16, k, P) % P for k in range(16)]
        acc = (0, 0)
        for k in range(16):                                        # plain Lagrange over the coset shift * <w16>
            num, den = (1, 0), 1
            for m in range(16):
                if m != k:
                    num = e_mul(num, e_sub(point, (xs[m], 0)))
                    den = den * ((xs[k] - xs[m]) % P) % P
            acc = e_add(acc, e_mul(vals[k], e_scale(num, pow(den, P - 2, P))))
        assert shift != 0 and value == acc


def test_semantics_poseidon_mds(big):
    rows = rows_of(big, "PoseidonMds")
    assert len(rows) >= 3
    for r in rows:
        w = row_wires(big, r)
        for comp in range(2):
            x = [w[2 * i + comp] for i in range(12)]
            for o in range(12):
                want = (sum(x[(i + o) % 12] * MDS_C[i] for i in range(12)) + x[o] * MDS_D[o]) % P
                assert w[24 + 2 * o + comp] == want

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

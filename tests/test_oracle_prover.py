"""Oracle prover stages are pinned *through the verifier* (oracle/prover.hpp header): every proof the
restated prover emits on a synthetic wormhole-/voting-shaped circuit must be accepted by the restated
verifier, which is itself pinned on the reference's bench-data fixture. CPU only, small sizes."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def tiny(oracle):
    s = oracle.Synth(zk=False, seed=3, **oracle.Synth.TINY)
    return s, oracle.Circuit(s.common, s.const_sigma_values)


def test_synthetic_witness_satisfies_circuit(oracle, tiny):
    s, _ = tiny
    assert s.check() == ""
    assert s.info["num_gates"] == 6 and s.info["num_wires"] == 135


@pytest.mark.parametrize("zk", [False, True])
def test_oracle_proof_accepted_by_pinned_verifier(oracle, zk):
    s = oracle.Synth(zk=zk, seed=5, **oracle.Synth.TINY)
    c = oracle.Circuit(s.common, s.const_sigma_values)
    proof = c.prove(s.wires, s.public_inputs, salt_seed=99)
    assert c.verify(proof) == ""
    assert oracle.proof_roundtrip(s.common, proof) == proof
    # deterministic: same witness, salts and PoW rule -> same bytes
    assert c.prove(s.wires, s.public_inputs, salt_seed=99) == proof
    if zk:
        assert c.prove(s.wires, s.public_inputs, salt_seed=100) != proof


def test_voting_shaped_circuit(oracle):
    s = oracle.Synth(zk=False, seed=2, **oracle.Synth.VOTING)
    assert s.info["degree_bits"] == 8 and s.info["num_public_inputs"] == 13
    c = oracle.Circuit(s.common, s.const_sigma_values)
    assert c.verify(c.prove(s.wires, s.public_inputs)) == ""


def test_explicit_salts_match_seeded_salts(oracle):
    s = oracle.Synth(zk=True, seed=8, **oracle.Synth.TINY)
    c = oracle.Circuit(s.common, s.const_sigma_values)
    N = s.n * 8
    salts = np.zeros((3, 4, N), dtype=np.uint64)
    for b in range(3):
        for k in range(4):
            for l in range(0, N, max(1, N // 64)):
                salts[b, k, l] = oracle.salt_value(123, b, k, l)
    # fill everything (vectorised restatement of splitmix64 finalizer)
    def mix(z):
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))
    with np.errstate(over="ignore"):
        for b in range(3):
            for k in range(4):
                z = np.uint64(123) + np.uint64(0x9E3779B97F4A7C15) * np.uint64(b * 4 + k + 1) + np.arange(N, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15)
                v = mix(z)
                v = np.where(v >= np.uint64(oracle.P), v - np.uint64(oracle.P), v)
                probe = salts[b, k] != 0
                assert np.array_equal(v[probe], salts[b, k][probe])
                salts[b, k] = v
    assert c.prove(s.wires, s.public_inputs, salts=salts) == c.prove(s.wires, s.public_inputs, salt_seed=123)


def test_tampered_witness_is_not_provable(oracle, tiny):
    # the reference surfaces bad inputs as prove Err / unverifiable proofs (voting/src/lib.rs:386-434)
    s, c = tiny
    w = s.wires.copy()
    w[3, 2] ^= np.uint64(1)
    assert c.verify(c.prove(w, s.public_inputs)) != ""
    pis = s.public_inputs.copy()
    pis[0] ^= np.uint64(1)
    assert c.verify(c.prove(s.wires, pis)) != ""


def test_stage_functions_match_full_prover_trace(oracle, tiny):
    s, c = tiny
    proof, tr = c.prove(s.wires, s.public_inputs, trace=True)
    zs = c.partial_products(s.wires, tr.betas, tr.gammas)
    assert np.array_equal(zs, tr.zs_pp_values)
    assert np.all(zs[0:2, 0] == 1)                       # Z(1) = 1
    q = c.quotient(s.wires, zs, s.public_inputs, tr.betas, tr.gammas, tr.alphas)
    assert np.array_equal(q, tr.quotient_chunks)


def test_lde_definition(oracle):
    # lde_out[c][l] = P_c(g * w^bitrev(l)); check a few leaves by direct evaluation
    rng = np.random.default_rng(1)
    n, rb = 16, 3
    vals = rng.integers(0, oracle.P, size=(2, n), dtype=np.uint64)
    coeffs, lde = oracle.lde_batch(vals, rb)
    assert np.array_equal(oracle.ntt(coeffs), vals)
    N, lg = n << rb, 7
    g, w = 0xC65C18B67785D900, oracle.root_of_unity(lg)
    for l in (0, 1, 5, 77, 127):
        i = int(format(l, "07b")[::-1], 2)
        x = g * pow(w, i, oracle.P) % oracle.P
        for c in range(2):
            acc = 0
            for k in reversed(range(n)):
                acc = (acc * x + int(coeffs[c, k])) % oracle.P
            assert acc == int(lde[c, l])


@pytest.mark.parametrize("spec,zk", [("TINY", True), ("VOTING", False), ("RECURSION_TINY", True)])
def test_fast_mode_proof_is_byte_identical(oracle, spec, zk):
    """The CPU-baseline arm (oracle.set_fast(True): AVX2 Poseidon, coset-wise LDE, the gate constraints evaluated four LDE points
    per call through the 4-lane Ops of oracle/vec_ops.hpp) emits the same proof bytes as the readable restatement — on the
    6-gate set and on the 14-gate recursion set (every gate's generic code instantiated over the vector type)."""
    s = oracle.Synth(zk=zk, seed=5, **getattr(oracle.Synth, spec))
    c = oracle.Circuit(s.common, s.const_sigma_values)
    try:
        oracle.set_fast(False)
        want = c.prove(s.wires, s.public_inputs, salt_seed=3)
        oracle.set_fast(True)
        got = c.prove(s.wires, s.public_inputs, salt_seed=3)
        assert got == want and c.verify(got) == ""
    finally:
        oracle.set_fast(False)

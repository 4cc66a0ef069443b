"""End-to-end parity: proofs produced by the CUDA prover through the C ABI are byte-identical to the
oracle prover's on the same witness, salts and PoW rule, and are accepted by the pinned verifier
(SURVEY.md §8c levels L2 + L3)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def zkb():
    import zkb200

    if zkb200.device_count() == 0:
        pytest.fail("no CUDA device: -m gpu tests must run on the B200 box")
    return zkb200


def run_case(zkb, oracle, spec, zk, seed, salt_seed=77, min_degree_bits=0):
    s = oracle.Synth(zk=zk, seed=seed, min_degree_bits=min_degree_bits, **spec)
    oc = oracle.Circuit(s.common, s.const_sigma_values)
    gc = zkb.ProverCircuit(s.common, s.const_sigma_values, is_values=True, circuit_digest=oc.digest)
    want = oc.prove(s.wires, s.public_inputs, salt_seed=salt_seed)
    got = gc.prove(s.wires, s.public_inputs, salt_seed=salt_seed)
    assert len(got) == gc.proof_size == len(want)
    assert oc.verify(got) == "", "CUDA proof rejected by the pinned verifier"
    assert got == want, "CUDA proof bytes differ from the oracle prover"
    return s, oc, gc, got


@pytest.mark.parametrize("zk", [False, True])
def test_tiny_circuit_proof_bytes(zkb, oracle, zk):
    s, oc, gc, proof = run_case(zkb, oracle, oracle.Synth.TINY, zk, seed=5)
    # deterministic and re-usable context
    assert gc.prove(s.wires, s.public_inputs, salt_seed=77) == proof
    t = gc.timings()
    assert t["total"] > 0


def test_voting_shaped_proof_bytes(zkb, oracle):
    run_case(zkb, oracle, oracle.Synth.VOTING, False, seed=2)


def test_voting_shaped_proof_bytes_at_the_surveyed_degree(zkb, oracle):
    """Config #2 as SURVEY.md §8d states it: n = 2^9, non-zk, 13 public inputs (voting/src/lib.rs:72-76)."""
    s, oc, gc, proof = run_case(zkb, oracle, oracle.Synth.VOTING, False, seed=2, min_degree_bits=9)
    assert s.info["degree_bits"] == 9 and s.info["num_public_inputs"] == 13


def test_recursion_gate_set_proof_bytes(zkb, oracle):
    """Config #4/#5 gate set (the 14 gates of a recursive-verifier circuit, 4 selector groups): the CUDA
    quotient's third launch evaluates ArithmeticExtension, MulExtension, Reducing(Extension), RandomAccess, Exponentiation,
    CosetInterpolation and PoseidonMds; proof bytes equal the oracle's. (These gates are unpinned against qp-plonky2 —
    oracle/gates.hpp — so this is GPU-vs-oracle parity plus acceptance by the restated verifier.)"""
    s, oc, gc, proof = run_case(zkb, oracle, oracle.Synth.RECURSION_TINY, False, seed=3)
    assert s.info["num_gates"] == 14
    # a tampered wire of a recursion gate gives the same (unverifiable) bytes on both sides
    w = s.wires.copy()
    w[70, int(np.nonzero(s.const_sigma_values[0] == 5)[0][0])] ^= np.uint64(1)      # ReducingExtension accumulator
    bad = gc.prove(w, s.public_inputs, salt_seed=77)
    assert bad == oc.prove(w, s.public_inputs, salt_seed=77) and oc.verify(bad) != ""


def test_recursion_shaped_circuit_proof_bytes(zkb, oracle):
    """A recursion-shaped circuit at the size class of one aggregation chunk, NON-zk: n = 2^12. (The reference's own tree tests
    build their toy circuits with the non-zk `standard_recursion_config`, tree.rs:165; the production aggregator inherits the
    leaf circuit's zk config — aggregator.rs:21, tree.rs:111 — which is the next test's shape.)"""
    s, oc, gc, proof = run_case(zkb, oracle, oracle.Synth.RECURSION, False, seed=4)
    assert s.info["degree_bits"] == 12 and len(proof) == gc.proof_size


def test_zero_knowledge_recursion_circuit_proof_bytes(zkb, oracle):
    """The aggregator's chunk circuits inherit the leaf circuit's zk config (aggregator.rs:21, tree.rs:111): 14-gate set,
    salted batches, n = 2^14."""
    s, oc, gc, proof = run_case(zkb, oracle, oracle.Synth.RECURSION_TINY, True, seed=3)
    assert s.info["degree_bits"] == 14 and s.info["num_gates"] == 14


@pytest.mark.parametrize("npi", [0, 1, 8, 9, 17])
def test_minimal_degree_and_public_input_counts(zkb, oracle, npi):
    """Edge shapes: the smallest circuits the generator makes (n = 8: 64 LDE points, no FRI reduction layer, a 16-digest cap
    two levels above the leaves) with 0 / 1 / a full sponge block / one more / two blocks and one more public inputs
    (hash_no_pad([]) is the zero digest: no permutation)."""
    spec = dict(n_poseidon=3, n_base_sum=1, n_arith=1, n_const=2, num_public_inputs=npi)
    s, oc, gc, proof = run_case(zkb, oracle, spec, False, seed=2)
    assert s.info["degree_bits"] == 3 and s.info["reduction_arity_bits"] == []


def test_explicit_salts(zkb, oracle):
    s = oracle.Synth(zk=True, seed=8, **oracle.Synth.TINY)
    oc = oracle.Circuit(s.common, s.const_sigma_values)
    gc = zkb.ProverCircuit(s.common, s.const_sigma_values, is_values=True)
    rng = np.random.default_rng(3)
    salts = rng.integers(0, 0xFFFFFFFF00000001, size=(3, 4, s.n * 8), dtype=np.uint64)
    got = gc.prove(s.wires, s.public_inputs, salts=salts)
    assert got == oc.prove(s.wires, s.public_inputs, salts=salts)
    assert oc.verify(got) == ""


def test_bad_arguments(zkb, oracle):
    s = oracle.Synth(zk=False, seed=3, **oracle.Synth.TINY)
    gc = zkb.ProverCircuit(s.common, s.const_sigma_values, is_values=True)
    with pytest.raises(zkb.ZkbError) as e:
        gc.prove(s.wires, s.public_inputs[:-1])
    assert e.value.status == "ZKB_E_ARG"
    bad = s.public_inputs.copy()
    bad[0] = 0xFFFFFFFF00000001
    with pytest.raises(zkb.ZkbError) as e:
        gc.prove(s.wires, bad)
    assert e.value.status == "ZKB_E_ARG"


def test_tampered_witness_gives_unverifiable_proof(zkb, oracle):
    s = oracle.Synth(zk=False, seed=3, **oracle.Synth.TINY)
    oc = oracle.Circuit(s.common, s.const_sigma_values)
    gc = zkb.ProverCircuit(s.common, s.const_sigma_values, is_values=True)
    w = s.wires.copy()
    w[3, 2] ^= np.uint64(1)
    got = gc.prove(w, s.public_inputs)
    assert got == oc.prove(w, s.public_inputs)      # still bit-identical to the CPU prover ...
    assert oc.verify(got) != ""                     # ... and, like it, not accepted


def test_wormhole_shaped_nonzk_proof_bytes(zkb, oracle):
    """C1': wormhole-shaped circuit, non-zk, n = 2^13 — same size as aggregator/data/dummy_proof.bin."""
    s, oc, gc, proof = run_case(zkb, oracle, oracle.Synth.WORMHOLE, False, seed=1)
    assert s.info["degree_bits"] == 13 and len(proof) == 132712


def test_wormhole_shaped_zk_proof_bytes(zkb, oracle):
    """C1: wormhole-shaped circuit, zk, n = 2^14 — same size as wormhole/bench-data/proof.bin."""
    s, oc, gc, proof = run_case(zkb, oracle, oracle.Synth.WORMHOLE, True, seed=1)
    assert s.info["degree_bits"] == 14 and len(proof) == 148932


@pytest.mark.gpu
def test_batch_of_proofs_on_two_streams(zkb, oracle):
    """zkb200.batch (the aggregator-style fan-out over independent proofs): 5 proofs on 2 prover contexts / streams of one
    GPU, each byte-identical to the oracle prover's proof for the same salt seed."""
    from zkb200 import batch

    s = zkb.SynthCircuit(zk=True, seed=21, **zkb.TINY)
    so = oracle.Synth(zk=True, seed=21, **oracle.Synth.TINY)
    assert so.check() == "" and np.array_equal(s.wires, so.wires) and np.array_equal(s.public_inputs, so.public_inputs)
    provers = [zkb.ProverCircuit(s.common, s.const_sigma_values, is_values=True) for _ in range(2)]
    oc = oracle.Circuit(s.common, s.const_sigma_values)
    seeds = [7, 8, 9, 10, 11]
    proofs = batch.prove_batch(seeds, provers, lambda p, w, i: p.prove(s.wires, s.public_inputs, salt_seed=w))
    assert len(proofs) == len(seeds)
    assert np.array_equal(s.wires, so.wires), "the witness changed under the provers"
    for seed, proof in zip(seeds, proofs):
        assert proof == oc.prove(s.wires, s.public_inputs, salt_seed=seed)
        assert oc.verify(proof) == ""


@pytest.mark.gpu
def test_non_canonical_witness_is_rejected(zkb):
    """A wire value >= p is an argument error on both the one-shot and the split upload path (checked on the device)."""
    s = zkb.SynthCircuit(zk=False, seed=4, **zkb.TINY)
    c = zkb.ProverCircuit(s.common, s.const_sigma_values, is_values=True)
    bad = s.wires.copy()
    bad[3, 5] = np.uint64(0xFFFFFFFF00000001)          # = p
    for call in (lambda: c.prove(bad, s.public_inputs), lambda: c.upload_witness(bad)):
        with pytest.raises(zkb.ZkbError) as e:
            call()
        assert e.value.status == "ZKB_E_ARG"
    assert len(c.prove(s.wires, s.public_inputs)) == c.proof_size      # the context is still usable


@pytest.mark.gpu
def test_c_client_proof_is_accepted_and_byte_identical(zkb, oracle, tmp_path):
    """examples/prove_example.c (plain C over the ABI, no Python in the loop) proves the same synthetic circuit; its bytes
    are accepted by the pinned verifier and equal the oracle prover's."""
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe, out = tmp_path / "prove_example", tmp_path / "proof.bin"
    lib_dir = os.path.dirname(zkb.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c11", "-O2", "-I", os.path.join(root, "include"), os.path.join(root, "examples", "prove_example.c"),
                           "-L", lib_dir, "-lzkb200", "-lzkb200_synth", f"-Wl,-rpath,{lib_dir}", "-o", str(exe)])
    subprocess.check_call([str(exe), "1", str(out)])
    proof = out.read_bytes()
    s = zkb.SynthCircuit(zk=True, seed=42, **zkb.TINY)
    oc = oracle.Circuit(s.common, s.const_sigma_values)
    assert oc.verify(proof) == ""
    assert proof == oc.prove(s.wires, s.public_inputs, salt_seed=7)


@pytest.mark.parametrize("zk,min_degree_bits", [(False, 15), (True, 16)])
def test_proof_bytes_above_the_single_block_transform_size(zkb, oracle, zk, min_degree_bits):
    """n = 2^15 / 2^16 (larger than any circuit of the reference, which stop at 2^14): every transform of the proof takes the
    multi-step path — plain and coset LDEs with the in-register first step, in-place inverse transforms with the tiled bit
    reversal, the quotient's coset iNTT at 8n and the FRI extension-field LDE — and the bytes still equal the oracle's."""
    s, oc, gc, proof = run_case(zkb, oracle, oracle.Synth.VOTING, zk, seed=11, min_degree_bits=min_degree_bits)
    assert s.info["degree_bits"] == min_degree_bits

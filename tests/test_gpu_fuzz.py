"""Seeded random circuit shapes through the whole prover: row mixes, public-input counts, degrees and both gate sets drawn at
random; every proof must be byte-identical to the oracle's and accepted by the pinned verifier. Complements the fixed shapes of
test_gpu_prove.py (a shape-dependent slip — a tile boundary, a slice count, a selector group — shows up here)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def zkb():
    import zkb200

    if zkb200.device_count() == 0:
        pytest.fail("no CUDA device: -m gpu tests must run on the B200 box")
    return zkb200


def random_spec(rng, recursion):
    spec = dict(n_poseidon=int(rng.integers(1, 40)), n_base_sum=int(rng.integers(0, 30)), n_arith=int(rng.integers(0, 40)),
                n_const=int(rng.integers(1, 8)), num_public_inputs=int(rng.integers(0, 20)))
    if recursion:
        keys = ("n_arith_ext", "n_mul_ext", "n_reducing", "n_reducing_ext", "n_random_access", "n_exp", "n_coset", "n_mds")
        counts = rng.integers(0, 12, size=len(keys))
        counts[int(rng.integers(0, len(keys)))] += 1           # at least one recursion row selects the 14-gate set
        spec.update({k: int(c) for k, c in zip(keys, counts)})
    return spec


@pytest.mark.parametrize("case", range(10))
def test_random_shapes_prove_byte_identically(zkb, oracle, case):
    rng = np.random.default_rng(1000 + case)
    recursion = case % 2 == 1
    spec = random_spec(rng, recursion)
    min_bits = (0, 0, 9, 10, 11, 12, 5, 13, 0, 8)[case]          # padding rows: degrees from the natural one up to 2^13
    zk = case == 8                                               # one zero-knowledge case (n = 2^14 with the blinding rows)
    s = oracle.Synth(zk=zk, seed=500 + case, min_degree_bits=min_bits, **spec)
    assert s.check() == ""
    oc = oracle.Circuit(s.common, s.const_sigma_values)
    gc = zkb.ProverCircuit(s.common, s.const_sigma_values, is_values=True, circuit_digest=oc.digest)
    want = oc.prove(s.wires, s.public_inputs, salt_seed=case)
    got = gc.prove(s.wires, s.public_inputs, salt_seed=case)
    assert oc.verify(got) == "", f"rejected: spec {spec}, degree_bits {s.info['degree_bits']}"
    assert got == want, f"bytes differ: spec {spec}, degree_bits {s.info['degree_bits']}"

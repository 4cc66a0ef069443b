"""Parity of the CUDA stages (through the C ABI) against the CPU oracle on seeded inputs — bit exact.
Covers SURVEY.md §8 rows a1-a8: field/Poseidon, Merkle, LDE-NTT, partial products, quotient."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
P = 0xFFFFFFFF00000001


@pytest.fixture(scope="module")
def zkb():
    import zkb200

    if zkb200.device_count() == 0:
        pytest.fail("no CUDA device: -m gpu tests must run on the B200 box")
    return zkb200


def rand_felts(rng, shape):
    v = rng.integers(0, P, size=shape, dtype=np.uint64)
    return v


def test_poseidon_permutation_matches_oracle(zkb, oracle):
    rng = np.random.default_rng(1)
    states = rand_felts(rng, (4099, 12))
    states[0] = np.arange(12)
    states[1] = 0
    states[2] = P - 1
    states[3, ::2] = P - 1
    states[4] = 0xFFFFFFFF
    states[5] = 0xFFFFFFFF00000000
    got = zkb.poseidon_permute_batch(states)
    assert np.array_equal(got, oracle.poseidon_permute(states))
    assert [int(x) for x in got[0][:4]] == [0xD64E1E3EFC5B8E9E, 0x53666633020AAA47, 0xD40285597C6A8825, 0x613A4F81E81231D2]
    assert zkb.poseidon_permute_batch(np.zeros((0, 12), dtype=np.uint64)).shape == (0, 12)


def test_noncanonical_input_rejected(zkb):
    bad = np.zeros((1, 12), dtype=np.uint64)
    bad[0, 3] = P
    with pytest.raises(zkb.ZkbError) as e:
        zkb.poseidon_permute_batch(bad)
    assert e.value.status == "ZKB_E_ARG"


@pytest.mark.parametrize("width,lg,cap_h", [(1, 5, 0), (4, 6, 2), (5, 6, 6), (8, 7, 4), (9, 8, 4), (16, 4, 4),
                                            (24, 10, 4), (135, 11, 4), (139, 9, 4), (84, 12, 0), (135, 13, 4), (20, 14, 3), (7, 13, 13)])
def test_merkle_commit_matches_oracle(zkb, oracle, width, lg, cap_h):
    rng = np.random.default_rng(width * 100 + lg)
    leaves = rand_felts(rng, (width, 1 << lg))
    digests, cap = zkb.merkle_commit(leaves, cap_h)
    odig, ocap = oracle.merkle_commit(leaves, cap_h)
    assert np.array_equal(cap, ocap)
    assert np.array_equal(digests, odig)


@pytest.mark.parametrize("lg_n,ncols,rate_bits", [(1, 3, 3), (2, 1, 3), (5, 7, 3), (8, 4, 0), (10, 5, 1), (12, 3, 3),
                                                  (13, 2, 3), (14, 3, 3), (15, 2, 3), (16, 1, 2), (17, 2, 3), (18, 1, 1), (19, 1, 3)])
def test_lde_matches_oracle(zkb, oracle, lg_n, ncols, rate_bits):
    rng = np.random.default_rng(lg_n * 10 + ncols)
    vals = rand_felts(rng, (ncols, 1 << lg_n))
    coeffs, lde = zkb.lde_batch(vals, rate_bits)
    oc, ol = oracle.lde_batch(vals, rate_bits)
    assert np.array_equal(coeffs, oc)
    assert np.array_equal(lde, ol)
    # from_coeffs path (quotient chunks)
    c2, l2 = zkb.lde_batch(oc, rate_bits, from_coeffs=True)
    assert np.array_equal(c2, oc) and np.array_equal(l2, ol)


def test_lde_edge_inputs(zkb, oracle):
    # all-zero, all p-1 and a delta column; empty batch is a no-op
    n = 1 << 9
    vals = np.zeros((3, n), dtype=np.uint64)
    vals[1] = P - 1
    vals[2, 7] = 1
    coeffs, lde = zkb.lde_batch(vals, 3)
    oc, ol = oracle.lde_batch(vals, 3)
    assert np.array_equal(coeffs, oc) and np.array_equal(lde, ol)
    assert not lde[0].any()
    zkb.lde_batch(np.zeros((0, 8), dtype=np.uint64), 3)
    with pytest.raises(zkb.ZkbError):
        zkb.lde_batch(np.zeros((1, 12), dtype=np.uint64), 3)     # not a power of two


def test_lde_large_properties(zkb, oracle):
    """Full-size microbench shape (BASELINE config #3, n = 2^20): size-independent properties —
    linearity of the transform and direct evaluation of a few leaves."""
    rng = np.random.default_rng(5)
    lg_n, rb = 20, 3
    n = 1 << lg_n
    a = rand_felts(rng, (1, n))
    b = rand_felts(rng, (1, n))
    s = ((a.astype(object) + b.astype(object)) % P).astype(np.uint64)
    ca, la = zkb.lde_batch(np.concatenate([a, b, s]), rb)
    lsum = ((la[0].astype(object) + la[1].astype(object)) % P).astype(np.uint64)
    assert np.array_equal(lsum, la[2])
    # spot-check: lde[c][l] = P_c(g * w^bitrev(l)) evaluated from the coefficients in Python
    lgN = lg_n + rb
    g, w = 0xC65C18B67785D900, oracle.root_of_unity(lgN)
    coeffs = [int(x) for x in ca[0]]
    for l in (0, 1, 12345, (1 << lgN) - 1):
        i = int(format(l, f"0{lgN}b")[::-1], 2)
        x = g * pow(w, i, P) % P
        acc = 0
        for k in reversed(range(n)):
            acc = (acc * x + coeffs[k]) % P
        assert acc == int(la[0, l])
    # coefficients interpolate the values: evaluate at w_n^j for a few j
    wn = oracle.root_of_unity(lg_n)
    for j in (0, 3, n - 1):
        x = pow(wn, j, P)
        acc = 0
        for k in reversed(range(n)):
            acc = (acc * x + coeffs[k]) % P
        assert acc == int(a[0, j])


def test_commit_batch_matches_separate_stages(zkb, oracle):
    rng = np.random.default_rng(9)
    vals = rand_felts(rng, (20, 1 << 10))
    cap, times = zkb.commit_batch(vals, 3, 4)
    _, lde = oracle.lde_batch(vals, 3)
    _, ocap = oracle.merkle_commit(lde, 4)
    assert np.array_equal(cap, ocap)
    assert times["lde_ms"] > 0 and times["merkle_ms"] > 0


@pytest.fixture(scope="module")
def tiny(oracle, zkb):
    s = oracle.Synth(zk=False, seed=3, **oracle.Synth.TINY)
    oc = oracle.Circuit(s.common, s.const_sigma_values)
    gc = zkb.ProverCircuit(s.common, s.const_sigma_values, is_values=True)
    return s, oc, gc


def test_circuit_context_matches_oracle(zkb, oracle, tiny):
    s, oc, gc = tiny
    cap, digest = gc.verifier_only()
    assert np.array_equal(cap, oc.cap) and np.array_equal(digest, oc.digest)
    # coefficient-form input (prover_only.constants_sigmas_commitment.polynomials) gives the same context
    gc2 = zkb.ProverCircuit(s.common, oc.const_sigma_coeffs(), is_values=False, circuit_digest=oc.digest)
    cap2, digest2 = gc2.verifier_only()
    assert np.array_equal(cap2, oc.cap) and np.array_equal(digest2, oc.digest)
    with pytest.raises(zkb.ZkbError) as e:
        zkb.ProverCircuit(s.common, s.const_sigma_values, is_values=True, circuit_digest=np.array([1, 2, 3, 4], dtype=np.uint64))
    assert e.value.status == "ZKB_E_DIGEST"


def test_partial_products_and_quotient_match_oracle(zkb, oracle, tiny):
    s, oc, gc = tiny
    _, tr = oc.prove(s.wires, s.public_inputs, trace=True)
    zs = gc.partial_products(s.wires, tr.betas, tr.gammas, 20, s.n)
    assert np.array_equal(zs, tr.zs_pp_values)
    q = gc.quotient(s.wires, zs, s.public_inputs, tr.betas, tr.gammas, tr.alphas, 16, s.n)
    assert np.array_equal(q, tr.quotient_chunks)


def test_recursion_gate_quotient_matches_oracle(zkb, oracle):
    """compute_quotient_polys over the recursion gate set: every quotient chunk bit-exact against the oracle."""
    s = oracle.Synth(seed=9, n_poseidon=20, n_base_sum=6, n_arith=10, n_const=4, num_public_inputs=7, n_arith_ext=20, n_mul_ext=8,
                     n_reducing=6, n_reducing_ext=6, n_random_access=9, n_exp=5, n_coset=7, n_mds=3)
    oc = oracle.Circuit(s.common, s.const_sigma_values)
    gc = zkb.ProverCircuit(s.common, s.const_sigma_values, is_values=True, circuit_digest=oc.digest)
    _, tr = oc.prove(s.wires, s.public_inputs, trace=True)
    zs = gc.partial_products(s.wires, tr.betas, tr.gammas, 20, s.n)
    assert np.array_equal(zs, tr.zs_pp_values)
    q = gc.quotient(s.wires, zs, s.public_inputs, tr.betas, tr.gammas, tr.alphas, 16, s.n)
    assert np.array_equal(q, tr.quotient_chunks)


@pytest.mark.parametrize("recursion,min_bits", [(False, 0), (True, 0), (True, 14)])
def test_quotient_on_edge_values(zkb, oracle, recursion, min_bits):
    """Carry / borrow corners of the device field arithmetic through every gate evaluator: wires, Z / partial-product
    columns, public inputs and challenges drawn from {0, 1, 2, p-1, p-2, 2^32-1, 2^32, 2^32+1, 2^63, p - 2^32, ...}. The
    witness does not satisfy the circuit — compute_quotient_polys is a pointwise computation on the coset followed by an
    inverse transform, defined for any input — and every chunk must still equal the oracle's bit for bit."""
    P = oracle.P
    edge = np.array([0, 1, 2, 7, P - 1, P - 2, P - 7, 2**32 - 1, 2**32, 2**32 + 1, 2**63, 2**63 - 1, P - 2**32, P - 2**32 + 1,
                     2**64 - 2**33, 0xFFFFFFFE00000002, 0x00000001FFFFFFFF], dtype=np.uint64)
    spec = oracle.Synth.RECURSION_TINY if recursion else oracle.Synth.TINY
    s = oracle.Synth(seed=21, min_degree_bits=min_bits, **spec)     # 2^14: the single-stream (unsliced) launch sequence
    oc = oracle.Circuit(s.common, s.const_sigma_values)
    gc = zkb.ProverCircuit(s.common, s.const_sigma_values, is_values=True)
    rng = np.random.default_rng(5)
    for trial in range(3 if min_bits == 0 else 1):
        wires = edge[rng.integers(0, len(edge), size=s.wires.shape)]
        zs = edge[rng.integers(0, len(edge), size=(20, s.n))]
        pis = edge[rng.integers(0, len(edge), size=s.public_inputs.shape)]
        ch = [int(x) for x in edge[rng.integers(0, len(edge), size=6)]]
        if trial == 0:
            ch = [P - 1, 2**32 - 1, 1, P - 2**32, 2**32, P - 2]
        want = oc.quotient(wires, zs, pis, ch[0:2], ch[2:4], ch[4:6])
        got = gc.quotient(wires, zs, pis, ch[0:2], ch[2:4], ch[4:6], 16, s.n)
        assert np.array_equal(got, want)


@pytest.mark.parametrize("lg_n,ncols", [(9, 7), (14, 3), (16, 2)])
def test_coset_sharded_commit_parts_concatenate_to_the_full_cap(zkb, oracle, lg_n, ncols):
    """zkb_commit_cosets (one GPU's share of a coset-sharded commitment) for G = 1, 2, 4, 8 emulated ranks on one GPU:
    the parts, in rank order, are the cap of the unsharded commitment (which is bit-exact with the oracle)."""
    rng = np.random.default_rng(lg_n)
    vals = rand_felts(rng, (ncols, 1 << lg_n))
    want, _ = zkb.commit_batch(vals, 3, 4)
    _, lde = oracle.lde_batch(vals, 3)
    _, ocap = oracle.merkle_commit(lde, 4)
    assert np.array_equal(want, ocap)
    for G in (1, 2, 4, 8):
        per = 8 // G
        parts = [zkb.commit_cosets(vals, 3, 4, r * per, (r + 1) * per)[0] for r in range(G)]
        assert np.array_equal(np.concatenate(parts), want)
    with pytest.raises(zkb.ZkbError):
        zkb.commit_cosets(vals, 3, 4, 1, 3)          # unaligned block range
    with pytest.raises(zkb.ZkbError):
        zkb.commit_cosets(vals, 3, 2, 0, 4)          # cap_height < rate_bits


@pytest.mark.parametrize("lg_n", [19, 20, 21])
def test_coset_sharded_commit_parts_at_the_staged_and_three_step_sizes(zkb, lg_n):
    """The same identity where the first transform step runs in the staged kernels (n1 = 32, 64) and in three steps (2^21),
    with a non-zero first coset and fewer than 8 blocks per call; the unsharded cap at 2^20 is pinned by the golden test."""
    rng = np.random.default_rng(lg_n)
    vals = rand_felts(rng, (1, 1 << lg_n))
    want, _ = zkb.commit_batch(vals, 3, 4)
    for G in (2, 8):
        per = 8 // G
        parts = [zkb.commit_cosets(vals, 3, 4, r * per, (r + 1) * per)[0] for r in range(G)]
        assert np.array_equal(np.concatenate(parts), want)


def test_lde_max_microbench_size_properties(zkb, oracle):
    """BASELINE config #3's largest degree, n = 2^22 (three-step transform, 4 x 64 x 2^14): linearity of the
    whole from_values map, and one leaf + one subgroup value re-evaluated from the returned coefficients by Horner."""
    rng = np.random.default_rng(22)
    lg_n, rb = 22, 3
    n = 1 << lg_n
    a = rand_felts(rng, (1, n))
    b = rand_felts(rng, (1, n))
    s = ((a.astype(object) + b.astype(object)) % P).astype(np.uint64)
    ca, la = zkb.lde_batch(np.concatenate([a, b, s]), rb)
    assert np.array_equal(((la[0].astype(object) + la[1].astype(object)) % P).astype(np.uint64), la[2])
    assert np.array_equal(((ca[0].astype(object) + ca[1].astype(object)) % P).astype(np.uint64), ca[2])
    lgN = lg_n + rb
    g, w = 0xC65C18B67785D900, oracle.root_of_unity(lgN)
    coeffs = [int(x) for x in ca[0]]

    def horner(x):
        acc = 0
        for c in reversed(coeffs):
            acc = (acc * x + c) % P
        return acc

    l = 0x1234567
    i = int(format(l, f"0{lgN}b")[::-1], 2)
    assert horner(g * pow(w, i, P) % P) == int(la[0, l])
    j = 3141592
    assert horner(pow(oracle.root_of_unity(lg_n), j, P)) == int(a[0, j])


def _config3_columns(lg_n, cols):
    nn = 1 << lg_n
    idx = np.arange(nn * cols, dtype=np.uint64) + np.uint64(0xB200000000000001)
    with np.errstate(over="ignore"):
        z = idx
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return np.where(z >= np.uint64(P), z - np.uint64(P), z).reshape(cols, nn)


@pytest.mark.parametrize("lg_n,cols", [(14, 135), (14, 400), (16, 135), (16, 200), (18, 100), (20, 100), (19, 5), (21, 5), (22, 5)])
def test_commit_cap_matches_the_oracle_golden_at_microbench_sizes(zkb, lg_n, cols):
    """Exact parity at BASELINE.json config #3's sizes: the 16 cap digests of the fused from_values commitment (iNTT + coset
    LDE, single-block, two-step and — above 2^20 — three-step transforms, fused Merkle trees up to 2^25 leaves) equal the caps the
    CPU oracle computed once for the same seeded columns (tests/golden/make_caps.py -> config3_caps.json); the narrow shapes
    pin the staged 32-point first step (2^19) and the largest microbenchmark degree (2^22) exactly."""
    import json
    import os

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config3_caps.json")) as f:
        want = np.array([[int(x, 16) for x in d] for d in json.load(f)[f"{lg_n}x{cols}"]["cap"]], dtype=np.uint64)
    cap, _ = zkb.commit_batch(_config3_columns(lg_n, cols), 3, 4)
    assert np.array_equal(cap, want)

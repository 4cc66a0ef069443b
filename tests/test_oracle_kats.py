"""Oracle vs the reference's hash known-answer tests — CPU only."""
import hashlib

import numpy as np

P = 0xFFFFFFFF00000001


def inj(b):
    """common/src/utils.rs:162-174 injective_bytes_to_felts: 4-byte LE chunks."""
    out = []
    for i in range(0, len(b), 4):
        c = b[i:i + 4]
        out.append(int.from_bytes(c + b"\0" * (4 - len(c)), "little"))
    return out


def digest_bytes(d):
    """common/src/utils.rs:203-215 digest_felts_to_bytes."""
    return b"".join(int(x).to_bytes(8, "little") for x in d)


def test_round_constants_checksum(oracle):
    rc = oracle.round_constants()
    assert int(rc[0]) == 0xB585F766F2144405 and int(rc[359]) == 0xBC8DFB627FE558FC
    assert hashlib.sha256(rc.tobytes()).hexdigest() == "d2fcbb5be293c50ab4b1ddcd9c81005b12d689816a54c91a054f97f6588a20a8"
    assert int(rc.max()) < P


def test_permutation_kats(oracle):
    out = oracle.poseidon_permute(np.array([list(range(12)), [0] * 12], dtype=np.uint64))
    assert [int(x) for x in out[0][:4]] == [0xD64E1E3EFC5B8E9E, 0x53666633020AAA47, 0xD40285597C6A8825, 0x613A4F81E81231D2]
    assert [int(x) for x in out[1][:4]] == [0x3C18A9786CB0B359, 0xC4055E3364A246C3, 0x7953DB0AB48808F4, 0xC71603F33A1144CA]
    assert [int(x) for x in oracle.hash_pad([])] == [0xF9AD7EFEEE338AC6, 0x70014F06AE45AC42, 0x393D1B035A725D35, 0x2A6CE778AA4FB823]


def test_unspendable_account_kats(oracle, kats):
    # wormhole/tests/src/circuit/unspendable_account_tests.rs:12-27,47-61
    for secret, address in zip(kats["secrets"], kats["addresses"]):
        pre = inj(kats["unspendable_salt"].encode()) + inj(bytes.fromhex(secret))
        assert digest_bytes(oracle.hash_no_pad(oracle.hash_no_pad(pre))).hex() == address
    pre = inj(b"wormhole") + inj(bytes.fromhex(kats["default_secret"]))
    assert list(digest_bytes(oracle.hash_no_pad(oracle.hash_no_pad(pre)))) == kats["default_to_account"]


def test_nullifier_kat(oracle, kats):
    # wormhole/circuit/src/nullifier.rs:53-73; wormhole/tests/src/prover/prover_tests.rs:31-35
    cnt = kats["default_transfer_count"]
    pre = inj(kats["nullifier_salt"].encode()) + inj(bytes.fromhex(kats["default_secret"])) + [cnt >> 32, cnt & 0xFFFFFFFF]
    assert list(digest_bytes(oracle.hash_no_pad(oracle.hash_no_pad(pre)))) == kats["nullifier"]
    assert bytes(kats["root_hash_bytes"]).hex() == kats["default_root_hash"]


def test_storage_proof_chain(oracle, kats):
    # wormhole/tests/test-helpers/src/lib.rs:68-80; node padding circuit/src/storage_proof/mod.rs:23,279
    nodes, idx = kats["storage_proof"], kats["storage_proof_indices"]
    for i, nd in enumerate(nodes):
        f = inj(bytes.fromhex(nd))
        assert len(f) <= 188
        h = digest_bytes(oracle.hash_no_pad(f + [0] * (188 - len(f)))).hex()
        if i == 0:
            assert h == kats["default_root_hash"]
        else:
            assert nodes[i - 1][idx[i - 1]: idx[i - 1] + 64] == h


def test_sponge_edge_cases(oracle):
    # hash_or_noop semantics are exercised through merkle_commit with narrow leaves (A.2/A.3)
    leaves = np.arange(3 * 32, dtype=np.uint64).reshape(3, 32)
    digests, cap = oracle.merkle_commit(leaves, 5)
    assert digests.shape == (32, 4)
    assert np.array_equal(digests[:, :3], leaves.T) and not digests[:, 3].any()
    assert np.array_equal(cap, digests)
    # two_to_one == permutation of (l ‖ r ‖ 0000)
    l, r = np.arange(4, dtype=np.uint64), np.arange(4, 8, dtype=np.uint64)
    st = oracle.poseidon_permute(np.concatenate([l, r, np.zeros(4, dtype=np.uint64)])[None, :])
    assert np.array_equal(oracle.two_to_one(l, r), st[0][:4])
    # short final chunk leaves the rest of the rate untouched (overwrite mode)
    v = list(range(1, 12))
    st = oracle.poseidon_permute(np.array([v[:8] + [0] * 4], dtype=np.uint64))[0]
    st[:3] = v[8:]
    st = oracle.poseidon_permute(st[None, :])[0]
    assert np.array_equal(oracle.hash_no_pad(v), st[:4])


def test_fast_poseidon_equals_the_definition(oracle):
    """oracle/poseidon_fast.hpp (the CPU-baseline arm's AVX2 permutation) against the naive round form on random, corner and
    known-answer states; hash_no_pad through it reproduces the reference KATs."""
    import numpy as np

    P = oracle.P
    rng = np.random.default_rng(7)
    st = rng.integers(0, P, size=(5000, 12), dtype=np.uint64)
    st[0] = np.arange(12); st[1] = 0; st[2] = P - 1; st[3] = 0xFFFFFFFF; st[4] = 0xFFFFFFFF00000000; st[5, ::2] = P - 1
    try:
        oracle.set_fast(False)
        want = oracle.poseidon_permute(st)
        oracle.set_fast(True)
        got = oracle.poseidon_permute(st)
        assert np.array_equal(got, want)
        assert [int(x) for x in got[0][:4]] == [0xD64E1E3EFC5B8E9E, 0x53666633020AAA47, 0xD40285597C6A8825, 0x613A4F81E81231D2]
        assert [int(x) for x in oracle.hash_pad([])] == [0xF9AD7EFEEE338AC6, 0x70014F06AE45AC42, 0x393D1B035A725D35, 0x2A6CE778AA4FB823]
    finally:
        oracle.set_fast(False)

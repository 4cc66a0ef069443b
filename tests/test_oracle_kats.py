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


def test_vector_field_ops_of_the_fast_gate_evaluator(oracle):
    """oracle/vec_ops.hpp (the 4-lane AVX2 Ops the CPU-baseline arm instantiates the gate code over): add, sub, mul, mulc and
    the MDS layer against Python big-integer arithmetic on corner values (0, 1, p - 1, 2^32 +- 1, 2^64 - 2^32, sums that wrap
    64 bits, products that reduce to exactly p - 1 or 0) and random ones."""
    import ctypes

    import numpy as np

    P = oracle.P
    L = oracle.lib()
    u64p = ctypes.POINTER(ctypes.c_uint64)
    corners = [0, 1, 2, P - 1, P - 2, 0xFFFFFFFF, 0x100000000, 0x100000001, 0xFFFFFFFF00000000, 0xFFFFFFFEFFFFFFFF,
               0x8000000000000000, 0x7FFFFFFF80000001, (P + 1) // 2, (P - 1) // 2, pow(7, (P - 1) // 2 - 1, P), 0xFFFFFFFE00000002]
    rng = np.random.default_rng(11)
    a = np.array([x for x in corners for _ in corners] + [int(v) for v in rng.integers(0, P, size=4096, dtype=np.uint64)], dtype=np.uint64)
    b = np.array([y for _ in corners for y in corners] + [int(v) for v in rng.integers(0, P, size=4096, dtype=np.uint64)], dtype=np.uint64)
    n = len(a)
    assert n % 4 == 0
    outs = [np.zeros(n, dtype=np.uint64) for _ in range(4)]
    for c in (0, 1, 7, 41, P - 1, 0xFFFFFFFF00000000):
        rc = L.orc_vecops_check(a.ctypes.data_as(u64p), b.ctypes.data_as(u64p), ctypes.c_size_t(n), ctypes.c_uint64(c),
                                *[o.ctypes.data_as(u64p) for o in outs])
        assert rc == 0, oracle.last_error() if hasattr(oracle, "last_error") else rc
        ai, bi = [int(x) for x in a], [int(x) for x in b]
        assert [int(x) for x in outs[0]] == [(x + y) % P for x, y in zip(ai, bi)]
        assert [int(x) for x in outs[1]] == [(x - y) % P for x, y in zip(ai, bi)]
        assert [int(x) for x in outs[2]] == [(x * y) % P for x, y in zip(ai, bi)]
        assert [int(x) for x in outs[3]] == [(x * c) % P for x in ai]
    # MDS layer, four states per call
    C = [17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20]
    states = rng.integers(0, P, size=(8, 4, 12), dtype=np.uint64)
    states[0, 0] = P - 1
    states[0, 1] = 0
    states[0, 2, ::2] = P - 1
    for blk in states:
        out = np.zeros((4, 12), dtype=np.uint64)
        assert L.orc_vecops_mds(np.ascontiguousarray(blk).ctypes.data_as(u64p), out.ctypes.data_as(u64p)) == 0
        for k in range(4):
            x = [int(v) for v in blk[k]]
            want = [(sum(x[(i + r) % 12] * C[i] for i in range(12)) + (8 * x[0] if r == 0 else 0)) % P for r in range(12)]
            assert [int(v) for v in out[k]] == want

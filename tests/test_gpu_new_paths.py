"""Round-2 paths through the C ABI: fixture replay (coefficient-form context + explicit salts), CSPRNG salts, the witness
self-check (ZKB_E_UNSAT) and the asynchronous proof engine."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def zkb():
    import zkb200

    if zkb200.device_count() == 0:
        pytest.fail("no CUDA device: -m gpu tests must run on the B200 box")
    return zkb200


@pytest.mark.parametrize("zk", [False, True])
def test_fixture_replay_is_byte_identical(zkb, oracle, tmp_path, zk):
    """A fixture directory (what rust/zkb200::Fixture::write_dir exports on a Rust host; here written from the oracle) fed to
    zkb_prove: coefficient-form constants/sigmas + digest check + explicit salts -> the CPU proof's bytes. This is also the
    only path the Rust shim uses (is_values = 0): the sigma VALUES the partial products need are derived on the device."""
    from zkb200 import fixture

    s = oracle.Synth(zk=zk, seed=12, **oracle.Synth.TINY)
    oc = oracle.Circuit(s.common, s.const_sigma_values)
    salts = np.random.default_rng(7).integers(0, oracle.P, size=(3, 4, s.n * 8), dtype=np.uint64) if zk else None
    proof = oc.prove(s.wires, s.public_inputs, salts=salts)
    fixture.write(tmp_path / "fx", s.common, oc.const_sigma_coeffs(), oc.digest, s.wires, s.public_inputs, proof, salts=salts)
    got, want = fixture.replay(tmp_path / "fx")
    assert got == want == proof


def test_coefficient_form_context_proves_like_the_value_form(zkb, oracle):
    s = oracle.Synth(zk=True, seed=1, **oracle.Synth.WORMHOLE)
    oc = oracle.Circuit(s.common, s.const_sigma_values)
    a = zkb.ProverCircuit(s.common, s.const_sigma_values, is_values=True)
    b = zkb.ProverCircuit(s.common, oc.const_sigma_coeffs(), is_values=False, circuit_digest=oc.digest)
    pa = a.prove(s.wires, s.public_inputs, salt_seed=3, check_witness=True)
    pb = b.prove(s.wires, s.public_inputs, salt_seed=3, check_witness=True)
    assert pa == pb == oc.prove(s.wires, s.public_inputs, salt_seed=3)


def test_default_salts_come_from_a_csprng(zkb, oracle):
    """No salts and no seed: the device draws them from ChaCha20 keyed by the OS per proof — two proofs of the same witness
    differ in every salted leaf, both verify, and neither equals the seeded test stream."""
    s = oracle.Synth(zk=True, seed=8, **oracle.Synth.TINY)
    oc = oracle.Circuit(s.common, s.const_sigma_values)
    gc = zkb.ProverCircuit(s.common, s.const_sigma_values, is_values=True)
    p1, p2 = gc.prove(s.wires, s.public_inputs), gc.prove(s.wires, s.public_inputs)
    assert oc.verify(p1) == "" and oc.verify(p2) == ""
    assert p1 != p2 and p1[:512] != p2[:512]                       # different wires caps
    assert p1 != gc.prove(s.wires, s.public_inputs, salt_seed=0)
    # the first opened wires leaf: 135 values + 4 salts; the salts must look uniform (not small, not equal)
    info = s.info
    start = 3 * 512 + 16 * (4 + 80 + 135 + 2 + 2 + 18 + 16) + 512 * len(info["reduction_arity_bits"])
    ncs = info["num_constants"] + 80
    off = start + 8 * ncs + 1 + 32 * (info["degree_bits"] + 3 - 4)
    salts1 = np.frombuffer(p1[off + 8 * 135: off + 8 * 139], dtype="<u8")
    assert len(set(int(x) for x in salts1)) == 4 and all(int(x) >> 40 for x in salts1)


def test_non_canonical_salts_are_rejected(zkb, oracle):
    s = oracle.Synth(zk=True, seed=8, **oracle.Synth.TINY)
    gc = zkb.ProverCircuit(s.common, s.const_sigma_values, is_values=True)
    salts = np.zeros((3, 4, s.n * 8), dtype=np.uint64)
    salts[1, 2, 17] = oracle.P
    with pytest.raises(zkb.ZkbError) as e:
        gc.prove(s.wires, s.public_inputs, salts=salts)
    assert e.value.status == "ZKB_E_ARG"
    assert len(gc.prove(s.wires, s.public_inputs)) == gc.proof_size          # context still usable


@pytest.mark.parametrize("shape", ["tiny", "tiny_zk", "recursion"])
def test_witness_check_returns_unsat_instead_of_an_unverifiable_proof(zkb, oracle, shape):
    """ZKB_CHECK_WITNESS (reference behaviour: bad inputs are an Err, voting/src/lib.rs:399-403): a satisfied witness proves
    to the same bytes as without the flag; flipping one wire of a gate row, or breaking a copy constraint, gives ZKB_E_UNSAT —
    and the context proves the good witness again afterwards."""
    spec = oracle.Synth.RECURSION_TINY if shape == "recursion" else oracle.Synth.TINY
    s = oracle.Synth(zk=shape == "tiny_zk", seed=3, **spec)
    oc = oracle.Circuit(s.common, s.const_sigma_values)
    gc = zkb.ProverCircuit(s.common, s.const_sigma_values, is_values=True)
    good = gc.prove(s.wires, s.public_inputs, salt_seed=4, check_witness=True)
    assert good == oc.prove(s.wires, s.public_inputs, salt_seed=4)
    sel = s.const_sigma_values[0]
    rows = {int(v): int(np.nonzero(sel == v)[0][0]) for v in np.unique(sel) if int(v) < 0xFFFFFFFF}
    cases = []
    for gate_index, row in rows.items():
        w = s.wires.copy()
        w[0, row] ^= np.uint64(1)
        if oc.verify(oc.prove(w, s.public_inputs, salt_seed=4)) != "":        # the flip really breaks something
            cases.append(w)
    assert len(cases) >= 3
    for w in cases:
        with pytest.raises(zkb.ZkbError) as e:
            gc.prove(w, s.public_inputs, salt_seed=4, check_witness=True)
        assert e.value.status == "ZKB_E_UNSAT"
        assert oc.verify(gc.prove(w, s.public_inputs, salt_seed=4)) != ""     # without the flag: bytes, but unverifiable
    assert gc.prove(s.wires, s.public_inputs, salt_seed=4, check_witness=True) == good


def test_engine_proofs_are_byte_identical_and_overlap(zkb, oracle):
    """zkb_engine: 3 contexts, 5 slots, 9 proofs submitted from two caller threads that fill the pinned slots themselves; every
    proof equals the oracle prover's for its seed; a bad witness fails only its own ticket."""
    s = oracle.Synth(zk=True, seed=21, **oracle.Synth.TINY)
    oc = oracle.Circuit(s.common, s.const_sigma_values)
    eng = zkb.Engine(s.common, s.const_sigma_values, is_values=True, contexts=3, slots=5)
    assert eng.proof_size == len(oc.prove(s.wires, s.public_inputs, salt_seed=0))
    results = {}

    def client(seeds):
        pending = []
        for seed in seeds:
            slot, buf = eng.acquire()
            buf[:] = s.wires
            if seed == 104:
                buf[3, 2] ^= np.uint64(1)                 # unsatisfied
            eng.submit(slot, s.public_inputs, salt_seed=seed, check_witness=True)
            pending.append((seed, slot))
            if len(pending) == 2:
                sd, sl = pending.pop(0)
                results[sd] = _wait(eng, sl)
        for sd, sl in pending:
            results[sd] = _wait(eng, sl)

    def _wait(e, slot):
        try:
            return e.wait(slot).tobytes()
        except zkb.ZkbError as err:
            return err.status

    th = [threading.Thread(target=client, args=(range(100, 105),)), threading.Thread(target=client, args=(range(200, 204),))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert len(results) == 9
    for seed, proof in results.items():
        if seed == 104:
            assert proof == "ZKB_E_UNSAT"
        else:
            assert proof == oc.prove(s.wires, s.public_inputs, salt_seed=seed)
    # resident re-prove (benchmark arm) and argument errors at submit
    slot, buf = eng.acquire()
    buf[:] = s.wires
    eng.submit(slot, s.public_inputs, salt_seed=5)
    first = eng.wait(slot).tobytes()
    slot, _ = eng.acquire()
    with pytest.raises(zkb.ZkbError) as e:
        eng.submit(slot, s.public_inputs[:-1], salt_seed=5)
    assert e.value.status == "ZKB_E_ARG"
    eng.release(slot)
    assert first == oc.prove(s.wires, s.public_inputs, salt_seed=5)
    eng.close()


def _chunks_to_coset_values(oracle, chunks, n, rate_bits=3):
    """LDE (leaf order) of t(X) = sum_m X^(m n) t_m(X) from its chunks: on coset j every x has x^n = c_j = (g w_N^j)^n, so
    t = sum_m c_j^m t_m there — built from the oracle's LDE of the chunks with plain integer arithmetic."""
    R = 1 << rate_bits
    _, lde = oracle.lde_batch(chunks, rate_bits, from_coeffs=True)          # [R][R n], leaf order
    g, w_N = 0xC65C18B67785D900, oracle.root_of_unity(n.bit_length() - 1 + rate_bits)
    out = np.zeros(R * n, dtype=np.uint64)
    for jb in range(R):
        j = int(format(jb, f"0{rate_bits}b")[::-1], 2)
        c = pow(g * pow(w_N, j, oracle.P) % oracle.P, n, oracle.P)
        acc = [0] * n
        for m in range(R):
            cm = pow(c, m, oracle.P)
            blk = lde[m, jb * n:(jb + 1) * n]
            acc = [(a + cm * int(v)) % oracle.P for a, v in zip(acc, blk)]
        out[jb * n:(jb + 1) * n] = acc
    return out


def test_nccl_comm_single_rank_commit_and_quotient_chunks(zkb, oracle):
    """The NCCL-backed sharded entry points on a one-rank communicator (the collectives degenerate; the N > 1 exchange is run by
    bench.py --gpus N and, for the host logic, by tests/test_batch_gloo.py): commit cap = the plain commitment's cap, and the
    chunk recovery from coset-local evaluations inverts the LDE of X^(m n)-stacked chunks."""
    comm = zkb.Comm(zkb.comm_unique_id(), 1, 0)
    rng = np.random.default_rng(4)
    vals = rng.integers(0, oracle.P, size=(37, 1 << 12), dtype=np.uint64)
    cap, tm = comm.commit(vals, 3, 4, reps=2)
    want, _ = zkb.commit_batch(vals, 3, 4)
    assert np.array_equal(cap, want) and tm["lde_ms"] > 0
    n = 1 << 7
    chunks = rng.integers(0, oracle.P, size=(2, 8, n), dtype=np.uint64)
    q = np.stack([_chunks_to_coset_values(oracle, chunks[ch], n) for ch in range(2)])
    got, _ = comm.quotient_chunks(q, n, 3)
    assert np.array_equal(got, chunks)
    comm.close()

"""N > 1 host path on CPU: two gloo ranks shard a batch of independent proofs (replica mode, no data-path collective),
gather them on rank 0 and reduce timings with max-over-ranks. The per-rank prover here is the ORACLE (CPU) standing in
for a GPU prover context; on the GPU box the same zkb200.batch code drives ProverCircuit objects (tests/test_gpu_prove.py)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_covers_everything_once():
    sys.path.insert(0, os.path.join(ROOT, "zk-circuits_b200"))
    from zkb200.batch import shard_range

    for total in (0, 1, 7, 8, 9, 64):
        for ws in (1, 2, 3, 4, 8):
            seen = []
            sizes = []
            for r in range(ws):
                lo, hi = shard_range(total, r, ws)
                seen += list(range(lo, hi))
                sizes.append(hi - lo)
            assert seen == list(range(total))
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _rank_main(rank, world_size, port, tmpdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size))
    for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "zk-circuits_b200")):
        sys.path.insert(0, p)
    import torch.distributed as dist
    import oracle as O
    from zkb200 import batch

    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        s = O.Synth(zk=False, seed=3, **O.Synth.TINY)
        circ = O.Circuit(s.common, s.const_sigma_values)
        witnesses = list(range(5))               # 5 proofs over 2 ranks: 3 + 2; the "witness" picks the salt seed
        calls = []

        def prove_fn(prover, w, i):
            calls.append(i)
            return prover.prove(s.wires, s.public_inputs, salt_seed=100 + w)

        proofs = batch.prove_batch(witnesses, [circ, circ], prove_fn)
        lo, hi = batch.shard_range(len(witnesses), rank, world_size)
        assert sorted(calls) == list(range(lo, hi))
        t = batch.max_over_ranks([float(rank + 1), 10.0 - rank])
        assert t == [float(world_size), 10.0]
        if rank == 0:
            assert len(proofs) == len(witnesses)
            want = circ.prove(s.wires, s.public_inputs, salt_seed=100)
            assert all(p == want for p in proofs)            # non-zk: salts unused, every proof identical and in order
            assert circ.verify(proofs[-1]) == ""
            open(os.path.join(tmpdir, "ok"), "w").write("ok")
        else:
            assert proofs is None
    finally:
        dist.destroy_process_group()


def test_two_rank_batch_over_gloo(tmp_path):
    import torch.multiprocessing as mp

    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_rank_main, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").read_text() == "ok"


def _sharded_rank_main(rank, world_size, port, tmpdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size))
    for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "zk-circuits_b200")):
        sys.path.insert(0, p)
    import numpy as np
    import torch.distributed as dist
    import oracle as O
    from zkb200 import sharded

    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        rng = np.random.default_rng(11)                     # same values on every rank
        vals = rng.integers(0, O.P, size=(5, 64), dtype=np.uint64)
        rb, ch = 3, 4
        _, lde = O.lde_batch(vals, rb)
        _, want = O.merkle_commit(lde, ch)

        def cpu_commit(v, rate_bits, cap_height, lo, hi):   # oracle stand-in for zkb_commit_cosets: sub-range of the leaves
            n = v.shape[1]
            _, l = O.lde_batch(v, rate_bits)
            nb = hi - lo
            local = np.ascontiguousarray(l[:, lo * n:hi * n])
            _, part = O.merkle_commit(local, cap_height - rate_bits + nb.bit_length() - 1)
            return part, {}

        cap, _ = sharded.sharded_commit(vals, rb, ch, commit_fn=cpu_commit)
        assert cap.shape == want.shape and np.array_equal(cap, want)
        assert sharded.block_range(rb, rank, world_size) == (rank * 8 // world_size, (rank + 1) * 8 // world_size)
        if rank == 0:
            open(os.path.join(tmpdir, "ok2"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world_size", [2, 4])
def test_coset_sharded_commit_over_gloo(tmp_path, world_size):
    """Sharding + the single all-gather of cap digests, with the oracle computing each rank's block range on the CPU."""
    import torch.multiprocessing as mp

    port = 31500 + (os.getpid() % 2000) + world_size
    mp.spawn(_sharded_rank_main, args=(world_size, port, str(tmp_path)), nprocs=world_size, join=True)
    assert (tmp_path / "ok2").read_text() == "ok"


def _quotient_rank_main(rank, world_size, port, tmpdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size))
    for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "zk-circuits_b200")):
        sys.path.insert(0, p)
    import numpy as np
    import torch.distributed as dist
    import oracle as O
    from zkb200 import sharded

    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        n, rb, R = 64, 3, 8
        rng = np.random.default_rng(23)                     # same chunks on every rank
        chunks = rng.integers(0, O.P, size=(2, R, n), dtype=np.uint64)
        w_N = O.root_of_unity(6 + rb)
        B, sl = R // world_size, n // world_size
        # t on this rank's coset blocks: t = sum_m c_j^m t_m on coset j (c_j = (g w_N^j)^n)
        q = np.zeros((2, B * n), dtype=np.uint64)
        for ch in range(2):
            _, lde = O.lde_batch(chunks[ch], rb, from_coeffs=True)
            for i in range(B):
                jb = rank * B + i
                c = pow(sharded.GEN * pow(w_N, sharded._bitrev(jb, rb), O.P) % O.P, n, O.P)
                acc = [0] * n
                for m in range(R):
                    cm = pow(c, m, O.P)
                    acc = [(a + cm * int(v)) % O.P for a, v in zip(acc, lde[m, jb * n:(jb + 1) * n])]
                q[ch, i * n:(i + 1) * n] = acc

        def interpolate(vals_bitrev, shift):                # oracle stand-in for the device coset iNTT
            nat = np.array([vals_bitrev[sharded._bitrev(k, 6)] for k in range(n)], dtype=np.uint64)[None, :]
            c = O.ntt(nat, inverse=True)[0]
            sinv = pow(shift, O.P - 2, O.P)
            return np.array([int(c[k]) * pow(sinv, k, O.P) % O.P for k in range(n)], dtype=np.uint64)

        got = sharded.quotient_chunks(q, n, rb, interpolate, O.root_of_unity)
        assert got.shape == (2, R, sl)
        assert np.array_equal(got, chunks[:, :, rank * sl:(rank + 1) * sl])
        dist.barrier()
        if rank == 0:
            open(os.path.join(tmpdir, "ok3"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world_size", [2, 4])
def test_quotient_chunk_exchange_over_gloo(tmp_path, world_size):
    """the quotient is evaluated coset-locally; ONE all-to-all of coefficient slices + an 8 x 8 Vandermonde
    solve per index give every rank its slice of the 8 degree-n chunks. Host logic of zkb_quotient_chunks_sharded over gloo,
    with the oracle doing each rank's coset interpolation."""
    import torch.multiprocessing as mp

    port = 33500 + (os.getpid() % 2000) + world_size
    mp.spawn(_quotient_rank_main, args=(world_size, port, str(tmp_path)), nprocs=world_size, join=True)
    assert (tmp_path / "ok3").read_text() == "ok"


def test_aggregate_tree_mirrors_the_reference_level_order():
    """zkb200.batch.aggregate_tree = aggregate_to_tree (aggregator/src/circuits/tree.rs:55-77): chunks of `branching`,
    every chunk of a level before any chunk of the next, children in order, a short last chunk allowed."""
    sys.path.insert(0, os.path.join(ROOT, "zk-circuits_b200"))
    import threading
    from zkb200.batch import aggregate_tree

    lock, log = threading.Lock(), []

    def prove_chunk(prover, chunk, level, index):
        with lock:
            log.append((level, index, prover))
        return "(" + "".join(chunk) + ")"

    root, widths = aggregate_tree(list("abcdefgh"), 2, prove_chunk, ["s0", "s1", "s2"])
    assert root == "(((ab)(cd))((ef)(gh)))" and widths == [4, 2, 1]          # the default 8-leaf, depth-3 tree
    levels = [lv for lv, _, _ in log]
    assert levels == sorted(levels), "a level started before the one below it finished"
    assert {(lv, i) for lv, i, _ in log} == {(0, 0), (0, 1), (0, 2), (0, 3), (1, 0), (1, 1), (2, 0)}
    assert all(p == f"s{i % 3}" for _, i, p in log)                          # chunk i of a level runs on stream i mod S
    root, widths = aggregate_tree(list("abcde"), 3, prove_chunk, ["s0"])
    assert root == "((abc)(de))" and widths == [2, 1]
    root, widths = aggregate_tree(["a"], 2, prove_chunk, ["s0"])
    assert root == "(a)" and widths == [1]
    with pytest.raises(ValueError):
        aggregate_tree([], 2, prove_chunk, ["s0"])
    with pytest.raises(ValueError):
        aggregate_tree(["a"], 1, prove_chunk, ["s0"])


def test_context_pool_builds_each_circuit_once_per_device():
    sys.path.insert(0, os.path.join(ROOT, "zk-circuits_b200"))
    from zkb200.batch import ContextPool

    built = []

    def make(device):
        built.append(device)
        return object()

    pool = ContextPool(streams=2)
    a = pool.get([1, 2, 3, 4], 0, make)
    assert len(a) == 2 and built == [0, 0]
    assert pool.get((1, 2, 3, 4), 0, make) is a and built == [0, 0]          # same circuit, same GPU: cached
    b = pool.get([1, 2, 3, 4], 1, make)                                      # same circuit on another GPU
    c = pool.get([9, 2, 3, 4], 0, make)                                      # next level of the tree: another circuit
    assert b is not a and c is not a and len(pool) == 3 and (pool.hits, pool.misses) == (1, 3)
    with pytest.raises(ValueError):
        ContextPool(0)


def test_aggregator_padding_and_public_input_split():
    """util.rs:11-29 and inputs.rs:57-89 semantics."""
    sys.path.insert(0, os.path.join(ROOT, "zk-circuits_b200"))
    from zkb200 import batch

    cfg = batch.TreeAggregationConfig()
    assert (cfg.tree_branching_factor, cfg.tree_depth, cfg.num_leaf_proofs) == (2, 3, 8)
    assert batch.TreeAggregationConfig(3, 2).num_leaf_proofs == 9
    assert batch.pad_with_dummy_proofs([b"a", b"b"], 4, b"D") == [b"a", b"b", b"D", b"D"]
    assert batch.pad_with_dummy_proofs([], 2, b"D") == [b"D", b"D"]
    with pytest.raises(ValueError, match="more than the maximum"):
        batch.pad_with_dummy_proofs([b"a"] * 5, 4, b"D")
    assert batch.split_aggregated_public_inputs(range(32), 16, 2) == [list(range(16)), list(range(16, 32))]
    with pytest.raises(ValueError, match="aggregated public inputs should contain: 128"):
        batch.split_aggregated_public_inputs(range(127), 16, 8)


def test_aggregate_forest_is_dependency_driven():
    """Every chunk proof receives exactly its children in order, a parent never starts before its children finished, trees are
    padded with the dummy proof, and chunks of different trees overlap (the level barrier of one tree does not idle the
    workers)."""
    sys.path.insert(0, os.path.join(ROOT, "zk-circuits_b200"))
    import threading
    import time
    from zkb200 import batch

    cfg = batch.TreeAggregationConfig(2, 3)
    forest = [[f"t{t}l{i}" for i in range(8 if t != 2 else 5)] for t in range(4)]
    lock, active, peak = threading.Lock(), [0], [0]

    def prove_chunk(w, chunk, level, index, tree):
        with lock:
            active[0] += 1
            peak[0] = max(peak[0], active[0])
        time.sleep(0.02)
        with lock:
            active[0] -= 1
        return "(" + "+".join(chunk) + ")"

    roots, stats = batch.aggregate_forest(forest, cfg, prove_chunk, workers=4, dummy_proof="D")
    assert sorted(roots) == [0, 1, 2, 3]
    assert roots[0] == "(((t0l0+t0l1)+(t0l2+t0l3))+((t0l4+t0l5)+(t0l6+t0l7)))"
    assert roots[2] == "(((t2l0+t2l1)+(t2l2+t2l3))+((t2l4+D)+(D+D)))"
    spans = {(t, l, i): (a, b) for t, l, i, a, b in stats["spans"]}
    assert len(spans) == 4 * 7
    for (t, l, i), (a, b) in spans.items():
        if l > 0:
            assert a >= max(spans[(t, l - 1, 2 * i)][1], spans[(t, l - 1, 2 * i + 1)][1])
    assert peak[0] == 4 and len(stats["level_concurrency"]) == 3
    # 28 chunk proofs of 20 ms on 4 workers: close to 7 x 20 ms, far below the 4 trees x 3 levels x 20 ms of per-tree barriers
    t0 = min(a for a, _ in spans.values())
    t1 = max(b for _, b in spans.values())
    assert t1 - t0 < 0.21

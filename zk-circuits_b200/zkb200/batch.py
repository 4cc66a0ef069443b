"""Batch proving across GPUs — the host-side mirror of the reference aggregator's fan-out over independent leaf proofs
(/root/reference/wormhole/aggregator/src/circuits/tree.rs:93-103 maps chunks over rayon; each call owns its circuit data and
witness). Proofs are independent units, so ranks share NO data-path collective: rank r proves its slice on its own GPU with
`streams` prover contexts in flight, and the 130-150 KB proofs are gathered on the host.

Works with any initialised torch.distributed backend (NCCL on the GPU box, gloo in the CPU tests) or with none (world 1).
"""
import threading

try:
    import torch.distributed as dist
except Exception:  # torch absent: single-process use only
    dist = None


def world():
    if dist is not None and dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(total, rank, world_size):
    """Contiguous slice [lo, hi) of `total` proofs owned by `rank`; sizes differ by at most one, earlier ranks get the extra."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world size")
    base, extra = divmod(total, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def run_streams(jobs, provers):
    """Run jobs (list of zero-argument-free callables taking a prover) on len(provers) host threads, job i on prover
    i % len(provers) in submission order; returns results in job order. The C ABI call drops the GIL."""
    results = [None] * len(jobs)
    errors = []

    def worker(s):
        try:
            for i in range(s, len(jobs), len(provers)):
                results[i] = jobs[i](provers[s])
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(s,)) for s in range(min(len(provers), len(jobs)))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return results


def prove_batch(witnesses, provers, prove_fn, gather=True):
    """Prove `witnesses` (a list, identical on every rank) as one distributed batch.

    provers   this rank's prover contexts (one per stream), e.g. [ProverCircuit(..., device=local_rank)] * B distinct objects
    prove_fn  prove_fn(prover, witness, global_index) -> proof bytes
    Returns the full list of proofs in witness order on rank 0 (None elsewhere) when gather is set, else this rank's slice.
    """
    rank, ws = world()
    lo, hi = shard_range(len(witnesses), rank, ws)
    jobs = [(lambda p, i=i: prove_fn(p, witnesses[i], i)) for i in range(lo, hi)]
    mine = run_streams(jobs, provers)
    if not gather:
        return mine
    if ws == 1:
        return mine
    parts = [None] * ws if rank == 0 else None
    dist.gather_object(mine, parts, dst=0)
    if rank != 0:
        return None
    return [p for part in parts for p in part]


class ContextPool:
    """Prover contexts cached by (circuit digest, device): SURVEY.md §8f rank 1. The reference builds its circuit inside every
    bench iteration (/root/reference/wormhole/prover/benches/prover.rs:14-19) and once per chunk in `aggregate_chunk`
    (aggregator/src/circuits/tree.rs:111-127) although every chunk of a level shares one CommonCircuitData (tree.rs:64-70);
    a GPU context costs a constants/sigmas commitment plus ~0.6 GB of buffers, so it is built once per distinct circuit and
    handed out again. `make(device)` builds one context; `streams` contexts are kept per key."""

    def __init__(self, streams=1):
        if streams < 1:
            raise ValueError("streams must be positive")
        self.streams = streams
        self._lock = threading.Lock()
        self._ctx = {}
        self.hits = self.misses = 0

    def get(self, circuit_digest, device, make):
        key = (tuple(int(x) for x in circuit_digest), int(device))
        with self._lock:
            got = self._ctx.get(key)
            if got is not None:
                self.hits += 1
                return got
            self.misses += 1
        made = [make(device) for _ in range(self.streams)]
        with self._lock:
            return self._ctx.setdefault(key, made)

    def __len__(self):
        return len(self._ctx)

    def clear(self):
        with self._lock:
            self._ctx.clear()


def aggregate_tree(leaf_proofs, branching, prove_chunk, provers):
    """Level-by-level aggregation of `leaf_proofs` into one root proof — the host-side mirror of the reference's
    `aggregate_to_tree` / `aggregate_level` (/root/reference/wormhole/aggregator/src/circuits/tree.rs:55-103): a level's
    proofs are cut into chunks of `branching` (the last chunk may be short), every chunk of a level is proved concurrently
    (the reference maps them over rayon; here over this rank's prover contexts = CUDA streams), and a level starts only when
    the one below is complete, so the GPU sees 4, 2, 1 concurrent proofs for the default 8-leaf tree.

    prove_chunk(prover, chunk, level, index) -> proof   (level 0 = the chunks made of leaf proofs)
    Returns (root_proof, [number of chunk proofs per level])."""
    if branching < 2:
        raise ValueError("branching factor must be at least 2")
    if not leaf_proofs:
        raise ValueError("no leaf proofs")
    level, depth, widths = list(leaf_proofs), 0, []
    while len(level) > 1 or depth == 0:
        chunks = [level[i:i + branching] for i in range(0, len(level), branching)]
        jobs = [(lambda p, c=c, d=depth, i=i: prove_chunk(p, c, d, i)) for i, c in enumerate(chunks)]
        level = run_streams(jobs, provers)
        widths.append(len(chunks))
        depth += 1
    return level[0], widths


def max_over_ranks(values, device=None):
    """Element-wise max of a list of floats over all ranks (timing rule: a multi-GPU number is the slowest rank's)."""
    rank, ws = world()
    if ws == 1:
        return list(values)
    import torch

    t = torch.tensor(list(values), dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


# ---------------------------------------------------------------------------------------------------------------------------
# Aggregator orchestration: the host logic around the chunk proofs of
# /root/reference/wormhole/aggregator/src/{aggregator.rs:74-93, util.rs:11-29, circuits/tree.rs:17-103}, scheduled so that a box
# of GPUs stays busy: trees are dealt to ranks (one process per GPU, no proof ever crosses a GPU), and inside a rank the chunk
# proofs of ALL its trees form one dependency graph — a chunk is ready as soon as its `branching` children are proved — fed to
# the rank's prover contexts. With the reference's level-by-level schedule an 8-leaf tree shows the GPU 4, then 2, then 1
# concurrent proofs; with several trees in flight the narrow upper levels of one tree overlap the wide lower levels of the next.
# ---------------------------------------------------------------------------------------------------------------------------
class TreeAggregationConfig:
    """tree.rs:32-52: num_leaf_proofs = branching ** depth (default 2, 3 -> 8 leaves, 4 + 2 + 1 chunk proofs)."""

    def __init__(self, tree_branching_factor=2, tree_depth=3):
        if tree_branching_factor < 2 or tree_depth < 1:
            raise ValueError("branching factor >= 2 and depth >= 1")
        self.tree_branching_factor = tree_branching_factor
        self.tree_depth = tree_depth
        self.num_leaf_proofs = tree_branching_factor ** tree_depth


def pad_with_dummy_proofs(proofs, proof_len, dummy_proof):
    """util.rs:11-29: append copies of the dummy leaf proof up to proof_len; more proofs than that is an error."""
    proofs = list(proofs)
    if len(proofs) > proof_len:
        raise ValueError("proofs to aggregate was more than the maximum allowed")
    return proofs + [dummy_proof] * (proof_len - len(proofs))


def split_aggregated_public_inputs(public_inputs, leaf_pi_len, num_leaves):
    """circuit/src/inputs.rs:57-89 (`try_from_aggregated`): the root proof's public inputs are the leaves' public inputs in
    order (every chunk circuit registers its children's, tree.rs:121-123); returns num_leaves slices of leaf_pi_len."""
    pis = list(public_inputs)
    expected = leaf_pi_len * num_leaves
    if len(pis) != expected:
        raise ValueError(f"aggregated public inputs should contain: {expected} (= {num_leaves} leaves x {leaf_pi_len} fields), "
                         f"got: {len(pis)}")
    return [pis[i:i + leaf_pi_len] for i in range(0, expected, leaf_pi_len)]


def tree_for_rank(num_trees, rank, world_size):
    """trees dealt round-robin: tree t is aggregated wholly on rank t % world_size"""
    return [t for t in range(num_trees) if t % world_size == rank]


def aggregate_forest(forest, config, prove_chunk, workers, dummy_proof=None, clock=None):
    """Aggregate this rank's share of `forest` (a list of leaf-proof lists, identical on every rank) — every tree padded to
    config.num_leaf_proofs with dummy_proof as `aggregate()` does (aggregator.rs:79-83) — with a dependency-driven schedule
    over `workers` concurrent prove calls.

    prove_chunk(worker_index, chunk, level, index, tree) -> proof     (level 0 = chunks of leaf proofs)
    Returns ({tree: root_proof} for this rank's trees, stats) where stats['level_concurrency'][k] is the average number of
    level-k chunk proofs in flight while any was (the per-level occupancy of this GPU's proof contexts) and stats['spans'] the
    (tree, level, index, start, end) records."""
    import queue
    import time as _time

    clock = clock or _time.perf_counter
    rank, ws = world()
    mine = tree_for_rank(len(forest), rank, ws)
    b, depth = config.tree_branching_factor, config.tree_depth
    done = {}                    # (tree, level, index) -> proof; level -1 = (padded) leaves
    pending = {}                 # (tree, level, index) -> children still missing
    ready = queue.Queue()
    lock = threading.Lock()
    for t in mine:
        leaves = pad_with_dummy_proofs(forest[t], config.num_leaf_proofs, dummy_proof) if dummy_proof is not None else list(forest[t])
        if len(leaves) != config.num_leaf_proofs:
            raise ValueError("a tree needs exactly branching ** depth leaf proofs (pass dummy_proof to pad)")
        for i, p in enumerate(leaves):
            done[(t, -1, i)] = p
        for i in range(config.num_leaf_proofs // b):
            ready.put((t, 0, i))
        for lvl in range(1, depth):
            for i in range(config.num_leaf_proofs // b ** (lvl + 1)):
                pending[(t, lvl, i)] = b
    total = sum(config.num_leaf_proofs // b ** (lvl + 1) for lvl in range(depth)) * len(mine)
    spans, errors = [], []
    remaining = [total]

    def worker(w):
        while True:
            job = ready.get()
            if job is None:
                return
            t, lvl, i = job
            try:
                chunk = [done[(t, lvl - 1, b * i + k)] for k in range(b)]
                t0 = clock()
                proof = prove_chunk(w, chunk, lvl, i, t)
                t1 = clock()
            except Exception as e:  # noqa: BLE001
                errors.append(e)
                for _ in range(workers):
                    ready.put(None)
                return
            with lock:
                done[(t, lvl, i)] = proof
                spans.append((t, lvl, i, t0, t1))
                remaining[0] -= 1
                parent = (t, lvl + 1, i // b)
                if parent in pending:
                    pending[parent] -= 1
                    if pending[parent] == 0:
                        ready.put(parent)
                if remaining[0] == 0:
                    for _ in range(workers):
                        ready.put(None)

    if total == 0:
        return {}, {"level_concurrency": [], "spans": []}
    threads = [threading.Thread(target=worker, args=(w,)) for w in range(workers)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    if errors:
        raise errors[0]
    conc = []
    for lvl in range(depth):
        s = [(t0, t1) for (_, l, _, t0, t1) in spans if l == lvl]
        busy = sum(t1 - t0 for t0, t1 in s)
        span = max(t1 for _, t1 in s) - min(t0 for t0, _ in s)
        conc.append(busy / span if span > 0 else float(len(s)))
    return {t: done[(t, depth - 1, 0)] for t in mine}, {"level_concurrency": conc, "spans": spans}

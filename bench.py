#!/usr/bin/env python3
"""bench.py — wormhole prove throughput on B200 through libzkb200.so (BASELINE.json metric:
"Wormhole prove ms & proofs/sec at 1/2/4/8 B200; LDE-NTT GB/s vs HBM peak").

    python bench.py --gpus N --steps K --warmup W            # CUDA prover (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle's restated Plonky2 prover

A step = `--streams` (default 8) independent proofs per GPU of the synthetic wormhole-shaped zk circuit (config #1:
n = 2^14, 135 wires, 6-gate set, 28 FRI queries, 16 PoW bits; proof = 148 932 bytes), each on its own prover context
and CUDA stream, driven by one host thread each — the way the reference's rayon callers invoke prove(). `value` is
measured with the witnesses resident in HBM, `e2e` through the host-buffer C-ABI call zkb_prove() (H2D of the wire
matrix and D2H of the proof inside the timed region); `prove_ms_single_stream` is the latency of one proof alone.
Multi-GPU is replica mode (independent proofs, no data-path collective): weak scaling.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "zk-circuits_b200"))

WORKLOAD = "wormhole_zk_synth_n2^14"
NCU_LDE_TRAFFIC_BYTES = 19055104 + 83806720   # dram__bytes_read.sum + dram__bytes_write.sum, one lde_block_kernel_t<3, 1024, 14> launch (profiles/r02_ncu_lde_block_v2.md)
METRIC = "wormhole_proofs_per_sec"
UNIT = "proofs/s"


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = "index,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().splitlines()
                for line in out:
                    f = [x.strip() for x in line.split(",")]
                    self.samples.append(float(f[1]))
                    self.max_mhz = float(f[2])
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                        if v.lower().startswith("active"):
                            self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def synth_columns(np, lg_n, cols):
    """config #3 inputs: x[c][i] = SplitMix64-finalizer(SEED + c n + i) mod p."""
    nn = 1 << lg_n
    idx = np.arange(nn * cols, dtype=np.uint64) + np.uint64(0xB200000000000001)
    with np.errstate(over="ignore"):
        z = idx
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return np.where(z >= np.uint64(0xFFFFFFFF00000001), z - np.uint64(0xFFFFFFFF00000001), z).reshape(cols, nn)


def golden_cap(lg_n, cols):
    """16 cap digests of the config-#3 commitment at this shape, computed once by the CPU oracle (tests/golden/make_caps.py);
    None if the shape has no stored cap."""
    try:
        with open(os.path.join(ROOT, "tests", "golden", "config3_caps.json")) as f:
            e = json.load(f).get(f"{lg_n}x{cols}")
        return None if e is None else [[int(x, 16) for x in d] for d in e["cap"]]
    except Exception:
        return None


def cap_matches(np, cap, lg_n, cols):
    want = golden_cap(lg_n, cols)
    if want is None:
        return None
    ok = bool(np.array_equal(np.asarray(cap, dtype=np.uint64), np.asarray(want, dtype=np.uint64)))
    if not ok:
        raise SystemExit(f"bench.py: commitment cap at 2^{lg_n} x {cols} differs from the oracle's golden cap")
    return ok


def bench_config(nw, proof_bytes, B):
    """The `config` object, identical in the GPU arm and the reference arm (same workload, same step definition)."""
    return {"workload": WORKLOAD, "degree_bits": 14, "zero_knowledge": True, "num_wires": nw, "proof_bytes": proof_bytes,
            "proofs_per_step_per_gpu": B,
            "l2": "no flush: one proof streams ~0.5 GB of LDE/leaf data, far above the 126 MB L2"}


def reference_arm(args, rank):
    """CPU arm: the reference's own prover cannot be built here (Rust crate qp-plonky2, no toolchain), so this times the
    oracle's restatement of it (kind "port") on every host core, same circuit. The port's hot loop — Poseidon, 3.6 M
    permutations per proof — runs in its optimised AVX2 form (oracle/poseidon_fast.hpp; qp-plonky2's is hand-tuned too), the
    rest is the readable restatement with OpenMP over columns / points. A step is a bounded sample of the GPU arm's step: ONE
    proof of the `proofs_per_step_per_gpu` (proofs are independent, the metric is proofs/s)."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O

    O.build()
    O.set_num_threads(os.cpu_count() or 1)     # torchrun exports OMP_NUM_THREADS=1; the CPU arm gets every host core
    O.set_fast(True)
    s = O.Synth(zk=True, seed=1, **O.Synth.WORMHOLE)
    c = O.Circuit(s.common, s.const_sigma_values)
    steps, warm = max(1, args.steps), max(0, args.warmup)
    B = args.streams if args.streams > 0 else 8
    for i in range(warm):
        c.prove(s.wires, s.public_inputs, salt_seed=7 + i)
    t0 = time.perf_counter()
    for i in range(steps):
        proof = c.prove(s.wires, s.public_inputs, salt_seed=100 + i)
    dt = time.perf_counter() - t0
    assert c.verify(proof) == ""
    val = steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": 1000 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 (Goldilocks field, F_p^2 extension)", "data": "synthetic",
        "config": bench_config(len(s.wires), len(proof), B),
        "parallelism": "host cores (OpenMP); GPUs unused",
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": O.num_threads(), "kind": "port",
                         "sample": f"{steps} full proofs of the bench circuit (one proof per step = 1/{B} of the GPU arm's step) with the "
                                   "oracle's restated Plonky2 prover, AVX2 Poseidon + 4-lane AVX2 gate evaluation + OpenMP"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the process's original stdout; everything else written to fd 1 by libraries
    (e.g. NCCL's version banner) was redirected to stderr by quiet_stdout()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def quiet_stdout():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="zkb200", choices=["zkb200", "reference"])
    ap.add_argument("--streams", type=int, default=0,
                    help="proofs in flight per GPU (independent prover contexts, one host thread each); 0 = 8. With more proofs "
                         "in flight than host cores per rank the stream waits sleep-poll instead of spinning (8 GPUs / 32 cores: "
                         "1622 proofs/s with 8 streams sleep-polling, 1419 with 4 streams spinning)")
    ap.add_argument("--driver", default="engine", choices=["engine", "threads"],
                    help="engine: zkb_engine (one driver thread per GPU steps all proof contexts); threads: one blocking host thread "
                         "per proof in flight (round 1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the LDE/Merkle microbench points (config #3)")
    ap.add_argument("--full-grid", action="store_true",
                    help="config #3 as SURVEY 8d states it: n in 2^14..2^22 x cols in 100/135/200/400 (2^22 x 200/400 left out: > 120 GB)")
    ap.add_argument("--no-aggregation", action="store_true", help="skip the aggregation-tree measurement (config #4)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        reference_arm(args, rank)
        return

    import numpy as np
    import torch
    import zkb200 as Z
    from zkb200 import batch as zbatch   # same max-over-ranks helper the CPU gloo test exercises

    if not torch.cuda.is_available() or Z.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device — the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    W = max(3, args.warmup)
    K = max(1, args.steps)
    B = args.streams if args.streams > 0 else 8
    synth = Z.SynthCircuit(zk=True, seed=1, **Z.WORMHOLE)
    n, nw = synth.n, synth.wires.shape[0]
    # B independent prover contexts per GPU (one stream each), driven by B host threads: the reference's callers
    # already prove concurrently from rayon workers (aggregator/src/circuits/tree.rs:93-103), one circuit per call.
    t0 = time.perf_counter()
    circs = [Z.ProverCircuit(synth.common, synth.const_sigma_values, is_values=True, device=local_rank) for _ in range(B)]
    t_create = (time.perf_counter() - t0) / B      # includes the first context's one-time table setup
    t0 = time.perf_counter()
    Z.ProverCircuit(synth.common, synth.const_sigma_values, is_values=True, device=local_rank)
    t_create_warm = time.perf_counter() - t0
    circ = circs[0]
    # pinned host staging for the e2e arm (the reference-side caller would hand over its witness like this)
    pinned = [torch.empty((nw, n), dtype=torch.int64, pin_memory=True) for _ in range(B)]
    for t_ in pinned:
        t_.numpy().view(np.uint64)[:] = synth.wires
    host_addr = [t_.data_ptr() for t_ in pinned]
    pis = synth.public_inputs
    outs = [np.zeros(circ.proof_size, dtype=np.uint8) for _ in range(B)]

    def run_parallel(fn):
        """fn(b) on B host threads (ctypes drops the GIL inside the C ABI call); returns when all are done."""
        if B == 1:
            fn(0)
            return
        errs = []

        def wrap(b):
            try:
                fn(b)
            except Exception as e:  # noqa: BLE001
                errs.append(e)

        th = [threading.Thread(target=wrap, args=(b,)) for b in range(B)]
        for t_ in th:
            t_.start()
        for t_ in th:
            t_.join()
        if errs:
            raise errs[0]

    # ---- single-stream latency + per-stage device times (CUDA events on the circuit's stream) ----
    for b in range(B):
        circs[b].upload_witness(host_addr[b])
    for i in range(W):
        circ.prove_resident(pis, salt_seed=1000 * rank + i, out=outs[0])
    stage_sum = {}
    barrier()
    launches0 = Z.kernel_launch_count()
    t0 = time.perf_counter()
    for i in range(K):
        circ.prove_resident(pis, salt_seed=2000 * (rank + 1) + i, out=outs[0])
        for k, v in circ.timings().items():
            stage_sum[k] = stage_sum.get(k, 0.0) + v
    torch.cuda.synchronize()
    t_single = time.perf_counter() - t0
    launches = (Z.kernel_launch_count() - launches0) // K
    # the same with ZKB_CHECK_WITNESS (every constraint evaluated on the subgroup before Z is committed)
    circ.prove_resident(pis, salt_seed=1, check_witness=True, out=outs[0])
    t0 = time.perf_counter()
    for i in range(K):
        circ.prove_resident(pis, salt_seed=2000 * (rank + 1) + i, check_witness=True, out=outs[0])
    t_single_checked = time.perf_counter() - t0
    # ---- device-resident throughput arm (`value`): B proofs in flight per GPU ----
    # Driver: the proof ENGINE (default) — B prover contexts stepped by ONE host thread inside the library, proofs submitted
    # asynchronously (zkb_engine_*); `--driver threads` is round 1's form, one blocking host thread per proof in flight.
    use_engine = args.driver == "engine"
    eng = None
    if use_engine:
        for c in circs[1:]:
            c.close()
        del circs[1:]
        eng = Z.Engine(synth.common, synth.const_sigma_values, is_values=True, device=local_rank, contexts=B, slots=2 * B, num_wires=nw)
        held = [eng.acquire() for _ in range(2 * B)]          # fill every pinned slot once (what a witness generator would do)
        for slot, buf in held:
            buf[:] = synth.wires
        for slot, _ in held:
            eng.release(slot)
        eouts = [np.zeros(eng.proof_size, dtype=np.uint8) for _ in range(2 * B)]

        def engine_step(seed, resident):
            """one step = B proofs: submit all, then wait for all (the engine overlaps them on its B contexts)"""
            slots = []
            for b in range(B):
                slot, _ = eng.acquire()
                eng.submit(slot, pis, salt_seed=seed + 100 * b, resident=resident, out=eouts[slot])
                slots.append(slot)
            return [eng.wait(sl) for sl in slots]

        engine_step(1, False)                                  # every context now holds the witness
        for i in range(W):
            engine_step(5000 + i, True)
    else:
        for i in range(W):
            run_parallel(lambda b: circs[b].prove_resident(pis, salt_seed=5000 + 100 * b + i, out=outs[b]))
    barrier()

    class DeviceTimer:
        """CUDA events on torch's current stream bracketing the timed region: the start event is recorded after a device
        synchronise (nothing in flight), the end event after every proving stream has drained, so the elapsed time is the
        device-timeline time of the K steps, host-side Fiat-Shamir gaps included. The host clock is kept as a cross-check."""

        def __enter__(self):
            torch.cuda.synchronize()
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.t0 = time.perf_counter()
            self.e0.record()
            return self

        def __exit__(self, *a):
            torch.cuda.synchronize()
            self.e1.record()
            self.e1.synchronize()
            self.host_s = time.perf_counter() - self.t0
            self.seconds = self.e0.elapsed_time(self.e1) * 1e-3

    with ClockSampler(local_rank) as clk:
        with DeviceTimer() as tm_res:
            for i in range(K):
                if use_engine:
                    engine_step(2000 * (rank + 1) + i, True)
                else:
                    run_parallel(lambda b: circs[b].prove_resident(pis, salt_seed=2000 * (rank + 1) + 100 * b + i, out=outs[b]))
        t_res = tm_res.seconds
        barrier()
        # ---- end-to-end arm: host buffers in, proof bytes out, through the public C-ABI call (zkb_engine_submit / wait, or
        # zkb_prove with --driver threads): H2D of the wire matrix from pinned host memory and D2H of the proof inside ----
        proofs = [None] * B

        def e2e_step(b, seed):
            proofs[b] = circs[b].prove(host_addr[b], pis, salt_seed=seed + 100 * b)

        for i in range(2):
            if use_engine:
                engine_step(i, False)
            else:
                run_parallel(lambda b: e2e_step(b, i))
        barrier()
        with DeviceTimer() as tm_e2e:
            for i in range(K):
                if use_engine:
                    got = engine_step(3000 * (rank + 1) + i, False)
                else:
                    run_parallel(lambda b: e2e_step(b, 3000 * (rank + 1) + i))
        t_e2e = tm_e2e.seconds
    proof = got[0].tobytes() if use_engine else proofs[0]
    barrier()

    t_res, t_e2e, t_single = zbatch.max_over_ranks([t_res, t_e2e, t_single], device="cuda")
    stages = {k: v / K for k, v in stage_sum.items()}

    # ---- N > 1 only: ONE large commitment sharded by LDE coset across the GPUs (the mode where the path has a real
    # exchange step: one NCCL all-gather of 16 cap digests; SURVEY.md §8e(2), config #3 shape) ----
    sharded_line = None
    if world > 1 and not args.no_sweep and (8 % world) == 0:
        # the exchange runs INSIDE libzkb200.so (zkb_comm_* / zkb_commit_sharded / zkb_quotient_chunks_sharded: NCCL over
        # NVLink); torch.distributed only carries the 128-byte NCCL unique id from rank 0 to the other ranks
        uid = torch.from_numpy(Z.comm_unique_id() if rank == 0 else np.zeros(128, dtype=np.uint8)).cuda()
        dist.broadcast(uid, 0)
        comm = Z.Comm(uid.cpu().numpy(), world, rank, device=local_rank)
        lg_n, cols = 20, 100
        vals = synth_columns(np, lg_n, cols)
        barrier()
        t0 = time.perf_counter()
        cap, tm = comm.commit(vals, 3, 4, reps=2)
        torch.cuda.synchronize()
        t_wall = time.perf_counter() - t0
        lde_ms, merkle_ms, gather_ms, t_wall = zbatch.max_over_ranks([tm["lde_ms"], tm["merkle_ms"], tm["gather_ms"], t_wall], device="cuda")
        nn = 1 << lg_n
        sharded_line = {"lg_n": lg_n, "cols": cols, "ranks": world, "blocks_per_rank": 8 // world, "lde_ms": lde_ms,
                        "merkle_ms": merkle_ms, "peer_windows": tm["peer_windows"],
                        ("peer_copies_ms_overlapped" if tm["peer_windows"] else "coeff_allgather_ms_overlapped"): gather_ms,
                        "lde_gbs_aggregate": 80 * nn * cols / (lde_ms * 1e-3) / 1e9,
                        "perms_per_sec_aggregate": (8 * nn * ((cols + 7) // 8) + 8 * nn - 16) / (merkle_ms * 1e-3),
                        "collectives": ("inside libzkb200.so: column-sharded iNTT into CUDA-IPC windows, each rank pulls the peers' "
                                        "coefficients (8 n cols bytes) over NVLink with its copy engines, one slice ahead of the LDE; NCCL: one barrier + the all-gather of 16 x 32 B cap digests"
                                        if tm["peer_windows"] else
                                        "NCCL inside libzkb200.so: all-gather of the coefficients (column-sharded iNTT, 8 n cols bytes, "
                                        "4 chunks pipelined behind the LDE) + all-gather of 16 x 32 B cap digests"),
                        "wall_ms_3_passes_incl_h2d_of_the_rank_slice": 1000 * t_wall,
                        "cap_word0": int(cap[0, 0]), "cap_equals_oracle_golden": cap_matches(np, cap, lg_n, cols)}
        # quotient chunks from coset-local evaluations: one all-to-all (2 challenges, n = 2^20); input = random field elements
        # (timing only: exactness is tests/test_gpu_new_paths.py + the gloo test of the same exchange)
        qn = 1 << 20
        qv = synth_columns(np, 20, 2 * (8 // world)).reshape(2, (8 // world) * qn)
        comm.quotient_chunks(qv, qn, 3)                      # untimed: NCCL sets up its point-to-point channels on first use
        qi = qx = float("inf")
        for _ in range(3):                                   # best of three (max over ranks each): one call's exchange time
            barrier()                                        # also holds whatever skew the ranks arrive with
            _, qt = comm.quotient_chunks(qv, qn, 3)
            a, b = zbatch.max_over_ranks([qt["interpolate_ms"], qt["exchange_ms"]], device="cuda")
            qi, qx = min(qi, a), min(qx, b)
        sharded_line["quotient_chunks_n2^20"] = {"coset_intt_ms": qi, "all_to_all_plus_solve_ms": qx,
                                                 "bytes_exchanged_per_rank": int(2 * (8 // world) * qn * 8 * (world - 1) / world)}
        comm.close()

    # ---- config #4/#5: a FOREST of aggregation trees over the GPUs. 8-leaf trees (branching 2, depth 3:
    # 4 + 2 + 1 chunk proofs each, aggregator/src/circuits/tree.rs:17-20) dealt round-robin to the ranks; inside a rank all chunk
    # proofs of its trees form one dependency graph fed to a proof engine (zkb200.batch.aggregate_forest), so the narrow upper
    # levels of one tree overlap the wide lower levels of the next. Chunk circuits: recursion-shaped synthetic, zero-knowledge
    # (n = 2^14), witness generation (Rust side) not included. ----
    forest_line = None
    if not args.no_aggregation:
        rs = Z.SynthCircuit(zk=True, seed=4, **Z.SynthCircuit.RECURSION)
        S = min(4, B)
        reng = Z.Engine(rs.common, rs.const_sigma_values, is_values=True, device=local_rank, contexts=S, slots=2 * S, num_wires=rs.wires.shape[0])
        held = [reng.acquire() for _ in range(2 * S)]
        for slot, buf in held:
            buf[:] = rs.wires                                  # every pinned slot holds the chunk witness (stand-in for the generator)
        for slot, _ in held:
            reng.release(slot)

        def prove_chunk_engine(wk, chunk, level, index, tree):
            slot, _ = reng.acquire()
            reng.submit(slot, rs.public_inputs, salt_seed=100000 * tree + 1000 * level + index)
            return reng.wait(slot).tobytes()

        cfg = zbatch.TreeAggregationConfig(2, 3)
        trees_per_rank = 4
        forest = [[b"leaf"] * 8 for _ in range(trees_per_rank * world)]
        zbatch.aggregate_forest(forest, cfg, prove_chunk_engine, workers=2 * S)
        barrier()
        t0 = time.perf_counter()
        roots, fst = zbatch.aggregate_forest(forest, cfg, prove_chunk_engine, workers=2 * S)
        torch.cuda.synchronize()
        t_forest = time.perf_counter() - t0
        t0 = time.perf_counter()
        one_root, one = zbatch.aggregate_forest(forest[:1] if rank == 0 else [], cfg, prove_chunk_engine, workers=2 * S)
        t_one = time.perf_counter() - t0
        t_forest, t_one = zbatch.max_over_ranks([t_forest, t_one], device="cuda")
        forest_line = {"workload": "aggregation_forest_8_leaf_trees_recursion_zk_synth_n2^14", "trees": len(forest), "ranks": world,
                       "trees_per_rank": trees_per_rank, "chunk_proofs_per_tree": 7, "contexts_per_gpu": S,
                       "forest_ms": 1000 * t_forest, "trees_per_sec": len(forest) / t_forest,
                       "chunk_proofs_per_sec": 7 * len(forest) / t_forest, "single_tree_ms_one_gpu": 1000 * t_one,
                       "level_concurrency_rank0": fst["level_concurrency"],
                       "single_tree_level_concurrency": one["level_concurrency"] if rank == 0 else None,
                       "root_proof_bytes": len(next(iter(roots.values()))),
                       "note": "level_concurrency[k] = average number of level-k chunk proofs in flight on a GPU while any was; a lone "
                               "tree shows <= 4, 2, 1 (the reference's level barrier), the forest keeps the contexts filled"}
        reng.close()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = hbm_peak()
    # roofline of the coset-LDE launch on the wires batch (lde_block_kernel: read n coefficients, write 8n values per
    # column = 72 n bytes; the interpolation before it is timed separately as wires_intt, 16 n bytes per column)
    lde_bytes = 72 * n * nw
    lde_ms = stages["wires_lde"]
    lde_gbs = lde_bytes / (lde_ms * 1e-3) / 1e9
    salt = 4
    leaves = n * 8
    perms = leaves * ((nw + salt + 7) // 8) + leaves - 16
    pos_ms = stages["wires_merkle"]
    line = {
        "metric": METRIC, "value": world * K * B / t_res, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": 1000 * t_res / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64 (Goldilocks field, F_p^2 extension)", "data": "synthetic",
        "config": bench_config(nw, len(proof), B),
        "parallelism": f"replica x{world} (no collective), {B} proof streams per GPU", "driver": args.driver,
        "timer": "CUDA events bracketing the K steps (recorded after a device synchronise on both sides, rank barrier before; a step "
                 "contains host-side Fiat-Shamir work between launches, which the events include), max over ranks; stage_ms and the "
                 "roofline numbers are CUDA events on the proving stream",
        "host_clock_ms_per_step": 1000 * tm_res.host_s / K,
        "circuit_create_ms": {"first_contexts_avg": 1000 * t_create, "warm": 1000 * t_create_warm,
                              "note": "zkb_circuit_create from host values: upload + constants/sigmas iNTT, LDE and Merkle tree on the "
                                      "device + every work buffer (SURVEY 8f rank 1; cached per circuit by zkb200.batch.ContextPool). One "
                                      "cudaMalloc of ~0.6 GB dominates and varies by box (2 ms to 400 ms, fresh memory being cleared); the "
                                      "rest is ~5 ms (ZKB_TRACE=1)"},
        "prove_ms_single_stream": 1000 * t_single / K, "prove_ms_single_stream_with_witness_check": 1000 * t_single_checked / K, "device_ms_per_proof": stages["total"], "stage_ms": stages,
        "e2e": {"value": world * K * B / t_e2e, "unit": UNIT, "ms_per_step": 1000 * t_e2e / K,
                "h2d_bytes_per_step": int(B * (nw * n * 8 + pis.size * 8)), "d2h_bytes_per_step": int(B * len(proof))},
        "gpu_launches": int(launches) * B,
        "roofline": {"kernel": "lde_block_kernel_t<3, 1024, 14>: coset pre-scale + 8 x NTT of the 135 wire columns, n = 2^14 (one launch)",
                     "bound": "hbm", "achieved": lde_gbs, "peak": peak, "unit": "GB/s", "frac": lde_gbs / peak,
                     "traffic": NCU_LDE_TRAFFIC_BYTES,
                     "note": "integer-issue bound in practice (322 thread instructions per element and transform = 36 per algorithmic byte vs "
                             "5.7 the chip can issue per HBM byte; operation-count floor ~32: DESIGN.md 4.2; measured bound of the arithmetic alone, no data movement: 0.218 ms = 11.2 % of HBM on this shape, profiles/r02_ntt_floor.md); traffic = dram read + write of one "
                             "launch from profiles/r02_ncu_lde_block_v2.md (output partly still in L2)",
                     "from_values_gbs": 80 * n * nw / ((stages["wires_intt"] + lde_ms) * 1e-3) / 1e9,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": lde_bytes, "avg_ms": lde_ms},
        "poseidon": {"kernel": "wires Merkle commit (merkle_leaves_kernel + 7 level launches + merkle_cap_subtree_kernel, levels replayed as a CUDA graph)", "perms_per_launch": perms,
                     "avg_ms": pos_ms, "perms_per_sec": perms / (pos_ms * 1e-3), "bound": "integer pipes (fmaheavy/IMAD + alu)",
                     "imad_peak_per_s": 148 * 64 * 1.965e9,
                     "frac_of_imad_peak_algorithmic": perms / (pos_ms * 1e-3) * 6600 / (148 * 64 * 1.965e9),
                     "ncu_fmaheavy_pipe_cycles_active_pct": 75.2, "ncu_alu_pipe_pct": 68.3, "ncu_issue_slots_busy_pct": 69.3,
                     "ncu_thread_instructions_per_permutation": 23673,
                     "note": "6.6 k IMAD issue slots per permutation (SURVEY 8d) is the ALGORITHMIC fraction; the ncu figures are the leaf "
                             "kernel's measured pipe utilisation (profiles/r02_ncu_merkle_leaves.md): both integer pipes ~70 % busy"},
        "clocks": clk.summary(),
    }
    # the DOMINANT kernel of a proof is the Poseidon tree hashing (integer-pipe bound); the contract's `roofline` object
    # describes the HBM-side kernel the north star names (LDE-NTT) and carries the Poseidon block inside it so that both survive
    line["roofline"]["dominant_kernel"] = "poseidon (see roofline.poseidon): ~65 % of a proof's GPU time; the LDE launch is ~5 %"
    line["roofline"]["poseidon"] = line["poseidon"]
    if not args.no_sweep:
        # config #3 points: fused from_values commit of synthetic columns, inputs resident in HBM
        sweep = []
        grid = ((14, 135), (14, 200), (14, 400), (16, 135), (16, 200), (18, 100), (20, 100), (22, 100))
        if args.full_grid:
            grid = tuple((lg, c) for lg in (14, 16, 18, 20, 22) for c in (100, 135, 200, 400) if not (lg == 22 and c > 135))
        for lg_n, cols in grid:
            nn = 1 << lg_n
            vals = synth_columns(np, lg_n, cols)
            cap, tm = Z.commit_batch(vals, 3, 4, reps=3, device=local_rank)
            gbs = 80 * nn * cols / (tm["lde_ms"] * 1e-3) / 1e9
            pp = 8 * nn * ((cols + 7) // 8) + 8 * nn - 16
            sweep.append({"lg_n": lg_n, "cols": cols, "lde_ms": tm["lde_ms"], "lde_gbs": gbs, "lde_frac_hbm": gbs / peak,
                          "merkle_ms": tm["merkle_ms"], "perms_per_sec": pp / (tm["merkle_ms"] * 1e-3),
                          "cap_equals_oracle_golden": cap_matches(np, cap, lg_n, cols)})
        line["lde_merkle_sweep"] = sweep
    if world == 1 and not args.no_aggregation:
        # config #4: the aggregator's default tree (branching 2, depth 3: 8 leaf proofs -> 4 + 2 + 1 chunk proofs,
        # aggregator/src/circuits/tree.rs:17-20,55-77) with recursion-shaped chunk circuits (14-gate set, 4 selector groups,
        # zero-knowledge like the aggregator's own config — aggregator.rs:21 — so n = 2^14 with the blinding rows; the real
        # recursive-verifier circuit needs the Rust circuit builder). Chunks of a level run concurrently on their own prover
        # contexts, levels are sequential. Witness generation (Rust side) is not included.
        rs = Z.SynthCircuit(zk=True, seed=4, **Z.SynthCircuit.RECURSION)
        S = min(4, B)
        rcircs = [Z.ProverCircuit(rs.common, rs.const_sigma_values, is_values=True, device=local_rank) for _ in range(S)]
        rpinned = torch.empty(rs.wires.shape, dtype=torch.int64, pin_memory=True)
        rpinned.numpy().view(np.uint64)[:] = rs.wires
        raddr = rpinned.data_ptr()

        def prove_chunk(prover, chunk, level, index):
            return prover.prove(raddr, rs.public_inputs, salt_seed=1000 * level + index)

        leaves = [None] * 8
        for _ in range(3):
            zbatch.aggregate_tree(leaves, 2, prove_chunk, rcircs)
        reps = max(3, K // 2)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            root, widths = zbatch.aggregate_tree(leaves, 2, prove_chunk, rcircs)
        torch.cuda.synchronize()
        t_tree = (time.perf_counter() - t0) / reps
        t0 = time.perf_counter()
        for _ in range(reps):
            rcircs[0].prove(raddr, rs.public_inputs, salt_seed=0)
        t_chunk = (time.perf_counter() - t0) / reps
        line["aggregation"] = {"workload": "aggregation_tree_8_leaves_recursion_zk_synth_n2^14", "leaf_proofs": 8, "branching": 2,
                               "chunk_proofs_per_level": widths, "tree_ms": 1000 * t_tree, "chunk_prove_ms": 1000 * t_chunk,
                               "chunk_stage_ms": rcircs[0].timings(), "chunk_proof_bytes": len(root), "streams": S,
                               "chunk_degree_bits": int(rs.n).bit_length() - 1, "gates": 14, "zero_knowledge": True,
                               "note": "host-buffer zkb_prove() calls (H2D of the 135 x 2^12 wire matrix inside); synthetic "
                                       "recursion-shaped circuit, recursion gate formulas unpinned against qp-plonky2 (DESIGN.md)"}
    if world == 1 and not args.no_aggregation:
        # config #2: voting-shaped circuit (6-gate set, n = 2^9 as SURVEY.md §8d states it, non-zk, 13 public inputs —
        # voting/src/lib.rs:72-76): latency of one proof
        vs = Z.SynthCircuit(zk=False, seed=2, min_degree_bits=9, **Z.VOTING)
        vc = Z.ProverCircuit(vs.common, vs.const_sigma_values, is_values=True, device=local_rank)
        for i in range(3):
            vp = vc.prove(vs.wires, vs.public_inputs, salt_seed=i)
        t0 = time.perf_counter()
        for i in range(K):
            vp = vc.prove(vs.wires, vs.public_inputs, salt_seed=i)
        line["voting"] = {"workload": "voting_synth", "degree_bits": int(vs.n).bit_length() - 1, "prove_ms": 1000 * (time.perf_counter() - t0) / K,
                          "proof_bytes": len(vp), "stage_ms": vc.timings()}
        vref = vc.prove(vs.wires, vs.public_inputs, salt_seed=100 + K - 1)
        vc.close()
        # the same circuit through the proof engine: 16 proofs in flight (a proof of 4096 LDE points is latency-bound alone),
        # host buffers, the wire matrix copied into a pinned slot and uploaded inside the timed region
        VB = 16
        veng = Z.Engine(vs.common, vs.const_sigma_values, is_values=True, device=local_rank, contexts=VB, slots=2 * VB,
                        num_wires=vs.wires.shape[0])
        vouts = [np.zeros(veng.proof_size, dtype=np.uint8) for _ in range(2 * VB)]

        def voting_step(seed):
            slots = []
            for b in range(VB):
                slot, buf = veng.acquire()
                buf[:] = vs.wires
                veng.submit(slot, vs.public_inputs, salt_seed=seed + b, out=vouts[slot])
                slots.append(slot)
            return [veng.wait(sl) for sl in slots]

        for i in range(3):
            voting_step(i)
        t0 = time.perf_counter()
        for i in range(K):
            got = voting_step(100 + i)
        dt = time.perf_counter() - t0
        line["voting"]["engine_proofs_per_sec"] = VB * K / dt
        line["voting"]["engine_proofs_in_flight"] = VB
        if bytes(got[0]) != bytes(vref):
            raise SystemExit("bench.py: the engine's voting proof differs from the single-context proof of the same inputs")
        line["voting"]["engine_proof_equals_single_context_proof"] = True
        veng.close()
    if sharded_line is not None:
        line["sharded_commit"] = sharded_line
    if forest_line is not None:
        line["aggregation_forest"] = forest_line
    if world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle as O   # CPU baseline leg only

        O.build()
        O.set_num_threads(os.cpu_count() or 1)
        O.set_fast(True)                       # the port's Poseidon in its AVX2 form, as in `--impl reference`
        os_ = O.Synth(zk=True, seed=1, **O.Synth.WORMHOLE)
        oc = O.Circuit(os_.common, os_.const_sigma_values)
        t0 = time.perf_counter()
        ref = oc.prove(os_.wires, os_.public_inputs, salt_seed=3000 + K - 1)   # same seed as stream 0's last e2e proof
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": 1.0 / dt, "unit": UNIT, "cores": O.num_threads(), "kind": "port",
                                "sample": "1 full proof of the bench circuit, oracle's restated Plonky2 prover (AVX2 Poseidon, 4-lane AVX2 gate evaluation, OpenMP)",
                                "bytes_identical_to_gpu_proof": bool(ref == proof)}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

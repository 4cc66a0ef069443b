import os, sys, hashlib
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "zk-circuits_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, zkb200 as Z, oracle as O
O.build()
def H(a): return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:12]
step = sys.argv[1] if len(sys.argv) > 1 else "all"
s0 = Z.SynthCircuit(zk=True, seed=21, **Z.TINY)
so = O.Synth(zk=True, seed=21, **O.Synth.TINY)
print("before: wires", H(s0.wires), H(so.wires), "cs", H(s0.const_sigma_values), H(so.const_sigma_values))
comm = Z.Comm(Z.comm_unique_id(), 1, 0)
print("comm created")
s1 = Z.SynthCircuit(zk=True, seed=21, **Z.TINY)
print("after create: wires", H(s1.wires), "s0 still", H(s0.wires))
rng = np.random.default_rng(4)
if step in ("all", "commit"):
    vals = rng.integers(0, O.P, size=(37, 1 << 12), dtype=np.uint64)
    cap, tm = comm.commit(vals, 3, 4, reps=2)
    s2 = Z.SynthCircuit(zk=True, seed=21, **Z.TINY)
    print("after commit: wires", H(s2.wires), "s0 still", H(s0.wires), "oracle again", H(O.Synth(zk=True, seed=21, **O.Synth.TINY).wires))
if step in ("all", "quot"):
    n = 1 << 7
    q = rng.integers(0, O.P, size=(2, 8 * n), dtype=np.uint64)
    got, _ = comm.quotient_chunks(q, n, 3)
    s3 = Z.SynthCircuit(zk=True, seed=21, **Z.TINY)
    print("after quotient_chunks: wires", H(s3.wires), "s0 still", H(s0.wires))
comm.close()
s4 = Z.SynthCircuit(zk=True, seed=21, **Z.TINY)
print("after close: wires", H(s4.wires), "cs", H(s4.const_sigma_values), "pis", H(s4.public_inputs), H(so.public_inputs))
if H(s4.wires) != H(so.wires):
    d = np.argwhere(s4.wires != so.wires)
    print("diff count", len(d), "first", d[:5].tolist(), "cols", sorted(set(d[:, 0].tolist()))[:20], "rows range", d[:, 1].min(), d[:, 1].max())
print("---- now import torch (zkb200.batch) and repeat the batch test's steps")
from zkb200 import batch
s = Z.SynthCircuit(zk=True, seed=21, **Z.TINY)
print("after torch import: wires", H(s.wires), "cs", H(s.const_sigma_values), "oracle", H(O.Synth(zk=True, seed=21, **O.Synth.TINY).wires))
oc = O.Circuit(s.common, s.const_sigma_values)
provers = [Z.ProverCircuit(s.common, s.const_sigma_values, is_values=True) for _ in range(2)]
seeds = [7, 8, 9, 10, 11]
proofs = batch.prove_batch(seeds, provers, lambda p, w, i: p.prove(s.wires, s.public_inputs, salt_seed=w))
for seed, proof in zip(seeds, proofs):
    want = oc.prove(s.wires, s.public_inputs, salt_seed=seed)
    want2 = oc.prove(s.wires, s.public_inputs, salt_seed=seed)
    print(seed, "gpu==oracle", proof == want, "oracle deterministic", want == want2, "verify gpu", repr(oc.verify(proof)), "verify oracle", repr(oc.verify(want)),
          "serial gpu == oracle", provers[0].prove(s.wires, s.public_inputs, salt_seed=seed) == want, "threads", O.num_threads())

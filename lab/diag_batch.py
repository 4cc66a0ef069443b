"""Diagnostic for test_batch_of_proofs_on_two_streams: where does a verifier rejection come from?"""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "zk-circuits_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, zkb200 as Z, oracle as O
from zkb200 import batch
O.build()
s = Z.SynthCircuit(zk=True, seed=21, **Z.TINY)
so = O.Synth(zk=True, seed=21, **O.Synth.TINY)
print("synth equal:", s.common == so.common, np.array_equal(s.wires, so.wires), np.array_equal(s.const_sigma_values, so.const_sigma_values),
      np.array_equal(s.public_inputs, so.public_inputs), "check:", repr(so.check()))
oc = O.Circuit(s.common, s.const_sigma_values)
provers = [Z.ProverCircuit(s.common, s.const_sigma_values, is_values=True) for _ in range(2)]
bad = 0
for rep in range(6):
    seeds = [7, 8, 9, 10, 11]
    proofs = batch.prove_batch(seeds, provers, lambda p, w, i: p.prove(s.wires, s.public_inputs, salt_seed=w))
    for seed, proof in zip(seeds, proofs):
        want = oc.prove(s.wires, s.public_inputs, salt_seed=seed)
        v1, v2 = oc.verify(proof), oc.verify(want)
        again = provers[0].prove(s.wires, s.public_inputs, salt_seed=seed)
        if proof != want or v1 or v2 or again != want:
            bad += 1
            diff = [i for i in range(min(len(proof), len(want))) if proof[i] != want[i]]
            print(f"rep {rep} seed {seed}: gpu==oracle {proof == want} verify(gpu)={v1!r} verify(oracle)={v2!r} serial-gpu==oracle {again == want} "
                  f"first diff byte {diff[:3]} ndiff {len(diff)}")
print("bad:", bad)

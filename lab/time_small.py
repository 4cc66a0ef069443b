"""Latency of small-circuit proofs (recursion chunk n = 2^12, voting n = 2^8): per-stage device times, 20 proofs each."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "zk-circuits_b200"))
import time, zkb200 as Z
for name, s in (("recursion", Z.SynthCircuit(seed=4, **Z.SynthCircuit.RECURSION)), ("voting", Z.SynthCircuit(seed=2, **Z.VOTING))):
    c = Z.ProverCircuit(s.common, s.const_sigma_values, is_values=True)
    for i in range(3):
        c.prove(s.wires, s.public_inputs, salt_seed=0)
    acc = {}
    t0 = time.perf_counter()
    for i in range(20):
        c.prove(s.wires, s.public_inputs, salt_seed=0)
        for k, v in c.timings().items():
            acc[k] = acc.get(k, 0) + v / 20
    print(name, "ms/proof %.3f" % ((time.perf_counter() - t0) * 50), {k: round(v, 3) for k, v in acc.items()})

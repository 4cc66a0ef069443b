"""ncu target: a few proofs of the recursion-shaped chunk circuit (n = 2^12, 14 gates)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "zk-circuits_b200"))
import zkb200 as Z
s = Z.SynthCircuit(seed=4, **Z.SynthCircuit.RECURSION)
c = Z.ProverCircuit(s.common, s.const_sigma_values, is_values=True)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    p = c.prove(s.wires, s.public_inputs, salt_seed=i)
print("ok", len(p), c.timings())

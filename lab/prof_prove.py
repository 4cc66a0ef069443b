import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "zk-circuits_b200"))
import zkb200 as Z
s = Z.SynthCircuit(zk=True, seed=1, **Z.WORMHOLE)
c = Z.ProverCircuit(s.common, s.const_sigma_values, is_values=True)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    p = c.prove(s.wires, s.public_inputs, salt_seed=i)
print("ok", len(p), c.timings())

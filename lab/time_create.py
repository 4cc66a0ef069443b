"""Where zkb_circuit_create spends its time (ZKB_TRACE=1 checkpoints), wormhole-shaped zk circuit, 4 contexts in a row."""
import sys, os, time
os.environ["ZKB_TRACE"] = "1"
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "zk-circuits_b200"))
import zkb200 as Z
s = Z.SynthCircuit(zk=True, seed=1, **Z.WORMHOLE)
keep = []
for i in range(10):
    t0 = time.perf_counter()
    keep.append(Z.ProverCircuit(s.common, s.const_sigma_values, is_values=True))
    print("create %d: %.2f ms" % (i, 1e3 * (time.perf_counter() - t0)), flush=True)
t0 = time.perf_counter(); keep.clear(); print("destroy all: %.2f ms" % (1e3 * (time.perf_counter() - t0)))

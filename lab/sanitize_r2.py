"""compute-sanitizer target for the round-2 paths: a tiny zk proof with CSPRNG salts + witness check, a voting-size proof (cap
subtree over 8 levels), the engine, a Merkle commit that uses the cap-subtree kernel, the single-rank NCCL entry points."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "zk-circuits_b200"))
import numpy as np, zkb200 as Z
s = Z.SynthCircuit(zk=False, seed=3, n_poseidon=5, n_base_sum=3, n_arith=4, n_const=3, num_public_inputs=5, n_arith_ext=4, n_mul_ext=3,
                   n_reducing=3, n_reducing_ext=3, n_random_access=3, n_exp=3, n_coset=3, n_mds=2)
c = Z.ProverCircuit(s.common, s.const_sigma_values, is_values=True)
p = c.prove(s.wires, s.public_inputs, check_witness=True)
print("recursion tiny ok", len(p))
v = Z.SynthCircuit(zk=False, seed=2, min_degree_bits=9, **Z.VOTING)
vc = Z.ProverCircuit(v.common, v.const_sigma_values, is_values=True)
print("voting ok", len(vc.prove(v.wires, v.public_inputs, check_witness=True)))
rng = np.random.default_rng(0)
vals = rng.integers(0, 0xFFFFFFFF00000001, size=(9, 1 << 10), dtype=np.uint64)
cap, _ = Z.commit_batch(vals, 3, 4)
print("commit ok", int(cap[0, 0]) != 0)
if "--zk" in sys.argv:
    t = Z.SynthCircuit(zk=True, seed=11, **Z.TINY)
    eng = Z.Engine(t.common, t.const_sigma_values, is_values=True, contexts=2, slots=3)
    slots = []
    for i in range(3):
        slot, buf = eng.acquire()
        buf[:] = t.wires
        eng.submit(slot, t.public_inputs, check_witness=True)     # CSPRNG salts
        slots.append(slot)
    print("engine ok", [len(eng.wait(sl)) for sl in slots])
    eng.close()
if "--nccl" in sys.argv:
    comm = Z.Comm(Z.comm_unique_id(), 1, 0)
    cap2, _ = comm.commit(vals, 3, 4)
    q = rng.integers(0, 0xFFFFFFFF00000001, size=(2, 8 * 64), dtype=np.uint64)
    out, _ = comm.quotient_chunks(q, 64, 3)
    print("nccl ok", np.array_equal(cap, cap2), out.shape)
    comm.close()

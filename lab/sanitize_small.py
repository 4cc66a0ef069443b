"""Small end-to-end run for compute-sanitizer (memcheck): tiny zk proof, a voting-shaped proof, LDE on both NTT paths,
Merkle commit with narrow and wide leaves, coset-sharded commit part."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "zk-circuits_b200"))
import numpy as np, zkb200 as Z
rng = np.random.default_rng(0)
P = 0xFFFFFFFF00000001
for spec, zk in ((Z.TINY, True), (Z.VOTING, False)):
    s = Z.SynthCircuit(zk=zk, seed=3, **spec)
    c = Z.ProverCircuit(s.common, s.const_sigma_values, is_values=True)
    p = c.prove(s.wires, s.public_inputs, salt_seed=1)
    print("proof", len(p))
# recursion gate set (third quotient launch, sliced + concurrent for small circuits)
s = Z.SynthCircuit(seed=3, n_poseidon=5, n_base_sum=3, n_arith=4, n_const=3, num_public_inputs=5, n_arith_ext=4, n_mul_ext=3,
                   n_reducing=3, n_reducing_ext=3, n_random_access=3, n_exp=3, n_coset=3, n_mds=2)
c = Z.ProverCircuit(s.common, s.const_sigma_values, is_values=True)
print("recursion proof", len(c.prove(s.wires, s.public_inputs)))
for lg, cols in ((9, 5), (15, 2)):
    v = rng.integers(0, P, size=(cols, 1 << lg), dtype=np.uint64)
    Z.lde_batch(v, 3)
    Z.commit_batch(v, 3, 4)
    Z.commit_cosets(v, 3, 4, 2, 4)
Z.merkle_commit(rng.integers(0, P, size=(3, 64), dtype=np.uint64), 2)
Z.merkle_commit(rng.integers(0, P, size=(139, 512), dtype=np.uint64), 4)
print("done")

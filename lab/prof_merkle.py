import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "zk-circuits_b200"))
import numpy as np, zkb200 as Z
rng = np.random.default_rng(0)
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 14
vals = rng.integers(0, 0xFFFFFFFF00000001, size=(135, 1 << lg), dtype=np.uint64)
cap, t = Z.commit_batch(vals, 3, 4, reps=2)
print("ok", t)

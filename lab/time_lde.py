"""LDE timing at large n: python lab/time_lde.py lg:cols [lg:cols ...] (commit_batch, lde_ms = iNTT + 8 coset transforms)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "zk-circuits_b200"))
import numpy as np, zkb200 as Z
rng = np.random.default_rng(0)
reps = int(os.environ.get("REPS", "3"))
for spec in sys.argv[1:]:
    lg, cols = (int(x) for x in spec.split(":"))
    vals = rng.integers(0, 0xFFFFFFFF00000001, size=(cols, 1 << lg), dtype=np.uint64)
    cap, t = Z.commit_batch(vals, 3, 4, reps=reps)
    n = 1 << lg
    print(lg, cols, {k: round(v, 3) for k, v in t.items()}, "lde GB/s %.1f" % (80 * n * cols / (t["lde_ms"] * 1e-3) / 1e9), flush=True)

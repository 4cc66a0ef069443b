"""Leaf-hash throughput (config #3 point: 2^17 leaves x 135 columns, 5 repetitions)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "zk-circuits_b200"))
import numpy as np, zkb200 as Z
rng = np.random.default_rng(0)
for lg, cols in ((14, 135), (17, 100)):
    vals = rng.integers(0, 0xFFFFFFFF00000001, size=(cols, 1 << lg), dtype=np.uint64)
    cap, t = Z.commit_batch(vals, 3, 4, reps=5)
    n = 1 << lg
    perms = 8 * n * ((cols + 7) // 8) + 8 * n - 16
    print(lg, cols, t, "G perms/s %.4f" % (perms / (t["merkle_ms"] * 1e-3) / 1e9), "lde GB/s %.1f" % (80 * n * cols / (t["lde_ms"] * 1e-3) / 1e9))

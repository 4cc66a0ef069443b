#!/usr/bin/env python3
"""Golden Merkle caps of `PolynomialBatch::from_values` at the microbenchmark sizes (BASELINE.json config #3):
the 16 cap digests of the commitment to x[c][i] = SplitMix64-finalizer(SEED + c n + i) mod p at rate_bits 3, cap_height 4,
computed ONCE with the CPU oracle (OpenMP) and stored in tests/golden/config3_caps.json. The GPU tests and bench.py's sweep /
sharded-commit legs compare against them (the oracle itself would need minutes per shape at these sizes).

    python tests/golden/make_caps.py            # regenerates every shape (about 10 minutes on 8 cores, ~20 GB of RAM)
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

SEED = 0xB200000000000001
P = 0xFFFFFFFF00000001
SHAPES = [(14, 100), (14, 135), (14, 200), (14, 400), (16, 100), (16, 135), (16, 200), (18, 100), (18, 135), (20, 100),
          (19, 5), (21, 5), (22, 5)]      # narrow batches at the sizes of the staged first NTT step and of the three-step transform


def synth_columns(lg_n, cols):
    nn = 1 << lg_n
    idx = np.arange(nn * cols, dtype=np.uint64) + np.uint64(SEED)
    with np.errstate(over="ignore"):
        z = idx
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return np.where(z >= np.uint64(P), z - np.uint64(P), z).reshape(cols, nn)


def main():
    import oracle as O

    O.build()
    O.set_num_threads(os.cpu_count() or 1)
    path = os.path.join(HERE, "config3_caps.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    for lg_n, cols in SHAPES:
        key = f"{lg_n}x{cols}"
        if key in out and "--force" not in sys.argv:
            continue
        t0 = time.time()
        vals = synth_columns(lg_n, cols)
        _, lde = O.lde_batch(vals, 3)
        _, cap = O.merkle_commit(lde, 4)
        del lde
        out[key] = {"lg_n": lg_n, "cols": cols, "rate_bits": 3, "cap_height": 4, "cap": [[f"{int(x):016x}" for x in d] for d in cap]}
        print(key, f"{time.time() - t0:.1f}s", out[key]["cap"][0], flush=True)
        with open(path, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()

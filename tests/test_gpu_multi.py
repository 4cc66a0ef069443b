"""Multi-GPU exactness of the in-library NCCL sharding (zkb_comm_*): one process per GPU, 2 (or 4) ranks on one box. Skipped on a
single-GPU box (the driver's GPU tier); run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`.
The same exchange logic is covered on the CPU over gloo (tests/test_batch_gloo.py) and with one rank in test_gpu_new_paths.py."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, time
import numpy as np
root, rank, world, tmp = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
sys.path.insert(0, os.path.join(root, "zk-circuits_b200")); sys.path.insert(0, os.path.join(root, "oracle"))
import zkb200 as Z, oracle as O
idf = os.path.join(tmp, "nccl_id.bin")
if rank == 0:
    uid = Z.comm_unique_id()
    uid.tofile(idf + ".tmp"); os.replace(idf + ".tmp", idf)
else:
    for _ in range(600):
        if os.path.exists(idf): break
        time.sleep(0.05)
    uid = np.fromfile(idf, dtype=np.uint8)
comm = Z.Comm(uid, world, rank, device=rank)
rng = np.random.default_rng(4)                      # same data on every rank
P = O.P
# 1. sharded commitment: odd column count (padding of the column slices), both NTT paths
# (the second shape makes the peer windows grow and be re-mapped, the third re-uses them, reps = 2 overwrites a window in place)
modes = []
for lg_n, cols, reps in ((10, 37, 1), (15, 5, 1), (12, 9, 2)):
    vals = rng.integers(0, P, size=(cols, 1 << lg_n), dtype=np.uint64)
    cap, tm = comm.commit(vals, 3, 4, reps=reps)
    _, lde = O.lde_batch(vals, 3)
    _, want = O.merkle_commit(lde, 4)
    assert np.array_equal(cap, want), (lg_n, cols)
    modes.append(tm["peer_windows"])
assert len(set(modes)) == 1
if os.environ.get("ZKB_SHARDED_P2P") == "0":
    assert not modes[0]
# 2. quotient chunks from coset-local evaluations
n, R = 1 << 8, 8
chunks = rng.integers(0, P, size=(2, R, n), dtype=np.uint64)
B, sl = R // world, n // world
w_N = O.root_of_unity(8 + 3)
GEN = 0xC65C18B67785D900
q = np.zeros((2, B * n), dtype=np.uint64)
for ch in range(2):
    _, lde = O.lde_batch(chunks[ch], 3, from_coeffs=True)
    for i in range(B):
        jb = rank * B + i
        j = int(format(jb, "03b")[::-1], 2)
        c = pow(GEN * pow(w_N, j, P) % P, n, P)
        acc = [0] * n
        for m in range(R):
            cm = pow(c, m, P)
            acc = [(a + cm * int(v)) % P for a, v in zip(acc, lde[m, jb * n:(jb + 1) * n])]
        q[ch, i * n:(i + 1) * n] = acc
got, _ = comm.quotient_chunks(q, n, 3)
assert np.array_equal(got, chunks[:, :, rank * sl:(rank + 1) * sl])
comm.close()
open(os.path.join(tmp, f"ok{rank}"), "w").write("ok peer_windows=%s" % modes[0])
'''


@pytest.mark.parametrize("p2p", ["1", "0"])
@pytest.mark.parametrize("world", [2, 4])
def test_nccl_sharded_commit_and_quotient_exchange_across_gpus(tmp_path, world, p2p):
    """p2p = 1: the LDE kernels read the coefficients out of the peers' CUDA-IPC windows (falls back by itself where the mapping
    is unavailable; the mode used is printed); p2p = 0: the NCCL all-gather form of the same exchange."""
    sys.path.insert(0, os.path.join(ROOT, "zk-circuits_b200"))
    import zkb200

    if zkb200.device_count() < world:
        pytest.skip(f"needs {world} GPUs on one box")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, ZKB_SHARDED_P2P=p2p)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(r), str(world), str(tmp_path)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True, env=env) for r in range(world)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r}:\n{o[-3000:]}"
        assert (tmp_path / f"ok{r}").read_text().startswith("ok")
    print(f"world {world} ZKB_SHARDED_P2P={p2p}: {(tmp_path / 'ok0').read_text()}")

"""Coset-sharded commitment of one large polynomial batch across the GPUs of a node.

In the reference's leaf order (`reverse_index_bits` on the LDE index) coset j of the blow-up is the CONTIGUOUS leaf block
bitrev(j), and with cap_height >= rate_bits every block holds whole cap subtrees. So rank r of G takes the 2^rate_bits / G
consecutive leaf blocks [r * B, (r + 1) * B): it interpolates every column (cheap, redundant), extends and hashes only its
blocks, builds their subtrees, and the only exchange of the whole commitment is ONE all-gather of 2^cap_height digests
(16 x 32 bytes for the reference configuration) — NCCL on the GPU box, gloo in the CPU test.
"""
import numpy as np

from . import batch


def block_range(rate_bits, rank, world_size):
    nblk = 1 << rate_bits
    if world_size > nblk or nblk % world_size:
        raise ValueError(f"world size {world_size} must divide the {nblk} coset blocks")
    per = nblk // world_size
    return rank * per, (rank + 1) * per


def sharded_commit(values, rate_bits=3, cap_height=4, device=0, reps=1, commit_fn=None, gather_device=None):
    """Every rank passes the same `values` [ncols][n]; returns (cap [2^cap_height][4], timings of this rank).

    commit_fn(values, rate_bits, cap_height, blk_lo, blk_hi) -> (cap_part, timings) defaults to the CUDA entry point
    zkb_commit_cosets on `device`; the CPU test injects an oracle-backed stand-in to exercise the sharding and the gather."""
    rank, ws = batch.world()
    lo, hi = block_range(rate_bits, rank, ws)
    if commit_fn is None:
        from . import commit_cosets

        def commit_fn(v, rb, ch, a, b):
            return commit_cosets(v, rb, ch, a, b, reps=reps, device=device)

    part, timings = commit_fn(values, rate_bits, cap_height, lo, hi)
    part = np.ascontiguousarray(part, dtype=np.uint64)
    if ws == 1:
        return part, timings
    import torch
    import torch.distributed as dist

    dev = gather_device or "cpu"
    mine = torch.from_numpy(part.view(np.int64)).to(dev)
    out = torch.empty((ws * mine.shape[0], mine.shape[1]), dtype=torch.int64, device=dev)   # concatenated along dim 0
    dist.all_gather_into_tensor(out, mine)
    cap = out.cpu().numpy().view(np.uint64).reshape(-1, 4)
    return cap, timings

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


# ---------------------------------------------------------------------------------------------------------------------------
# Quotient chunks from coset-local evaluations: the host-side mirror of zkb_quotient_chunks_sharded (csrc/sharded.cpp) — same
# block ownership, same slicing of the coefficient index range, same R x R solve — over torch.distributed, so that the exchange
# logic is covered by a gloo world-size-2 CPU test. On the GPU box the C entry point does all of this with NCCL inside the library.
# ---------------------------------------------------------------------------------------------------------------------------
P = 0xFFFFFFFF00000001
GEN = 0xC65C18B67785D900


def _bitrev(x, bits):
    return int(format(x, f"0{bits}b")[::-1], 2) if bits else 0


def vandermonde_inverse(n, rate_bits, root_of_unity):
    """V^-1[m][j] = c_0^-m w_R^(-j m) / R for c_j = (g w_N^j)^n = c_0 w_R^j (chunk m of t sees the factor c_j^m on coset j)."""
    R = 1 << rate_bits
    w_N = root_of_unity(n.bit_length() - 1 + rate_bits)
    c0_inv, wr_inv, r_inv = pow(pow(GEN, n, P), P - 2, P), pow(pow(w_N, n, P), P - 2, P), pow(R, P - 2, P)
    return [[pow(c0_inv, m, P) * pow(wr_inv, j * m, P) % P * r_inv % P for j in range(R)] for m in range(R)]


def quotient_chunks(q_local, n, rate_bits, coset_interpolate, root_of_unity):
    """q_local [nch][B n]: t on this rank's leaf blocks (leaf order). coset_interpolate(block_values, shift) -> the n
    coefficients of the degree < n interpolant on shift * <w_n> from values in bit-reversed order. Returns [nch][R][n / G]:
    this rank's coefficient slice of every chunk. ONE all_to_all."""
    import torch
    import torch.distributed as dist

    rank, G = batch.world()
    R = 1 << rate_bits
    B, sl = R // G, n // G
    q_local = np.ascontiguousarray(q_local, dtype=np.uint64)
    nch = q_local.shape[0]
    w_N = root_of_unity(n.bit_length() - 1 + rate_bits)
    u = np.zeros((nch, B, n), dtype=np.uint64)
    for ch in range(nch):
        for i in range(B):
            j = _bitrev(rank * B + i, rate_bits)
            u[ch, i] = coset_interpolate(q_local[ch, i * n:(i + 1) * n], GEN * pow(w_N, j, P) % P)
    send = torch.from_numpy(np.stack([u[:, :, p * sl:(p + 1) * sl] for p in range(G)]).view(np.int64))     # [G][nch][B][sl]
    recv = torch.empty_like(send)
    if G == 1:
        recv.copy_(send)
    else:
        dist.all_to_all_single(recv, send)                       # equal splits along dim 0: block p goes to / comes from rank p
    allu = np.zeros((nch, R, sl), dtype=np.uint64)               # by coset index
    for p in range(G):
        part = recv[p].numpy().view(np.uint64)
        for i in range(B):
            allu[:, _bitrev(p * B + i, rate_bits), :] = part[:, i, :]
    vinv = vandermonde_inverse(n, rate_bits, root_of_unity)
    out = np.zeros((nch, R, sl), dtype=np.uint64)
    for ch in range(nch):
        cols = [[int(x) for x in allu[ch, j]] for j in range(R)]
        for m in range(R):
            out[ch, m] = [sum(vinv[m][j] * cols[j][k] for j in range(R)) % P for k in range(sl)]
    return out

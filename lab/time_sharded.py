"""Sharded commit timing only (torchrun): python -m torch.distributed.run --nproc-per-node N lab/time_sharded.py [lg_n cols]
ZKB_SHARDED_P2P=0 selects the NCCL gather."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "zk-circuits_b200"))
import numpy as np, torch, torch.distributed as dist
import zkb200 as Z
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
lg_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 100
uid = torch.from_numpy(Z.comm_unique_id() if rank == 0 else np.zeros(128, dtype=np.uint8)).cuda()
dist.broadcast(uid, 0)
comm = Z.Comm(uid.cpu().numpy(), world, rank, device=local)
rng = np.random.default_rng(1)
vals = rng.integers(0, 0xFFFFFFFF00000001, size=(cols, 1 << lg_n), dtype=np.uint64)
for it in range(2):
    dist.barrier()
    cap, tm = comm.commit(vals, 3, 4, reps=3)
    t = torch.tensor([tm["lde_ms"], tm["merkle_ms"], tm["gather_ms"]], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"ranks {world} 2^{lg_n} x {cols} peer_windows={tm['peer_windows']} lde {t[0]:.3f} ms merkle {t[1]:.3f} ms exchange {t[2]:.3f} ms cap0 {int(cap[0,0]):#x}", flush=True)
qn = 1 << 20
qv = rng.integers(0, 0xFFFFFFFF00000001, size=(2, (8 // world) * qn), dtype=np.uint64)
for it in range(4):
    dist.barrier()
    _, qt = comm.quotient_chunks(qv, qn, 3)
    t = torch.tensor([qt["interpolate_ms"], qt["exchange_ms"]], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"ranks {world} quotient chunks n=2^20 x 2: coset iNTT {t[0]:.3f} ms, exchange + solve {t[1]:.3f} ms", flush=True)
comm.close()
dist.destroy_process_group()

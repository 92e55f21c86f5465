"""Prints, per call and rank, the triangle row block skd_comm_self_dists_knn used and the times it measured
(adaptive split diagnostic).   torchrun --nproc-per-node N tools/split_probe.py [n] [calls]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import sketchlib.rust_b200 as sk  # noqa: E402
from sketchlib.rust_b200 import synth  # noqa: E402

KMERS = list(range(15, 31, 2))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = sk.Context(local)
grp = sk.PeerGroup(ctx, rank, world, sk.PeerGroup.window_bytes_knn(n, 50, world))
if world > 1:
    hs = [None] * world
    dist.all_gather_object(hs, grp.export())
    grp.connect(hs)
per = len(KMERS) * 16 * 14
bins = torch.empty(n * per, dtype=torch.int64, device="cuda")
for s0 in range(0, n, 50000):
    s1 = min(n, s0 + 50000)
    bins[s0 * per:s1 * per] = synth.make_bins(s1 - s0, 16, KMERS, seed=3 + s0, device="cuda")
torch.cuda.synchronize()
g = sk.SketchDb.upload(ctx, bins.data_ptr(), n, KMERS, 16, bins_loc=sk.SKD_DEVICE)
del bins
out = (np.zeros((n, 50), np.uint64), np.zeros((n, 50), np.float32), np.zeros((n, 50), np.float32)) if rank == 0 else (None, None, None)
for c in range(calls):
    if world > 1:
        dist.barrier()
    grp.self_dists_knn(g, 50, sk.set_k(g), root=0, out=out)
    a, b = grp.last_row_block
    print(f"call {c} rank {rank}: rows [{a}, {b}) pairs {int(ctx.lib.skd_self_pairs_in_rows(n, a, b))} op {grp.last_op_ms:.1f} ms "
          f"exchange {grp.last_exchange_ms:.1f} ms local {grp.last_op_ms - grp.last_exchange_ms:.1f} ms checksum "
          f"{float(out[1].astype(np.float64).sum()) if rank == 0 else 0:.6f}", flush=True)
grp.close()

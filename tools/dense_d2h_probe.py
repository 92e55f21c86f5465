"""How should dense results leave the GPU?  A/B of the three ways on the bench's dense leg (n = 10 000 sqrt(N) genomes,
every rank its equal-pair block of triangle rows, database resident in HBM, ~400 MB of f32 pairs per rank per step):
  zerocopy  the kernel stores straight into mapped pinned host memory (library default)
  staged    SKD_ZEROCOPY=0: row chunks in HBM, copy-engine D2H of chunk c overlapped with the kernel of chunk c + 1
  bulk      results left in HBM, then ONE cudaMemcpy D2H
plus the ceiling: a bare D2H copy of the same bytes with nothing else running (all ranks at once).
    python tools/dense_d2h_probe.py            torchrun --nproc-per-node N tools/dense_d2h_probe.py
Prints one JSON line (rank 0): per mode the max-over-ranks ms per step and the aggregate host-ingest GB/s."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import sketchlib.rust_b200 as sk  # noqa: E402
from sketchlib.rust_b200 import synth  # noqa: E402

KMERS = list(range(15, 31, 2))


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n = int(round(10000 * world ** 0.5 / 64.0)) * 64 if world > 1 else 10000
    r0, r1 = sk.triangle_row_partition(n, rank, world)
    steps = 5
    out = {"n_gpus": world, "n": n}
    bins = synth.make_bins(n, 16, KMERS, seed=world, device=f"cuda:{local}")
    torch.cuda.synchronize()
    for mode in ("zerocopy", "staged", "bulk"):
        os.environ["SKD_ZEROCOPY"] = "0" if mode == "staged" else "1"
        ctx = sk.Context(local)
        g = sk.SketchDb.upload(ctx, bins.data_ptr(), n, KMERS, 16, bins_loc=sk.SKD_DEVICE)
        dt = sk.set_k(g)
        my_pairs = int(ctx.lib.skd_self_pairs_in_rows(n, r0, r1))
        nbytes = my_pairs * 8
        pin = ctx.pinned((my_pairs * 2,), np.float32)
        dbuf = ctx.device_alloc(nbytes) if mode == "bulk" else None

        def step():
            if mode == "bulk":
                ctx._check(ctx.lib.skd_self_dists_all(ctx.handle, g.handle, dt.c, None, 0.64, r0, r1, dbuf, sk.SKD_DEVICE))
                ctx.memcpy(pin.array, sk.SKD_HOST, dbuf, sk.SKD_DEVICE, nbytes)
            else:
                ctx._check(ctx.lib.skd_self_dists_all(ctx.handle, g.handle, dt.c, None, 0.64, r0, r1,
                                                      pin.array.ctypes.data, sk.SKD_HOST))
        for _ in range(2):
            step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        barrier()
        ms = maxr((time.perf_counter() - t0) / steps * 1e3)
        tot = maxr(float(nbytes)) * world
        out[mode] = {"ms_per_step": ms, "aggregate_host_ingest_GBps": tot / (ms * 1e-3) / 1e9,
                     "checksum_rank0": float(pin.array[:1000].astype(np.float64).sum())}
        if mode == "bulk":  # the ceiling: the bare copy, all ranks at once
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                ctx.memcpy(pin.array, sk.SKD_HOST, dbuf, sk.SKD_DEVICE, nbytes)
            barrier()
            ms = maxr((time.perf_counter() - t0) / steps * 1e3)
            out["bare_d2h"] = {"ms_per_step": ms, "aggregate_host_ingest_GBps": tot / (ms * 1e-3) / 1e9,
                               "bytes_per_rank": nbytes}
            ctx.device_free(dbuf)
        pin.free()
        g.free()
        ctx.close()
    if rank == 0:
        print(json.dumps(out), flush=True)
    barrier()
    if world > 1:
        dist.destroy_process_group()


main()

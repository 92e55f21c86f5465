"""Runs BASELINE.json configs[2..4] (and reduced variants) on the visible GPU(s) and prints one JSON line each.

    python tools/run_configs.py [--scale 1.0] [--configs 3,4,5] [--out profiles/r1_configs.jsonl]
    torchrun --nproc-per-node N tools/run_configs.py ...      (rows / queries / refs sharded over ranks)

config 3: n=100k genomes, 1024 bins, K=8, self kNN 50 core/accessory (rows split evenly over ranks)
config 4: 10k queries vs 1M refs, 1024 bins, K=8: cross kNN 50 (queries split) and dense (ref row blocks,
          results left in HBM chunk by chunk: 80 GB of f32 pairs do not fit one host buffer)
config 5: n=200k, 10 240 bins (W=160), k=21: dense Jaccard (triangle row blocks, chunked in HBM) and self kNN 50
--scale shrinks n (not the sketch shape) for quick runs.  Kernel time = CUDA events on the library's stream.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import sketchlib.rust_b200 as sk  # noqa: E402
from sketchlib.rust_b200 import synth  # noqa: E402

KMERS8 = list(range(15, 31, 2))


def make_db(ctx, n, W, kmers, seed, dev):
    parts, chunk = [], 100_000
    t = time.time()
    bins = torch.empty(n * len(kmers) * W * 14, dtype=torch.int64, device=dev)
    per = len(kmers) * W * 14
    for s0 in range(0, n, chunk):
        s1 = min(n, s0 + chunk)
        bins[s0 * per:s1 * per] = synth.make_bins(s1 - s0, W, kmers, seed=seed + s0, device=dev)
    torch.cuda.synchronize()
    gen_s = time.time() - t
    t = time.time()
    g = sk.SketchDb.upload(ctx, bins.data_ptr(), n, kmers, W, bins_loc=sk.SKD_DEVICE)
    up_s = time.time() - t
    del bins, parts
    torch.cuda.empty_cache()
    return g, gen_s, up_s


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--configs", default="3,4,5")
    ap.add_argument("--out", default="")
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(dev))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ctx = sk.Context(local)
    want = [int(c) for c in args.configs.split(",")]
    lines = []

    def emit(d):
        d.update({"n_gpus": world, "scale": args.scale})
        if rank == 0:
            print(json.dumps(d), flush=True)
            lines.append(d)

    def timed(fn):
        best = None
        for _ in range(args.reps):
            barrier()
            fn()
            ms = maxr(ctx.last_kernel_ms_total)
            best = ms if best is None else min(best, ms)
        return best

    # accumulate kernel ms over several library calls
    class Acc:
        total = 0.0
    ctx.last_kernel_ms_total = 0.0

    def symmetric_knn(g, n, dt, knn, with_aux):
        """Symmetric self kNN over all ranks: each unordered pair once; ranks take equal-pair triangle row blocks and
        produce partial lists for all n rows, exchanged with ONE NCCL all-to-all and merged by the rank that owns the
        rows; final rows are all-gathered.  Returns (ms, exchange+merge ms, checksum of d0)."""
        r0, r1 = sk.triangle_row_partition(n, rank, world)
        n_pad = (n + world - 1) // world * world   # equal splits for the all-to-all
        rows_mine = n_pad // world
        keys = torch.full((n_pad, knn), -1, dtype=torch.int64, device=dev)   # -1 = UINT64_MAX padding
        aux = torch.zeros((n_pad, knn), dtype=torch.float32, device=dev) if with_aux else None
        idx_o = torch.empty((rows_mine, knn), dtype=torch.int64, device=dev)
        d0_o = torch.empty((rows_mine, knn), dtype=torch.float32, device=dev)
        d1_o = torch.empty((rows_mine, knn), dtype=torch.float32, device=dev) if with_aux else None
        ptr = lambda t: t.data_ptr() if t is not None else None
        if world > 1:
            recv_keys = torch.empty((world, rows_mine, knn), dtype=torch.int64, device=dev)
            recv_aux = torch.empty((world, rows_mine, knn), dtype=torch.float32, device=dev) if with_aux else None
            all_idx = torch.empty((n_pad, knn), dtype=torch.int64, device=dev)
            all_d0 = torch.empty((n_pad, knn), dtype=torch.float32, device=dev)
            all_d1 = torch.empty((n_pad, knn), dtype=torch.float32, device=dev) if with_aux else None
        xchg = {"ms": 0.0}

        def run():
            ctx._check(ctx.lib.skd_self_knn_partial(ctx.handle, g.handle, knn, dt.c, None, 0.64, r0, r1, keys.data_ptr(),
                                                    ptr(aux), sk.SKD_DEVICE))
            ms = ctx.last_kernel_ms
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            torch.cuda.synchronize()
            e0.record()
            if world > 1:
                # rank r receives, from every rank, the partial lists of ITS rows [r*rows_mine, (r+1)*rows_mine)
                dist.all_to_all_single(recv_keys.view(-1), keys.view(-1))
                if with_aux:
                    dist.all_to_all_single(recv_aux.view(-1), aux.view(-1))
                kp, ap, parts = recv_keys.data_ptr(), ptr(recv_aux), world
            else:
                kp, ap, parts = keys.data_ptr(), ptr(aux), 1
            e1.record()
            torch.cuda.synchronize()
            ctx._check(ctx.lib.skd_knn_merge(ctx.handle, kp, ap, parts, rows_mine, knn, dt.c, sk.SKD_DEVICE,
                                             idx_o.data_ptr(), d0_o.data_ptr(), ptr(d1_o), sk.SKD_DEVICE))
            merge_ms = ctx.last_kernel_ms
            e2.record()
            if world > 1:  # final sparse rows to every rank (rank 0 would write the text)
                dist.all_gather_into_tensor(all_idx, idx_o)
                dist.all_gather_into_tensor(all_d0, d0_o)
                if with_aux:
                    dist.all_gather_into_tensor(all_d1, d1_o)
            e3 = torch.cuda.Event(enable_timing=True)
            e3.record()
            torch.cuda.synchronize()
            xchg["ms"] = e0.elapsed_time(e1) + merge_ms + e2.elapsed_time(e3)
            ctx.last_kernel_ms_total = ms + xchg["ms"]
        ms = timed(run)
        checksum = float((all_d0 if world > 1 else d0_o)[:n].double().sum().item())
        return ms, maxr(xchg["ms"]), checksum

    if 3 in want:
        n = int(100_000 * args.scale)
        g, gen_s, up_s = make_db(ctx, n, 16, KMERS8, 3, dev)
        ms, xms, checksum = symmetric_knn(g, n, sk.set_k(g), 50, True)
        emit({"config": 3, "what": "self kNN 50 core/acc, 1024 bins, K=8 (symmetric: each unordered pair once)", "n": n,
              "kernel_ms": ms, "exchange_and_merge_ms": xms,
              "unique_pairs_per_s": n * (n - 1) / 2 / (ms * 1e-3),
              "reference_equivalent_ordered_pairs_per_s": n * (n - 1) / (ms * 1e-3),
              "lop3_T_per_s": n * (n - 1) / 2 * 3584 / (ms * 1e-3) / 1e12 / world, "db_hbm_bytes": g.hbm_bytes,
              "checksum": checksum,
              "note": "the reference evaluates every ordered pair (i, j != i); here each unordered pair is evaluated once "
                      "and credited to both rows; ranks own equal-pair triangle row blocks; partial lists are exchanged "
                      "with one NCCL all-to-all, each rank merges its rows, final rows are all-gathered; "
                      "lop3_T_per_s is per GPU"})
        g.free()

    if 4 in want:
        n_ref, n_q = int(1_000_000 * args.scale), int(10_000 * max(args.scale, 0.1))
        # references are SHARDED over ranks (each rank loads only its 1/N of the 14.3 GB .skd), queries replicated
        a0, a1 = sk.row_partition(n_ref, rank, world)
        gr, _, _ = make_db(ctx, a1 - a0, 16, KMERS8, 4 + a0, dev)
        gq, _, _ = make_db(ctx, n_q, 16, KMERS8, 5, dev)
        dt = sk.set_k(gq)
        knn = 50
        keys = torch.empty((n_q, knn), dtype=torch.int64, device=dev)
        aux = torch.empty((n_q, knn), dtype=torch.float32, device=dev)
        idx_o = torch.empty((n_q, knn), dtype=torch.int64, device=dev)
        d0_o = torch.empty((n_q, knn), dtype=torch.float32, device=dev)
        d1_o = torch.empty((n_q, knn), dtype=torch.float32, device=dev)
        if world > 1:
            all_keys = torch.empty((world, n_q, knn), dtype=torch.int64, device=dev)
            all_aux = torch.empty((world, n_q, knn), dtype=torch.float32, device=dev)
        xchg = {"ms": 0.0}

        def run_knn():
            ctx._check(ctx.lib.skd_cross_knn_partial(ctx.handle, gr.handle, gq.handle, knn, dt.c, None, None, 0.64, a0,
                                                     keys.data_ptr(), aux.data_ptr(), sk.SKD_DEVICE))
            ms = ctx.last_kernel_ms
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            if world > 1:
                dist.all_gather_into_tensor(all_keys, keys)
                dist.all_gather_into_tensor(all_aux, aux)
                kp, ap, parts = all_keys.data_ptr(), all_aux.data_ptr(), world
            else:
                kp, ap, parts = keys.data_ptr(), aux.data_ptr(), 1
            e1.record()
            torch.cuda.synchronize()
            ctx._check(ctx.lib.skd_knn_merge(ctx.handle, kp, ap, parts, n_q, knn, dt.c, sk.SKD_DEVICE, idx_o.data_ptr(),
                                             d0_o.data_ptr(), d1_o.data_ptr(), sk.SKD_DEVICE))
            xchg["ms"] = e0.elapsed_time(e1) + ctx.last_kernel_ms
            ctx.last_kernel_ms_total = ms + xchg["ms"]
        ms = timed(run_knn)
        emit({"config": 4, "what": "cross kNN 50 core/acc (10k queries x 1M refs), 1024 bins, K=8", "n_ref": n_ref,
              "n_query": n_q, "kernel_ms": ms, "exchange_and_merge_ms": maxr(xchg["ms"]),
              "pairs_per_s": n_ref * n_q / (ms * 1e-3), "lop3_T_per_s": n_ref * n_q * 3584 / (ms * 1e-3) / 1e12 / world,
              "ref_shard_hbm_bytes": gr.hbm_bytes, "checksum": float(d0_o.double().sum().item()),
              "note": "references sharded over ranks, queries replicated; per-rank partial top-k exchanged with one NCCL "
                      "all-gather (n_query x knn x 12 B per rank) and merged; lop3_T_per_s is per GPU"})
        # dense: this rank's ref rows, in chunks whose results stay in HBM (4 GB buffer)
        rows_per = max(64, (4 << 30) // (n_q * 8) // 64 * 64)
        dbuf = ctx.device_alloc(rows_per * n_q * 8)
        dtr = sk.set_k(gr)

        def run_dense():
            tot = 0.0
            for s0 in range(0, a1 - a0, rows_per):
                e = min(a1 - a0, s0 + rows_per)
                ctx._check(ctx.lib.skd_cross_dists_all(ctx.handle, gr.handle, gq.handle, dtr.c, None, None, 0.64, s0, e, dbuf,
                                                       sk.SKD_DEVICE))
                tot += ctx.last_kernel_ms
            ctx.last_kernel_ms_total = tot
        ms = timed(run_dense)
        emit({"config": 4, "what": "cross dense core/acc (ref-major), results left in HBM per 4 GB chunk", "n_ref": n_ref,
              "n_query": n_q, "kernel_ms": ms, "pairs_per_s": n_ref * n_q / (ms * 1e-3),
              "lop3_T_per_s": n_ref * n_q * 3584 / (ms * 1e-3) / 1e12 / world,
              "ref_stream_GBps_per_gpu": gr.hbm_bytes / (ms * 1e-3) / 1e9,
              "note": "refs sharded over ranks; ref_stream = each ref slab counted once (the tile kernel is INT-bound at "
                      "10k queries per streamed ref, not HBM-bound)"})
        ctx.device_free(dbuf)
        del keys, aux, idx_o, d0_o, d1_o
        gr.free()
        gq.free()

    if 5 in want:
        n = int(200_000 * args.scale)
        g, _, _ = make_db(ctx, n, 160, [21], 6, dev)
        r0, r1 = sk.triangle_row_partition(n, rank, world)
        dt = sk.set_k(g, 21)
        lop3_pair = 160 * 28
        # chunk rows so that each chunk's results fit a 4 GB device buffer
        chunks, s = [], r0
        while s < r1:
            e = s + 64
            while e < r1 and int(ctx.lib.skd_self_pairs_in_rows(n, s, min(r1, e + 64))) * 4 <= (4 << 30):
                e += 64
            chunks.append((s, min(e, r1)))
            s = min(e, r1)
        dbuf = ctx.device_alloc(max(int(ctx.lib.skd_self_pairs_in_rows(n, a, b)) for a, b in chunks) * 4 + 16)

        def run_dense():
            tot = 0.0
            for a, b in chunks:
                ctx._check(ctx.lib.skd_self_dists_all(ctx.handle, g.handle, dt.c, None, 0.64, a, b, dbuf, sk.SKD_DEVICE))
                tot += ctx.last_kernel_ms
            ctx.last_kernel_ms_total = tot
        ms = timed(run_dense)
        pairs = n * (n - 1) // 2
        emit({"config": 5, "what": "dense Jaccard k=21, 10240 bins (W=160), results left in HBM per <=4 GB chunk", "n": n,
              "kernel_ms": ms, "pairs_per_s": pairs / (ms * 1e-3), "lop3_T_per_s": pairs * lop3_pair / (ms * 1e-3) / 1e12 / world,
              "db_hbm_bytes": g.hbm_bytes, "note": "equal-pair triangle row blocks over ranks"})
        ctx.device_free(dbuf)
        dta = sk.set_k(g, 21, ani=True)
        ms, xms, checksum = symmetric_knn(g, n, dta, 50, False)
        emit({"config": 5, "what": "self kNN 50 ANI k=21, 10240 bins (symmetric: each unordered pair once)", "n": n,
              "kernel_ms": ms, "exchange_and_merge_ms": xms, "unique_pairs_per_s": n * (n - 1) / 2 / (ms * 1e-3),
              "reference_equivalent_ordered_pairs_per_s": n * (n - 1) / (ms * 1e-3),
              "lop3_T_per_s": n * (n - 1) / 2 * lop3_pair / (ms * 1e-3) / 1e12 / world, "checksum": checksum})
        g.free()

    if rank == 0 and args.out:
        with open(os.path.join(ROOT, args.out), "a") as fh:
            for d in lines:
                fh.write(json.dumps(d) + "\n")
    barrier()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

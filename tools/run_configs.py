"""Runs BASELINE.json configs[2..4] at full size (or scaled down) on the visible GPU(s), one JSON line each, through the
library's own multi-GPU entry points (skd_comm_* peer groups: no orchestration lives in this script), optionally
checking sampled rows against the oracle at the size that was timed.

    python tools/run_configs.py [--scale 1.0] [--configs 3,4,5] [--check] [--out profiles/r2_configs_1gpu.jsonl]
    torchrun --nproc-per-node N tools/run_configs.py ...

config 3: n=100k genomes, 1024 bins, K=8, self kNN 50 core/accessory     (skd_comm_self_dists_knn)
config 4: 10k queries vs 1M refs, 1024 bins, K=8, queries drawn from the references' families:
          cross kNN 50, refs sharded over ranks                           (skd_comm_cross_dists_knn)
          dense cross, this rank's ref rows, results left in HBM per 4 GB chunk (skd_cross_dists_all row ranges)
config 5: n=200k, 10 240 bins (W=160), k=21: dense Jaccard over this rank's triangle rows (skd_self_dists_all row
          ranges, results left in HBM per 4 GB chunk) and self kNN 50 ANI   (skd_comm_self_dists_knn)
--scale shrinks n (not the sketch shape).  Times are device times (CUDA events inside the library), max over ranks.
--check: rank 0 regenerates the planes on the host and compares sampled rows with oracle/ (CPU restatement of
src/distances/mod.rs) -- neighbour sets outside exact ties, values within 1e-6 (bit-exact for single k).
torch.distributed is used for the rendezvous (handle exchange) and for max-over-ranks only.
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
CHUNK = 100_000
TOL = 1e-6


def gen_chunk(s0, s1, W, kmers, seed, n_families, dev):
    """Samples [s0, s1) of a synthetic collection: the families depend on (seed, n_families) only, the members on the
    chunk, so any rank can generate any slice and a query set (another member stream) is related to it."""
    return synth.make_bins(s1 - s0, W, kmers, seed=seed, n_families=n_families, member_seed=1000 + s0, device=dev)


def gen_range(a, b, W, kmers, seed, n_families, dev):
    per = len(kmers) * W * 14
    bins = torch.empty((b - a) * per, dtype=torch.int64, device=dev)
    c0 = a // CHUNK * CHUNK
    for s0 in range(c0, b, CHUNK):
        s1 = s0 + CHUNK
        lo, hi = max(a, s0), min(b, s1)
        chunk = gen_chunk(s0, s1, W, kmers, seed, n_families, dev)
        bins[(lo - a) * per:(hi - a) * per] = chunk[(lo - s0) * per:(hi - s0) * per]
        del chunk
    torch.cuda.synchronize()
    return bins


def knn_row_check(idx, d0, d1, oidx, od0, od1, exact):
    """Neighbour sets identical outside exact ties (clearly-below-boundary entries must be in both rows), values of
    common neighbours within TOL (bit-equal keys when `exact`).  Returns (max |diff|, set mismatches, entries)."""
    worst, bad, entries = 0.0, 0, 0
    for r in range(idx.shape[0]):
        kth = float(od0[r, -1])
        ours, theirs = set(idx[r].tolist()), set(oidx[r].tolist())
        slack = 0.0 if exact else 2 * TOL
        lo_o = od0[r] < kth - slack if kth >= od0[r, 0] else od0[r] > kth + slack     # ANI rows descend
        lo_g = d0[r] < kth - slack if kth >= od0[r, 0] else d0[r] > kth + slack
        bad += len(set(oidx[r][lo_o].tolist()) - ours) + len(set(idx[r][lo_g].tolist()) - theirs)
        pos = {int(j): t for t, j in enumerate(oidx[r])}
        for t, j in enumerate(idx[r]):
            if int(j) in pos:
                entries += 1
                worst = max(worst, abs(float(d0[r, t]) - float(od0[r, pos[int(j)]])))
                if d1 is not None:
                    worst = max(worst, abs(float(d1[r, t]) - float(od1[r, pos[int(j)]])))
        worst = max(worst, float(np.abs(np.sort(d0[r].astype(np.float64)) - np.sort(od0[r].astype(np.float64))).max()))
    return worst, bad, entries


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--configs", default="3,4,5")
    ap.add_argument("--out", default="")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--check", action="store_true", help="compare sampled rows with the oracle at the timed size")
    ap.add_argument("--check-rows", type=int, default=64)
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
    n3 = int(100_000 * args.scale)
    n4r, n4q = int(1_000_000 * args.scale), int(10_000 * max(args.scale, 0.1))
    n5 = int(200_000 * args.scale)
    win = 1 << 20
    if 3 in want:
        win = max(win, sk.PeerGroup.window_bytes_knn(n3, 50, world))
    if 4 in want:
        win = max(win, sk.PeerGroup.window_bytes_knn(n4q, 50, world))
    if 5 in want:
        win = max(win, sk.PeerGroup.window_bytes_knn(n5, 50, world))
    grp = sk.PeerGroup(ctx, rank, world, win)
    if world > 1:
        handles = [None] * world
        dist.all_gather_object(handles, grp.export())
        grp.connect(handles)
    lines = []
    threads = os.cpu_count() or 1

    def emit(d):
        d.update({"n_gpus": world, "scale": args.scale})
        if rank == 0:
            print(json.dumps(d), flush=True)
            lines.append(d)

    def timed(fn):
        """fn() returns this rank's device ms; best of --reps, max over ranks."""
        best = None
        for _ in range(args.reps):
            barrier()
            ms = maxr(fn())
            best = ms if best is None else min(best, ms)
        return best

    def upload_whole(n, W, kmers, seed, n_families):
        bins = gen_range(0, n, W, kmers, seed, n_families, dev)
        g = sk.SketchDb.upload(ctx, bins.data_ptr(), n, kmers, W, bins_loc=sk.SKD_DEVICE)
        del bins
        torch.cuda.empty_cache()
        return g

    def host_planes(a, b, W, kmers, seed, n_families):
        return synth.as_u64_numpy(gen_range(a, b, W, kmers, seed, n_families, dev).cpu())

    def group_self_knn(g, n, dt, knn):
        """skd_comm_self_dists_knn, final rows on rank 0's host.  Returns (ms, exchange ms, (idx, d0, d1) on rank 0)."""
        with_aux = dt.kind == sk.CORE_ACC
        out = None
        if rank == 0:
            out = (np.zeros((n, knn), np.uint64), np.zeros((n, knn), np.float32),
                   np.zeros((n, knn), np.float32) if with_aux else None)
        x = {"ms": 0.0}

        def run():
            grp.self_dists_knn(g, knn, dt, root=0, out=out if rank == 0 else (None, None, None), out_loc=sk.SKD_HOST)
            x["ms"] = grp.last_exchange_ms
            return grp.last_op_ms
        ms = timed(run)
        return ms, maxr(x["ms"]), out

    def check_self_knn(out, n, W, kmers, seed, n_families, knn, k_idx, ani):
        import oracle
        rng = np.random.default_rng(11)
        db = oracle.Db(host_planes(0, n, W, kmers, seed, n_families), n, kmers, W)
        starts = sorted(set(int(s) for s in rng.integers(0, n - 8, size=max(1, args.check_rows // 8))) | {0, n - 8})
        worst, bad, entries, rows = 0.0, 0, 0, 0
        for s in starts:
            oidx, od0, od1 = oracle.self_dists_knn(db, knn, k_idx=k_idx, ani=ani, threads=threads, rows=(s, s + 8))
            w, b, e = knn_row_check(out[0][s:s + 8], out[1][s:s + 8], out[2][s:s + 8] if out[2] is not None else None,
                                    oidx, od0, od1, exact=k_idx >= 0)
            worst, bad, entries, rows = max(worst, w), bad + b, entries + e, rows + 8
        return {"rows": rows, "entries_compared": entries, "max_abs_diff": worst, "set_mismatches_outside_ties": bad,
                "tolerance": 0.0 if k_idx >= 0 else TOL, "checker": "oracle.self_dists_knn on the same planes"}

    if 3 in want:
        n, fam = n3, max(1, n3 // 50)
        g = upload_whole(n, 16, KMERS8, 3, fam)
        dt = sk.set_k(g)
        ms, xms, out = group_self_knn(g, n, dt, 50)
        line = {"config": 3, "what": "self kNN 50 core/acc, 1024 bins, K=8 (each unordered pair once)", "n": n,
                "device_ms": ms, "exchange_ms": xms, "unique_pairs_per_s": n * (n - 1) / 2 / (ms * 1e-3),
                "reference_equivalent_ordered_pairs_per_s": n * (n - 1) / (ms * 1e-3),
                "lop3_T_per_s_per_gpu": n * (n - 1) / 2 * 3584 / (ms * 1e-3) / 1e12 / world, "db_hbm_bytes": g.hbm_bytes,
                "entry": "skd_comm_self_dists_knn (triangle row blocks per rank, peer-memory merge, rows on rank 0's host)"}
        if rank == 0:
            line["checksum_core"] = float(out[1].astype(np.float64).sum())
            line["checksum_accessory"] = float(out[2].astype(np.float64).sum())
            if args.check:
                line["check"] = check_self_knn(out, n, 16, KMERS8, 3, fam, 50, -1, False)
        emit(line)
        g.free()

    if 4 in want:
        import oracle
        n_ref, n_q, fam = n4r, n4q, max(1, n4r // 50)
        a0, a1 = sk.slice_rows(n_ref, rank, world, 64)      # this rank's reference shard
        rb = gen_range(a0, a1, 16, KMERS8, 4, fam, dev)
        gr = sk.SketchDb.upload(ctx, rb.data_ptr(), a1 - a0, KMERS8, 16, bins_loc=sk.SKD_DEVICE)
        del rb
        qb = synth.make_bins(n_q, 16, KMERS8, seed=4, n_families=fam, member_seed=7, device=dev)  # the refs' families
        torch.cuda.synchronize()
        q_host = synth.as_u64_numpy(qb.cpu()) if rank == 0 and args.check else None
        gq = sk.SketchDb.upload(ctx, qb.data_ptr(), n_q, KMERS8, 16, bins_loc=sk.SKD_DEVICE)
        del qb
        torch.cuda.empty_cache()
        dt = sk.set_k(gq)
        knn = 50
        out = None
        if rank == 0:
            out = (np.zeros((n_q, knn), np.uint64), np.zeros((n_q, knn), np.float32), np.zeros((n_q, knn), np.float32))
        x = {"ms": 0.0}

        def run_knn():
            grp.cross_dists_knn(gr, gq, knn, dt, a0, n_ref, root=0, out=out if rank == 0 else (None, None, None),
                                out_loc=sk.SKD_HOST)
            x["ms"] = grp.last_exchange_ms
            return grp.last_op_ms
        ms = timed(run_knn)
        line = {"config": 4, "what": "cross kNN 50 core/acc (10k queries x 1M refs of the same families), 1024 bins, K=8",
                "n_ref": n_ref, "n_query": n_q, "device_ms": ms, "exchange_ms": maxr(x["ms"]),
                "pairs_per_s": n_ref * n_q / (ms * 1e-3),
                "lop3_T_per_s_per_gpu": n_ref * n_q * 3584 / (ms * 1e-3) / 1e12 / world, "ref_shard_hbm_bytes": gr.hbm_bytes,
                "entry": "skd_comm_cross_dists_knn (references sharded over ranks, queries replicated, peer-memory merge)"}
        if rank == 0:
            line["checksum_core"] = float(out[1].astype(np.float64).sum())
            line["checksum_accessory"] = float(out[2].astype(np.float64).sum())
            line["rows_with_a_relative"] = float((out[1][:, 0] < 1.0).mean())
            if args.check:
                rng = np.random.default_rng(12)
                qs = np.sort(rng.choice(n_q, size=min(n_q, args.check_rows), replace=False))
                qsub = oracle.Db(q_host.reshape(n_q, -1)[qs].reshape(-1), len(qs), KMERS8, 16)
                cand = [[], [], []]
                for s0 in range(0, n_ref, CHUNK):   # the oracle over one reference chunk at a time (bounded host memory)
                    s1 = min(n_ref, s0 + CHUNK)
                    rdb = oracle.Db(host_planes(s0, s1, 16, KMERS8, 4, fam), s1 - s0, KMERS8, 16)
                    oi, o0, o1 = oracle.cross_dists_knn(rdb, qsub, min(knn, s1 - s0), threads=threads)
                    cand[0].append(oi.astype(np.uint64) + s0)
                    cand[1].append(o0)
                    cand[2].append(o1)
                ci, c0, c1 = (np.concatenate(c, axis=1) for c in cand)
                order = np.lexsort((ci, c0), axis=1)[:, :knn]   # ascending core, ties by index
                oidx, od0, od1 = (np.take_along_axis(c, order, axis=1) for c in (ci, c0, c1))
                w, b, e = knn_row_check(out[0][qs], out[1][qs], out[2][qs], oidx, od0, od1, exact=False)
                line["check"] = {"rows": len(qs), "entries_compared": e, "max_abs_diff": w,
                                 "set_mismatches_outside_ties": b, "tolerance": TOL,
                                 "checker": "oracle.cross_dists_knn per 100k-reference chunk, merged by (core, index)"}
        emit(line)
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
            return tot
        ms = timed(run_dense)
        emit({"config": 4, "what": "cross dense core/acc (ref-major), results left in HBM per 4 GB chunk", "n_ref": n_ref,
              "n_query": n_q, "device_ms": ms, "pairs_per_s": n_ref * n_q / (ms * 1e-3),
              "lop3_T_per_s_per_gpu": n_ref * n_q * 3584 / (ms * 1e-3) / 1e12 / world,
              "entry": "skd_cross_dists_all row ranges (refs sharded over ranks; no exchange)"})
        ctx.device_free(dbuf)
        gr.free()
        gq.free()

    if 5 in want:
        import oracle
        n, fam = n5, max(1, n5 // 50)
        g = upload_whole(n, 160, [21], 6, fam)
        r0, r1 = sk.triangle_row_partition(n, rank, world)
        dt = sk.set_k(g, 21)
        lop3_pair = 160 * 28
        chunks, s = [], r0   # rows whose results fit a 4 GB device buffer
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
            return tot
        ms = timed(run_dense)
        pairs = n * (n - 1) // 2
        line = {"config": 5, "what": "dense Jaccard k=21, 10240 bins (W=160), results left in HBM per <=4 GB chunk", "n": n,
                "device_ms": ms, "pairs_per_s": pairs / (ms * 1e-3),
                "lop3_T_per_s_per_gpu": pairs * lop3_pair / (ms * 1e-3) / 1e12 / world, "db_hbm_bytes": g.hbm_bytes,
                "entry": "skd_self_dists_all row ranges (equal-pair triangle row blocks over ranks; no exchange)"}
        if rank == 0 and args.check:   # the first rows of the last chunk (still in dbuf) against the oracle, bit for bit
            a, b = chunks[-1]
            rows = min(4, b - a)
            cnt = int(ctx.lib.skd_self_pairs_in_rows(n, a, a + rows))
            got = np.empty(cnt, np.float32)
            ctx.memcpy(got, sk.SKD_HOST, dbuf, sk.SKD_DEVICE, cnt * 4)
            db = oracle.Db(host_planes(0, n, 160, [21], 6, fam), n, [21], 160)
            first = int(oracle.lib().oracle_square_to_condensed(a, a + 1, n))
            pad = first % 1000   # the oracle keeps the reference's 1000-distance chunking: aligned starts only
            ref = oracle.self_dists_all(db, k_idx=0, first=first - pad, count=cnt + pad, threads=threads)[pad:]
            line["check"] = {"rows": rows, "entries_compared": cnt, "bit_identical": bool(np.array_equal(got, ref)),
                             "max_abs_diff": float(np.abs(got.astype(np.float64) - ref).max()),
                             "checker": "oracle.self_dists_all (k=21 Jaccard) on the same planes"}
            del db
        emit(line)
        ctx.device_free(dbuf)
        dta = sk.set_k(g, 21, ani=True)
        ms, xms, out = group_self_knn(g, n, dta, 50)
        line = {"config": 5, "what": "self kNN 50 ANI k=21, 10240 bins (each unordered pair once)", "n": n,
                "device_ms": ms, "exchange_ms": xms, "unique_pairs_per_s": n * (n - 1) / 2 / (ms * 1e-3),
                "reference_equivalent_ordered_pairs_per_s": n * (n - 1) / (ms * 1e-3),
                "lop3_T_per_s_per_gpu": n * (n - 1) / 2 * lop3_pair / (ms * 1e-3) / 1e12 / world,
                "entry": "skd_comm_self_dists_knn"}
        if rank == 0:
            line["checksum"] = float(out[1].astype(np.float64).sum())
            if args.check:
                line["check"] = check_self_knn(out, n, 160, [21], 6, fam, 50, 0, True)
        emit(line)
        g.free()

    if rank == 0 and args.out:
        with open(os.path.join(ROOT, args.out), "a") as fh:
            for d in lines:
                fh.write(json.dumps(d) + "\n")
    barrier()
    grp.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    t_all = time.time()
    main()

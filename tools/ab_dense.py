"""A/B probe of the dense match-count kernels (not the contract bench): dense self core/acc on the bench's family data,
results resident in HBM, plus a symmetric kNN-50 run.  Select the library with SKD_LIB_OVERRIDE, the old kernel with
SKD_DENSE_V1=1.     python tools/ab_dense.py [n_dense] [n_knn] [tag]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import sketchlib.rust_b200 as sk  # noqa: E402
from sketchlib.rust_b200 import synth  # noqa: E402

KMERS = list(range(15, 31, 2))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    n_knn = int(sys.argv[2]) if len(sys.argv) > 2 else 30000
    tag = sys.argv[3] if len(sys.argv) > 3 else ""
    ctx = sk.Context(0)
    out = {"tag": tag, "lib": os.environ.get("SKD_LIB_OVERRIDE", "default"), "v1": os.environ.get("SKD_DENSE_V1", "0")}
    bins = synth.make_bins(n, 16, KMERS, seed=1, device="cuda")
    torch.cuda.synchronize()   # the generator ran on torch's stream, the loader reads on the library's
    g = sk.SketchDb.upload(ctx, bins.data_ptr(), n, KMERS, 16, bins_loc=sk.SKD_DEVICE)
    dt = sk.set_k(g)
    pairs = n * (n - 1) // 2
    dout = ctx.device_alloc(pairs * 8)
    ms = []
    for it in range(6):
        ctx._check(ctx.lib.skd_self_dists_all(ctx.handle, g.handle, dt.c, None, 0.64, 0, n, dout, sk.SKD_DEVICE))
        ms.append(ctx.last_kernel_ms)
    best = min(ms[1:])
    out["dense_ms"] = best
    out["dense_tlop3"] = pairs * 3584 / (best * 1e-3) / 1e12
    chk = torch.empty(pairs * 2, dtype=torch.float32, device="cuda")
    ctx.memcpy(chk.data_ptr(), sk.SKD_DEVICE, dout, sk.SKD_DEVICE, pairs * 8)
    out["dense_checksum"] = float(chk.double().sum().item())
    ctx.device_free(dout)
    g.free()
    del bins
    if n_knn:
        per = len(KMERS) * 16 * 14
        bins = torch.empty(n_knn * per, dtype=torch.int64, device="cuda")
        for s0 in range(0, n_knn, 50000):
            s1 = min(n_knn, s0 + 50000)
            bins[s0 * per:s1 * per] = synth.make_bins(s1 - s0, 16, KMERS, seed=3 + s0, device="cuda")
        torch.cuda.synchronize()
        g = sk.SketchDb.upload(ctx, bins.data_ptr(), n_knn, KMERS, 16, bins_loc=sk.SKD_DEVICE)
        del bins
        ms, tile = [], []
        for it in range(3):
            r = sk.self_dists_knn(g, knn=50, dist_type=sk.set_k(g))
            ms.append(ctx.last_kernel_ms)
            tile.append(ctx.last_tile_kernel_ms)
        kp = n_knn * (n_knn - 1) // 2
        out["knn_n"] = n_knn
        out["knn_ms"] = min(ms[1:])
        out["knn_tile_ms"] = min(tile[1:])
        out["knn_tile_tlop3"] = kp * 3584 / (min(tile[1:]) * 1e-3) / 1e12
        out["knn_checksum"] = float(r.d0.astype("float64").sum() + r.d1.astype("float64").sum())
        ctx.work_counters(reset=True)
        sk.self_dists_knn(g, knn=50, dist_type=sk.set_k(g))
        done, total, deferred = ctx.work_counters()
        out["knn_steps_computed_frac"] = done / max(total, 1)
        out["knn_pairs_deferred_frac"] = deferred / kp
    print(json.dumps(out), flush=True)


main()

"""Streaming query-vs-ref with FEW queries: n_ref references (1024 bins x 8 k) against nq queries, dense cross
core/acc with results left in HBM.  Reports the reference-plane stream rate against the HBM peak."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sketchlib.rust_b200 as sk
from sketchlib.rust_b200 import synth
n_ref = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
kmers = list(range(15, 31, 2))
ctx = sk.Context(0)
per = 8 * 16 * 14
bins = torch.empty(n_ref * per, dtype=torch.int64, device="cuda")
for s0 in range(0, n_ref, 100000):
    s1 = min(n_ref, s0 + 100000)
    bins[s0 * per:s1 * per] = synth.make_bins(s1 - s0, 16, kmers, seed=s0, device="cuda")
gr = sk.SketchDb.upload(ctx, bins.data_ptr(), n_ref, kmers, 16, bins_loc=sk.SKD_DEVICE)
del bins
torch.cuda.empty_cache()
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6551.4
LOP3_PEAK = float(os.environ.get("SKD_LOP3_PEAK", 18.44e12))  # tools/int_peak on this pool
for nq in (1, 4, 8, 9, 12, 16, 24, 32, 48, 64, 128):
    q = synth.make_bins(nq, 16, kmers, seed=7, device="cuda")
    gq = sk.SketchDb.upload(ctx, q.data_ptr(), nq, kmers, 16, bins_loc=sk.SKD_DEVICE)
    dt = sk.set_k(gr)
    out = ctx.device_alloc(n_ref * nq * 8)
    best = 1e9
    for _ in range(3):
        ctx._check(ctx.lib.skd_cross_dists_all(ctx.handle, gr.handle, gq.handle, dt.c, None, None, 0.64, 0, n_ref, out, sk.SKD_DEVICE))
        best = min(best, ctx.last_kernel_ms)
    r = sk.cross_dists_knn(gr, gq, knn=50, dist_type=dt)
    knn_ms = ctx.last_kernel_ms
    print(json.dumps({"n_ref": n_ref, "n_query": nq, "dense_ms": best, "ref_stream_GBps": gr.hbm_bytes / best / 1e6,
                      "frac_of_hbm_peak": gr.hbm_bytes / best / 1e6 / peak, "pairs_per_s": n_ref * nq / best * 1e3,
                      "knn50_ms": knn_ms,
                      # the path's two rooflines for this shape: one pass over the reference planes / the LOP3 issue rate
                      "bound_ms": max(gr.hbm_bytes / peak / 1e6, n_ref * nq * 3584 / LOP3_PEAK * 1e3),
                      "frac_of_bound": max(gr.hbm_bytes / peak / 1e6, n_ref * nq * 3584 / LOP3_PEAK * 1e3) / best}))
    ctx.device_free(out); gq.free()

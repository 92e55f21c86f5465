"""Is the e2e step bound by the zero-copy result stream?  Dense self core/acc, n=10000 family data, database resident:
(a) results to HBM, (b) results to pinned host memory (zero-copy stores), (c) = (b) after an async upload."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sketchlib.rust_b200 as sk
from sketchlib.rust_b200 import synth
n, W, kmers = 10000, 16, list(range(15, 31, 2))
ctx = sk.Context(0)
bins = synth.make_bins(n, W, kmers, seed=1, device="cuda")
pin = ctx.pinned((bins.numel(),), np.uint64)
ctx.memcpy(pin.array, sk.SKD_HOST, bins.data_ptr(), sk.SKD_DEVICE, bins.numel() * 8)
g = sk.SketchDb.upload(ctx, bins.data_ptr(), n, kmers, W, bins_loc=sk.SKD_DEVICE)
dt = sk.set_k(g)
npairs = n * (n - 1) // 2
out = ctx.pinned((npairs * 2,), np.float32)
dout = ctx.device_alloc(npairs * 8)
def run(dst, loc, db):
    ctx._check(ctx.lib.skd_self_dists_all(ctx.handle, db.handle, dt.c, None, 0.64, 0, n, dst, loc))
for name, dst, loc in (("HBM", dout, sk.SKD_DEVICE), ("pinned host (zero-copy)", out.array.ctypes.data, sk.SKD_HOST)):
    for _ in range(2): run(dst, loc, g)
    t0 = time.perf_counter()
    for _ in range(5): run(dst, loc, g)
    print(f"resident db -> {name}: {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms wall, kernel {ctx.last_kernel_ms:.2f} ms")
for _ in range(6):
    t0 = time.perf_counter()
    ga = sk.SketchDb.upload(ctx, pin.array, n, kmers, W, async_upload=True)
    run(out.array.ctypes.data, sk.SKD_HOST, ga)
    t1 = time.perf_counter()
    ga.free()
print(f"async upload + zero-copy: {(t1 - t0) * 1e3:.2f} ms wall")
for _ in range(6):
    t0 = time.perf_counter()
    ga = sk.SketchDb.upload(ctx, pin.array, n, kmers, W, async_upload=True)
    run(dout, sk.SKD_DEVICE, ga)
    t1 = time.perf_counter()
    ga.free()
print(f"async upload -> HBM: {(t1 - t0) * 1e3:.2f} ms wall")

"""Small driver for ncu: the bench step (dense all-vs-all core/acc, configs[1] shape) without torch, so that
every captured launch is one of this library's kernels.  usage: profile_step.py [n] [steps] [knn]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sketchlib.rust_b200 as sk

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
knn = int(sys.argv[3]) if len(sys.argv) > 3 else 0
W, kmers = 16, list(range(15, 31, 2))
ctx = sk.Context(0)
rng = np.random.default_rng(0)
bins = rng.integers(0, 1 << 63, size=n * len(kmers) * W * 14, dtype=np.uint64)
g = sk.SketchDb.upload(ctx, bins, n, kmers, W)
dt = sk.set_k(g)
npairs = n * (n - 1) // 2
dout = ctx.device_alloc(npairs * 8)
for _ in range(steps):
    ctx._check(ctx.lib.skd_self_dists_all(ctx.handle, g.handle, dt.c, None, 0.64, 0, n, dout, sk.SKD_DEVICE))
    print("dense ms", ctx.last_kernel_ms, "Gpairs/s", npairs / ctx.last_kernel_ms / 1e6)
if knn:
    for _ in range(steps):
        r = sk.self_dists_knn(g, knn=knn, dist_type=dt)
        print("knn ms", ctx.last_kernel_ms, "ordered Gpairs/s", n * (n - 1) / ctx.last_kernel_ms / 1e6)

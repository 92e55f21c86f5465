"""e2e probe: upload from pinned + dense self core/acc to pinned host output, wall-clock."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sketchlib.rust_b200 as sk
n, W, kmers = 10000, 16, list(range(15, 31, 2))
ctx = sk.Context(0)
rng = np.random.default_rng(0)
pin = ctx.pinned((n * 8 * W * 14,), np.uint64)
pin.array[:] = rng.integers(0, 1 << 64, size=pin.array.size, dtype=np.uint64)
npairs = n * (n - 1) // 2
out = ctx.pinned((npairs * 2,), np.float32)
import os
ASYNC = os.environ.get('E2E_ASYNC', '1') == '1'
for it in range(6):
    t0 = time.perf_counter()
    g = sk.SketchDb.upload(ctx, pin.array, n, kmers, W, async_upload=ASYNC)
    t1 = time.perf_counter()
    dt = sk.set_k(g)
    ctx._check(ctx.lib.skd_self_dists_all(ctx.handle, g.handle, dt.c, None, 0.64, 0, n, out.array.ctypes.data, sk.SKD_HOST))
    t2 = time.perf_counter()
    g.free()
    print(f"upload {1e3*(t1-t0):.2f} ms  dists {1e3*(t2-t1):.2f} ms (kernel {ctx.last_kernel_ms:.2f})  total {1e3*(t2-t0):.2f} ms  sum {float(out.array[:1000].sum()):.3f}")

"""Ad-hoc throughput probe (not the contract bench): dense self core/acc on random sketches."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sketchlib.rust_b200 as sk

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    K = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    ctx = sk.Context(0)
    rng = np.random.default_rng(0)
    kmers = list(range(15, 15 + 2 * K, 2))
    bins = rng.integers(0, 1 << 63, size=n * K * W * 14, dtype=np.uint64)
    t = time.time(); g = sk.SketchDb.upload(ctx, bins, n, kmers, W); print("upload s", time.time() - t, "hbm MB", g.hbm_bytes / 1e6)
    dt = sk.set_k(g) if K > 1 else sk.set_k(g, kmers[0])
    npairs = n * (n - 1) // 2
    dout = ctx.device_alloc(npairs * dt.n_dist_cols * 4)
    for it in range(5):
        ctx._check(ctx.lib.skd_self_dists_all(ctx.handle, g.handle, dt.c, None, 0.64, 0, n, dout, sk.SKD_DEVICE))
        ms = ctx.last_kernel_ms
        print(f"dense self n={n} W={W} K={K}: {ms:.3f} ms  {npairs / ms / 1e6:.1f} Gpairs/s  LOP3/s={npairs*K*W*28/ms/1e9:.2f} T")
    t = time.time(); dm = sk.self_dists_all(g, dist_type=dt); e = time.time() - t
    print(f"e2e host-out: {e*1e3:.1f} ms kernel {ctx.last_kernel_ms:.2f}")
    if n <= 200000:
        for it in range(3):
            t = time.time(); kn = sk.self_dists_knn(g, knn=50, dist_type=dt); e = time.time() - t
            ms = ctx.last_kernel_ms
            print(f"knn50 self: kernel {ms:.2f} ms  ordered pairs/s {n*(n-1)/ms/1e6:.1f} G   wall {e*1e3:.1f} ms")

main()

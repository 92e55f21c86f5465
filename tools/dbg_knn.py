import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sketchlib.rust_b200 as sk
from sketchlib.rust_b200 import synth
KMERS = list(range(15, 31, 2))
n = int(sys.argv[1])
ctx = sk.Context(0)
bins = synth.make_bins(n, 16, KMERS, seed=5, device="cuda")
torch.cuda.synchronize()
g = sk.SketchDb.upload(ctx, bins.data_ptr(), n, KMERS, 16, bins_loc=sk.SKD_DEVICE)
r = sk.self_dists_knn(g, knn=50, dist_type=sk.set_k(g))
rows = [0, 63, 64, n // 2, n - 1]
bad = 0
for row in rows:
    d = sk.cross_dists_all(g, g, dist_type=sk.set_k(g), rows=(row, row + 1)).distances.reshape(n, 2)
    key = d[:, 0].copy(); key[row] = np.inf
    want = np.sort(key)[:50]
    ok = np.array_equal(want, r.d0[row])
    bad += not ok
    print(n, "row", row, "ok" if ok else "MISMATCH", r.d0[row][:4], want[:4], r.idx[row][:4])
print("checksum", float(r.d0.astype(np.float64).sum()), "bad", bad)

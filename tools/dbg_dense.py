import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sketchlib.rust_b200 as sk
from sketchlib.rust_b200 import synth
KMERS = list(range(15, 31, 2))
n = int(sys.argv[1]); r0 = int(sys.argv[2]); mode = sys.argv[3]
ctx = sk.Context(0)
bins = synth.make_bins(n, 16, KMERS, seed=5, device="cuda")
torch.cuda.synchronize()
torch.cuda.synchronize()
g = sk.SketchDb.upload(ctx, bins.data_ptr(), n, KMERS, 16, bins_loc=sk.SKD_DEVICE)
rows = (r0, r0 + 64)
got = sk.cross_dists_all(g, g, dist_type=sk.set_k(g), rows=rows).distances.reshape(64, n, 2).copy()
gs = sk.self_dists_all(g, dist_type=sk.set_k(g), rows=rows).distances.reshape(-1, 2).copy()
f = f"/tmp/dbg_{n}_{r0}.npz"
if mode == "save":
    np.savez(f, got=got, gs=gs)
    sys.exit(0)
ref = np.load(f)
want = got; got = ref["got"]            # this process = v1 = reference
bad = np.abs(got.astype(np.float64) - want).max(axis=2) > 1e-6      # [64, n]
print("n", n, "rows", rows, "cross bad fraction", bad.mean())
per_cb = bad.reshape(64, -1, 64).mean(axis=(0, 2))
nz = np.flatnonzero(per_cb > 0)
print("column blocks with mismatches:", nz[:12], "...", nz[-5:] if len(nz) else "", "count", len(nz), "of", len(per_cb))
if len(nz):
    cb = nz[0]
    sub = bad[:, cb * 64:(cb + 1) * 64]
    print("first bad block", cb, "frac", sub.mean(), "bad rows", np.flatnonzero(sub.any(axis=1))[:16], "bad cols", np.flatnonzero(sub.any(axis=0))[:16])
    i, j = np.argwhere(sub)[0]
    print("example", r0 + i, cb * 64 + j, "v2", got[i, cb * 64 + j], "v1", want[i, cb * 64 + j])
bs = np.abs(ref["gs"].astype(np.float64) - gs).max(axis=1) > 1e-6
print("self rows bad", int(bs.sum()), "of", len(bs), "first bad at", np.flatnonzero(bs)[:5])

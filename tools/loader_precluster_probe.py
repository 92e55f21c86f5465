"""Measures (1) the .skd loader: file -> pinned staging -> HBM -> re-tile (skd_db_load_skd), full and --subset;
(2) the candidate-list (precluster) kNN kernel on family-structured candidates."""
import os, sys, time, json, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import sketchlib.rust_b200 as sk
from sketchlib.rust_b200 import synth, precluster
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
kmers = list(range(15, 31, 2))
ctx = sk.Context(0)
d = tempfile.mkdtemp()
path = os.path.join(d, "db.skd")
per = 8 * 16 * 14
with open(path, "wb") as fh:
    for s0 in range(0, n, 50000):
        s1 = min(n, s0 + 50000)
        fh.write(synth.as_u64_numpy(synth.make_bins(s1 - s0, 16, kmers, seed=s0, device="cuda")).tobytes())
size = os.path.getsize(path)
for rep in range(3):
    t = time.perf_counter(); g = sk.SketchDb.load_skd(ctx, path, n, kmers, 16); e = time.perf_counter() - t
    print(json.dumps({"what": "load_skd full", "bytes": size, "seconds": e, "GBps": size / e / 1e9}))
    if rep < 2: g.free()
rng = np.random.default_rng(0)
sub = rng.permutation(n)[: n // 10]
t = time.perf_counter(); gs = sk.SketchDb.load_skd(ctx, path, len(sub), kmers, 16, sample_indices=sub); e = time.perf_counter() - t
print(json.dumps({"what": "load_skd subset 10% random order", "bytes": len(sub) * per * 8, "seconds": e, "GBps": len(sub) * per * 8 / e / 1e9}))
gs.free()
# precluster: candidates = 64 consecutive samples around each row (family-like), knn 10
m = 50000
offs = np.arange(m + 1, dtype=np.uint64) * 64
idx = ((np.arange(m)[:, None] // 64 * 64 + np.arange(64)[None, :]) % n).astype(np.uint64).ravel()
nidx = np.zeros((n, 10), np.uint64); nd0 = np.zeros((n, 10), np.float32)
full_off = np.concatenate([offs, np.full(n - m, offs[-1], dtype=np.uint64)])
dt = sk.set_k(g, 21)
import ctypes as C
p = lambda a: a.ctypes.data_as(C.c_void_p)
for rep in range(2):
    t = time.perf_counter()
    ctx._check(ctx.lib.skd_self_dists_knn_candidates(ctx.handle, g.handle, 10, dt.c, None, 0.64, p(full_off), p(idx), 0, p(nidx), p(nd0)))
    e = time.perf_counter() - t
    print(json.dumps({"what": "candidate-list kNN (precluster)", "rows": n, "candidate_pairs": int(idx.size), "kernel_ms": ctx.last_kernel_ms,
                      "pairs_per_s": idx.size / ctx.last_kernel_ms * 1e3, "wall_s": e}))
os.remove(path)

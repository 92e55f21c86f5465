"""Small run of every kernel family for compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_probe.py
    compute-sanitizer --tool racecheck python tools/sanitize_probe.py

Sizes are tiny (the tools slow kernels down 10-100x) but cover ragged last blocks, the narrow/stream few-query
kernels, both select kernels, row-owned and symmetric kNN, candidate lists and the async upload."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import sketchlib.rust_b200 as sk  # noqa: E402
from helpers import synth_db, synth_pair  # noqa: E402

ctx = sk.Context(0)


def up(db):
    return sk.SketchDb.upload(ctx, db.bins, db.n, db.kmers, db.W, db.sketch_size, db.sample_stride, db.kmer_stride, db.names)


kmers = [15, 19, 23, 27]
db = synth_db(333, 16, kmers, seed=1, n_families=5)
g = up(db)
ca, jac, ani = sk.set_k(g), sk.set_k(g, 19), sk.set_k(g, 19, ani=True)
sk.self_dists_all(g, dist_type=ca)
sk.self_dists_all(g, dist_type=jac, rows=(70, 200))
sk.self_dists_knn(g, knn=10, dist_type=ca)                 # symmetric: rect tiles + select_sym rows/cols
sk.self_dists_knn(g, knn=64, dist_type=ani)
sk.self_dists_knn(g, knn=70, dist_type=jac)                # row-owned kernel (knn > 64)
sk.self_dists_knn(g, knn=5, dist_type=ca, rows=(64, 192))  # row shard
for nq in (1, 5, 8, 13, 40, 70):
    ref, qry = synth_pair(300 + nq, nq, 16, kmers, seed=nq, n_families=4)
    gr, gq = up(ref), up(qry)
    sk.match_counts(gr, gq)
    sk.cross_dists_all(gr, gq, dist_type=sk.set_k(gr))
    sk.cross_dists_all(gr, gq, dist_type=sk.set_k(gr, 19), rows=(37, 290))
    sk.cross_dists_knn(gr, gq, knn=9, dist_type=sk.set_k(gr))
    sk.cross_dists_knn(gr, gq, knn=64, dist_type=sk.set_k(gr, 19, ani=True))
    gr.free()
    gq.free()
big = synth_db(130, 160, [21], seed=3)                      # config-5 shape
gb = up(big)
sk.self_dists_all(gb, dist_type=sk.set_k(gb, 21))
sk.self_dists_knn(gb, knn=7, dist_type=sk.set_k(gb, 21, ani=True))
r5, q5 = synth_pair(200, 3, 160, [21], seed=4)
g5, h5 = up(r5), up(q5)
sk.cross_dists_all(g5, h5, dist_type=sk.set_k(g5, 21))
# candidate lists
skq = np.random.default_rng(0).integers(0, 6, size=(333, 4)).astype(np.uint16)
sk.self_dists_knn_precluster(g, skq, db.names, 5, jac, retain_unmatched="bruteforce")
# async upload + interleaved dense self
pin = ctx.pinned((db.bins.size,), np.uint64)
pin.array[:] = db.bins
ga = sk.SketchDb.upload(ctx, pin.array, db.n, db.kmers, db.W, async_upload=True)
sk.self_dists_all(ga, dist_type=sk.set_k(ga))
# round 2: early exit (deferred pairs, producer cut), tiny deferred-pair list, sorted kNN path, many k-mer lengths
k8 = [15, 17, 19, 21, 23, 25, 27, 29]
mix = synth_db(450, 16, k8, seed=7, n_families=12, unrelated_frac=0.3)
gm = up(mix)
sk.self_dists_all(gm, dist_type=sk.set_k(gm))
sk.self_dists_knn(gm, knn=20, dist_type=sk.set_k(gm))
print("work counters (computed, total, deferred):", ctx.work_counters())
rm, qm = synth_pair(400, 150, 16, k8, seed=8, n_families=10, unrelated_frac=0.3)
grm, gqm = up(rm), up(qm)
sk.cross_dists_all(grm, gqm, dist_type=sk.set_k(grm))
sk.cross_dists_knn(grm, gqm, knn=12, dist_type=sk.set_k(grm))
sk.self_dists_knn(gm, knn=449, dist_type=sk.set_k(gm, 19))        # knn = n - 1: dense keys + full row sort
sk.self_dists_knn(gm, knn=449, dist_type=sk.set_k(gm), rows=(64, 128))
k30 = list(range(5, 65, 2))
many = synth_db(150, 16, k30, seed=9, n_families=3)
gk = up(many)
sk.self_dists_all(gk, dist_type=sk.set_k(gk))
sk.self_dists_knn(gk, knn=10, dist_type=sk.set_k(gk))
os.environ["SKD_DEFER_CAP"] = "40"
ctx2 = sk.Context(0)
gm2 = sk.SketchDb.upload(ctx2, mix.bins, mix.n, mix.kmers, mix.W)
sk.self_dists_all(gm2, dist_type=sk.set_k(gm2))
print("sanitize probe done, launches:", ctx.launch_count + ctx2.launch_count)

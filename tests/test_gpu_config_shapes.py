"""Parity at the SHAPES of BASELINE.json configs[3] and [4], on hardware, at sizes the
oracle can sample in seconds:
  (i)  config-5 shape: 10 240 bins (W = 160), one k-mer length, n = 20 000 -- dense Jaccard rows and symmetric
       kNN-50 ANI, both bit-exact against oracle rows;
  (ii) config-4 shape: 200 000 references x 2 000 queries drawn from the SAME families (so that the top-50 are real
       relatives, not 50 exact zeros), cross kNN-50 core/accessory through reference shards (partial lists + merge)
       and in one piece, 64 sampled queries against the oracle.
The planes come from sketchlib.rust_b200.synth on the GPU (torch as an array library) and are copied to the host for
the oracle."""
import numpy as np
import pytest

import oracle
from test_gpu_parity import TOL, assert_knn_matches

pytestmark = pytest.mark.gpu
KMERS8 = [15, 17, 19, 21, 23, 25, 27, 29]


@pytest.fixture(scope="module")
def sk():
    import sketchlib.rust_b200 as sk
    return sk


@pytest.fixture(scope="module")
def ctx(sk):
    c = sk.Context(0)
    yield c
    c.close()


def device_db(sk, ctx, n, W, kmers, **kw):
    import torch
    from sketchlib.rust_b200 import synth
    bins = synth.make_bins(n, W, kmers, device="cuda", **kw)
    torch.cuda.synchronize()  # generated on torch's stream, read on the library's
    g = sk.SketchDb.upload(ctx, bins.data_ptr(), n, kmers, W, bins_loc=sk.SKD_DEVICE)
    host = synth.as_u64_numpy(bins)
    del bins
    torch.cuda.empty_cache()
    return g, host


def test_config5_shape_dense_jaccard_and_symmetric_knn_ani(sk, ctx):
    n, W, kmers = 20000, 160, [21]
    g, host = device_db(sk, ctx, n, W, kmers, seed=60)
    db = oracle.Db(host, n, kmers, W)
    dt = sk.set_k(g, 21)
    # dense Jaccard: row blocks at the start, across a tile edge and at the end of the triangle, bit for bit
    for r0, r1 in ((0, 3), (9982, 9986), (19990, 20000)):
        got = sk.self_dists_all(g, dist_type=dt, rows=(r0, r1)).distances
        first = int(oracle.lib().oracle_square_to_condensed(r0, r0 + 1, n))
        pad = first % 1000  # the oracle keeps the reference's 1000-distance chunks: aligned starts only
        ref = oracle.self_dists_all(db, k_idx=0, first=first - pad, count=got.size + pad, threads=16)[pad:]
        assert np.array_equal(got, ref)
    # the large sketch really exercises 16-bit counts near their limit: identical samples match in all 10 240 bins
    assert sk.match_counts(g, g, pairs=[(5, 5)])[0, 0] == 64 * W
    # symmetric kNN-50, ANI: sampled rows, keys bit-exact, neighbour sets equal outside exact ties
    got = sk.self_dists_knn(g, knn=50, dist_type=sk.set_k(g, 21, ani=True))
    assert (np.diff(got.d0, axis=1) <= 0).all()  # highest ANI first
    rng = np.random.default_rng(5)
    for s in sorted({0, n - 8} | {int(x) for x in rng.integers(0, n - 8, size=6)}):
        oidx, od0, _ = oracle.self_dists_knn(db, 50, k_idx=0, ani=True, threads=16, rows=(s, s + 8))
        assert np.array_equal(got.d0[s:s + 8], od0)
        for r in range(8):
            last = od0[r, -1]
            assert set(got.idx[s + r][got.d0[s + r] != last].tolist()) == set(oidx[r][od0[r] != last].tolist())
    g.free()


def test_config4_shape_cross_knn_same_families_through_ref_shards(sk, ctx):
    n_ref, n_q, knn, fams = 200_000, 2_000, 50, 4_000
    gr, ref_host = device_db(sk, ctx, n_ref, 16, KMERS8, seed=61, n_families=fams, member_seed=1)
    gq, q_host = device_db(sk, ctx, n_q, 16, KMERS8, seed=61, n_families=fams, member_seed=2)
    dt = sk.set_k(gr)
    whole = sk.cross_dists_knn(gr, gq, knn=knn, dist_type=dt)
    assert whole.idx.shape == (n_q, knn)
    assert (whole.d0[:, 0] < 1.0).mean() > 0.9          # the queries do have relatives among the references
    assert whole.d0.astype(np.float64).sum() > 1.0      # ... and the rows are not 50 exact zeros (r1's config-4 data)
    # the same through four reference shards streamed one after the other (partial lists + merge)
    per = ref_host.size // n_ref
    bounds = [0, 50_048, 100_000, 150_016, n_ref]
    parts = []
    for a, b in zip(bounds, bounds[1:]):
        shard = sk.SketchDb.upload(ctx, ref_host[a * per:b * per], b - a, KMERS8, 16)
        parts.append(sk.cross_knn_partial(shard, gq, knn=knn, dist_type=dt, ref_index_offset=a))
        shard.free()
    merged = sk.knn_merge(ctx, [p[0] for p in parts], [p[1] for p in parts], dt, None, None)
    assert np.array_equal(merged.d0, whole.d0) and np.array_equal(merged.d1, whole.d1)
    assert (merged.idx == whole.idx).mean() > 0.999     # ties inside a row may be ordered differently
    # 64 sampled queries against the oracle
    qs = np.sort(np.random.default_rng(6).choice(n_q, size=64, replace=False))
    rdb = oracle.Db(ref_host, n_ref, KMERS8, 16)
    qdb = oracle.Db(q_host.reshape(n_q, -1)[qs].reshape(-1), 64, KMERS8, 16)
    oidx, od0, od1 = oracle.cross_dists_knn(rdb, qdb, knn, threads=16)
    for res in (whole, merged):
        assert_knn_matches(res.idx[qs], res.d0[qs], res.d1[qs], oidx, od0, od1, False)
    assert np.abs(whole.d0[qs].astype(np.float64) - od0).max() <= TOL
    gr.free()
    gq.free()

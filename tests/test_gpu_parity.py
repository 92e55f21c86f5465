"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the CPU oracle and the
reference's golden outputs.  Run on the B200 box: python -m pytest tests -m gpu."""
import os

import numpy as np
import pytest

import oracle
from helpers import FIXTURES, fixture_db, load_json_db, subset_db, synth_db, synth_pair

pytestmark = pytest.mark.gpu

TOL = 1e-6  # BASELINE.json: Jaccard / core / accessory / ANI within 1e-6 absolute of the reference arithmetic


@pytest.fixture(scope="module")
def sk():
    import sketchlib.rust_b200 as sk
    return sk


@pytest.fixture(scope="module")
def ctx(sk):
    c = sk.Context(0)
    yield c
    c.close()


def up(sk, ctx, db):
    return sk.SketchDb.upload(ctx, db.bins, db.n, db.kmers, db.W, db.sketch_size, db.sample_stride, db.kmer_stride,
                              db.names)


def golden(name):
    return open(os.path.join(FIXTURES, name + ".stdout")).read()


def all_pairs(n, m=None):
    m = n if m is None else m
    return np.array([(i, j) for i in range(n) for j in range(m)], dtype=np.uint64)


# ------------------------------------------------------------------ integer parity (bit exact)
@pytest.mark.parametrize("n,W,kmers", [(200, 16, [15, 19, 23, 27]), (65, 3, [21]), (130, 160, [21]), (7, 1, [11, 13])])
def test_match_counts_bit_exact(sk, ctx, n, W, kmers):
    db = synth_db(n, W, kmers, seed=n + W)
    g = up(sk, ctx, db)
    dense = sk.match_counts(g, g)  # tiled kernel: [n, n, K]
    pairs = all_pairs(n)
    ref = oracle.match_counts(db, db, pairs).reshape(n, n, len(kmers))
    assert np.array_equal(dense, ref)
    sel = pairs[:: max(1, len(pairs) // 997)]
    assert np.array_equal(sk.match_counts(g, g, pairs=sel), oracle.match_counts(db, db, sel))  # pair kernel


def test_match_counts_known_answers(sk, ctx):
    big = load_json_db("s10048")
    g = up(sk, ctx, big)
    c = sk.match_counts(g, g, pairs=[(0, 1), (2, 3), (1, 3), (0, 2)])
    assert c[0].tolist() == [6811, 6547, 6310, 6127]
    assert c[1].tolist() == [7167, 6871, 6623, 6342]
    assert c[2].tolist() == [7, 1, 1, 3]
    d = sk.match_counts(g, g)
    assert d[0, 1].tolist() == [6811, 6547, 6310, 6127] and d[3, 1].tolist() == [7, 1, 1, 3]
    g3 = up(sk, ctx, fixture_db("sketches3"))
    assert sk.match_counts(g3, g3, pairs=[(0, 1), (2, 3)]).ravel().tolist() == [660, 722]


def test_cross_match_counts_ragged(sk, ctx):
    a, b = synth_pair(150, 70, 16, [17, 21, 25], seed=5, n_families=6)
    ga, gb = up(sk, ctx, a), up(sk, ctx, b)
    got = sk.match_counts(ga, gb, rows=(10, 140))
    ref = oracle.match_counts(a, b, all_pairs(150, 70)).reshape(150, 70, 3)[10:140]
    assert np.array_equal(got, ref)


# ------------------------------------------------------------------ the reference's golden stdout files
def test_golden_knn_ca_jaccard_ani(sk, ctx):
    db = load_json_db("s10048")
    g = up(sk, ctx, db)
    assert str(sk.self_dists_knn(g, knn=1, dist_type=sk.set_k(g))) == golden("dists_knn_ca")
    assert str(sk.self_dists_knn(g, knn=1, dist_type=sk.set_k(g, 21))) == golden("dists_knn_jaccard")
    assert str(sk.self_dists_knn(g, knn=1, dist_type=sk.set_k(g, 21, ani=True))) == golden("dists_knn_ani")


def test_golden_subset_dense(sk, ctx):
    db = load_json_db("s10048")
    names = [l.strip() for l in open(os.path.join(FIXTURES, "subset.txt")) if l.strip()]
    g = up(sk, ctx, subset_db(db, names))
    assert str(sk.self_dists_all(g, dist_type=sk.set_k(g))) == golden("dists_subset")


def test_golden_precluster_values(sk, ctx):
    g = up(sk, ctx, fixture_db("sketches3"))
    assert str(sk.self_dists_knn(g, knn=1, dist_type=sk.set_k(g, 21))) == golden("inverted_precluster")
    assert str(sk.self_dists_knn(g, knn=1, dist_type=sk.set_k(g, 21, ani=True))) == golden("inverted_precluster_ani")


# ------------------------------------------------------------------ dense layouts and values vs the oracle
def check_close(got, ref):
    assert got.shape == ref.shape
    assert np.isfinite(got).all()
    assert np.abs(got.astype(np.float64) - ref.astype(np.float64)).max() <= TOL


@pytest.mark.parametrize("n,W,kmers", [(300, 16, [15, 17, 19, 21, 23, 25, 27, 29]), (70, 2, [17, 21, 25]), (2, 16, [13, 17, 21])])
def test_dense_self_core_acc(sk, ctx, n, W, kmers):
    db = synth_db(n, W, kmers, seed=1, unrelated_frac=0.1)
    g = up(sk, ctx, db)
    got = sk.self_dists_all(g, dist_type=sk.set_k(g)).distances
    ref = oracle.self_dists_all(db, threads=8)
    check_close(got, ref)
    # the discrete outcomes (early break -> (1,1), exact zero core) must agree exactly
    assert np.array_equal(got == 1.0, ref == 1.0) and np.array_equal(got == 0.0, ref == 0.0)
    # and in fact the f64 epilogue is restated op for op: expect bit equality almost everywhere
    assert (got == ref).mean() > 0.999


@pytest.mark.parametrize("ani", [False, True])
def test_dense_self_single_k_bit_exact(sk, ctx, ani):
    db = synth_db(257, 16, [15, 21, 27], seed=2, unrelated_frac=0.05)
    g = up(sk, ctx, db)
    got = sk.self_dists_all(g, dist_type=sk.set_k(g, 21, ani=ani)).distances
    ref = oracle.self_dists_all(db, k_idx=1, ani=ani, threads=8)
    assert np.array_equal(got, ref)  # tables come from the host libm: identical f32


def test_dense_self_row_shards_concatenate(sk, ctx):
    db = synth_db(333, 16, [17, 21, 25, 29], seed=3)
    g = up(sk, ctx, db)
    dt = sk.set_k(g)
    whole = sk.self_dists_all(g, dist_type=dt).distances
    parts = [sk.self_dists_all(g, dist_type=dt, rows=r).distances for r in ((0, 64), (64, 100), (100, 101), (101, 333))]
    assert np.array_equal(np.concatenate(parts), whole)
    for world in (2, 3, 8):
        shards = [sk.triangle_row_partition(333, r, world) for r in range(world)]
        assert shards[0][0] == 0 and shards[-1][1] == 333
        cat = np.concatenate([sk.self_dists_all(g, dist_type=dt, rows=s).distances for s in shards])
        assert np.array_equal(cat, whole)


def test_dense_cross_layout_and_values(sk, ctx):
    a, b = synth_pair(130, 75, 16, [17, 21, 25, 29], seed=7, n_families=4)
    ga, gb = up(sk, ctx, a), up(sk, ctx, b)
    got = sk.cross_dists_all(ga, gb, dist_type=sk.set_k(ga))
    check_close(got.distances, oracle.cross_dists_all(a, b, threads=8))
    gotj = sk.cross_dists_all(ga, gb, dist_type=sk.set_k(ga, 21)).distances
    assert np.array_equal(gotj, oracle.cross_dists_all(a, b, k_idx=1, threads=8))
    part = sk.cross_dists_all(ga, gb, dist_type=sk.set_k(ga), rows=(60, 129)).distances
    assert np.array_equal(part, got.distances.reshape(130, 75, 2)[60:129].ravel())
    # text: ref-outer / query-inner (distance_matrix.rs:178-189)
    first = str(got).split("\n")[0].split("\t")
    assert first[0] == a.names[0] and first[1] == b.names[0] and len(first) == 4


def test_self_equals_cross_on_same_db(sk, ctx):
    db = synth_db(100, 16, [15, 19, 23, 27], seed=9)
    g = up(sk, ctx, db)
    s = sk.self_dists_all(g, dist_type=sk.set_k(g)).distances.reshape(-1, 2)
    x = sk.cross_dists_all(g, g, dist_type=sk.set_k(g)).distances.reshape(100, 100, 2)
    iu = np.triu_indices(100, 1)
    assert np.array_equal(x[iu], s) and np.array_equal(x.transpose(1, 0, 2)[iu], s)
    assert np.array_equal(x[np.arange(100), np.arange(100)], np.zeros((100, 2), np.float32))  # NaN r -> (0, 0)


# ------------------------------------------------------------------ degenerate and guard cases
def test_need_two_kmers_and_unknown_k(sk, ctx):
    g = up(sk, ctx, fixture_db("sketches3"))
    with pytest.raises(sk.SketchDistError, match="Need at least two k-mer lengths"):
        sk.self_dists_all(g, dist_type=sk.set_k(g))
    with pytest.raises(sk.SketchDistError, match="K-mer size 33 not found in file"):
        sk.set_k(g, 33)


def test_two_kmers_gives_ones(sk, ctx):
    leg = fixture_db("legacy_db")
    two = oracle.Db(leg.bins, 2, [17, 21], 2, 128, leg.sample_stride, leg.kmer_stride)
    g = up(sk, ctx, two)
    assert sk.self_dists_all(g, dist_type=sk.set_k(g)).distances.tolist() == [1.0, 1.0]
    g3 = up(sk, ctx, leg)  # legacy layout, K=3, 128 bins
    check_close(sk.self_dists_all(g3, dist_type=sk.set_k(g3)).distances, oracle.self_dists_all(leg))


def test_expected_samebits_nonzero(sk, ctx):
    db = synth_db(20, 256, [21, 31], seed=11, n_families=1)  # 16384 bins -> expected_samebits = 1
    g = up(sk, ctx, db)
    got = sk.self_dists_all(g, dist_type=sk.set_k(g, 21)).distances
    assert np.array_equal(got, oracle.self_dists_all(db, k_idx=0))
    gota = sk.self_dists_all(g, dist_type=sk.set_k(g, 31, ani=True)).distances
    assert np.array_equal(gota, oracle.self_dists_all(db, k_idx=1, ani=True))


def test_large_sketch_uses_32bit_counts(sk, ctx):
    db = synth_db(6, 1100, [21], seed=12, n_families=1)  # 70400 bins > 65535
    g = up(sk, ctx, db)
    assert np.array_equal(sk.match_counts(g, g), oracle.match_counts(db, db, all_pairs(6)).reshape(6, 6, 1))
    assert np.array_equal(sk.self_dists_all(g, dist_type=sk.set_k(g, 21)).distances, oracle.self_dists_all(db, k_idx=0))


def test_completeness(sk, ctx):
    db = synth_db(90, 16, [17, 21, 25, 29], seed=13, n_families=2)
    rng = np.random.default_rng(0)
    cv = rng.uniform(0.6, 1.0, size=90)
    g = up(sk, ctx, db)
    for kidx, ani in ((-1, False), (1, False), (1, True)):
        dt = sk.set_k(g) if kidx < 0 else sk.set_k(g, 21, ani=ani)
        got = sk.self_dists_all(g, dist_type=dt, completeness_vec=cv).distances
        ref = oracle.self_dists_all(db, k_idx=kidx, ani=ani, completeness=cv)
        check_close(got, ref)
        plain = sk.self_dists_all(g, dist_type=dt).distances
        assert not np.array_equal(plain, got)
    # below the cutoff nothing changes (jaccard.rs:37)
    low = np.full(90, 0.5)
    a = sk.self_dists_all(g, dist_type=sk.set_k(g, 21), completeness_vec=low).distances
    assert np.array_equal(a, sk.self_dists_all(g, dist_type=sk.set_k(g, 21)).distances)
    # cross: separate vectors per side
    b = oracle.Db(db.bins.reshape(90, -1)[rng.permutation(90)[:40]].reshape(-1), 40, db.kmers, db.W)  # related to db
    gb = up(sk, ctx, b)
    cq = rng.uniform(0.7, 1.0, size=40)
    got = sk.cross_dists_all(g, gb, dist_type=sk.set_k(g), ref_completeness_vec=cv, query_completeness_vec=cq).distances
    check_close(got, oracle.cross_dists_all(db, b, ref_c=cv, query_c=cq))


# ------------------------------------------------------------------ kNN
def assert_knn_matches(idx, d0, d1, oidx, od0, od1, exact_values):
    """Rows must hold the same distances in the same ascending order; neighbour sets may differ only
    inside groups of exactly tied keys (BASELINE.json: identical apart from exact ties).  For core/accessory
    (exact_values=False: keys agree within TOL, not bitwise) the sets are compared wherever the oracle's gap to
    its k-th key exceeds the tolerance: everything clearly below the boundary must be in both rows."""
    assert idx.shape == oidx.shape
    if exact_values:
        assert np.array_equal(d0, od0)
    else:
        assert np.abs(d0.astype(np.float64) - od0).max() <= TOL
    for r in range(idx.shape[0]):
        if exact_values:
            last = od0[r, -1]
            strict = od0[r] != last  # entries not tied with the boundary value must be the same set
            assert set(idx[r][d0[r] != last].tolist()) == set(oidx[r][strict].tolist())
        else:
            kth = float(od0[r, -1])
            ours, theirs = set(idx[r].tolist()), set(oidx[r].tolist())
            assert set(oidx[r][od0[r] < kth - 2 * TOL].tolist()) <= ours, f"row {r}: a clear neighbour is missing"
            assert set(idx[r][d0[r] < kth - 2 * TOL].tolist()) <= theirs, f"row {r}: a wrong neighbour below the boundary"
            # per neighbour (not per slot): the same sample must carry the same distances
            pos = {int(j): t for t, j in enumerate(oidx[r])}
            for t, j in enumerate(idx[r]):
                if int(j) in pos:
                    assert abs(float(d0[r, t]) - float(od0[r, pos[int(j)]])) <= TOL
                    if d1 is not None:
                        assert abs(float(d1[r, t]) - float(od1[r, pos[int(j)]])) <= TOL
        if d1 is not None:
            same = idx[r] == oidx[r]
            assert np.abs(d1[r][same].astype(np.float64) - od1[r][same]).max(initial=0.0) <= TOL


@pytest.mark.parametrize("knn", [1, 5, 50, 64])
def test_self_knn_jaccard(sk, ctx, knn):
    db = synth_db(500, 16, [15, 21, 27], seed=20, n_families=8, unrelated_frac=0.05)
    g = up(sk, ctx, db)
    got = sk.self_dists_knn(g, knn=knn, dist_type=sk.set_k(g, 21))
    oidx, od0, _ = oracle.self_dists_knn(db, knn, k_idx=1, threads=8)
    assert_knn_matches(got.idx, got.d0, None, oidx, od0, None, True)
    assert (np.diff(got.d0, axis=1) >= 0).all()
    assert (got.idx != np.arange(500)[:, None]).all()  # never its own neighbour


def test_self_knn_ani_roundtrip(sk, ctx):
    db = synth_db(300, 16, [15, 21, 27], seed=21, n_families=5)
    g = up(sk, ctx, db)
    got = sk.self_dists_knn(g, knn=10, dist_type=sk.set_k(g, 21, ani=True))
    oidx, od0, _ = oracle.self_dists_knn(db, 10, k_idx=1, ani=True, threads=8)
    assert np.array_equal(got.d0, od0)  # includes the f32 1-(1-ani) round trip (mod.rs:183-189)
    assert (np.diff(got.d0, axis=1) <= 0).all()  # highest ANI first


def test_self_knn_core_acc_and_row_shards(sk, ctx):
    db = synth_db(400, 16, [15, 17, 19, 21, 23, 25, 27, 29], seed=22, n_families=6, unrelated_frac=0.05)
    g = up(sk, ctx, db)
    got = sk.self_dists_knn(g, knn=50, dist_type=sk.set_k(g))
    oidx, od0, od1 = oracle.self_dists_knn(db, 50, threads=8)
    assert_knn_matches(got.idx, got.d0, got.d1, oidx, od0, od1, False)
    # the kNN rows must be the dense rows' smallest entries
    dense = sk.cross_dists_all(g, g, dist_type=sk.set_k(g)).distances.reshape(400, 400, 2)
    for r in (0, 63, 64, 399):
        row = dense[r, :, 0].copy()
        row[r] = np.inf
        assert np.array_equal(np.sort(row)[:50], got.d0[r])
        assert np.array_equal(dense[r, got.idx[r].astype(int), 1], got.d1[r])
    parts = [sk.self_dists_knn(g, knn=50, dist_type=sk.set_k(g), rows=rows) for rows in ((0, 100), (100, 130), (130, 400))]
    assert np.array_equal(np.concatenate([p.idx for p in parts]), got.idx)
    assert np.array_equal(np.concatenate([p.d0 for p in parts]), got.d0)


def test_cross_knn(sk, ctx):
    ref, qry = synth_pair(700, 90, 16, [17, 21, 25, 29], seed=23, n_families=10)
    gr, gq = up(sk, ctx, ref), up(sk, ctx, qry)
    got = sk.cross_dists_knn(gr, gq, knn=7, dist_type=sk.set_k(gr, 21))
    oidx, od0, _ = oracle.cross_dists_knn(ref, qry, 7, k_idx=1, threads=8)
    assert got.idx.shape == (90, 7)
    assert_knn_matches(got.idx, got.d0, None, oidx, od0, None, True)
    gotc = sk.cross_dists_knn(gr, gq, knn=7, dist_type=sk.set_k(gr))
    oidx, od0, od1 = oracle.cross_dists_knn(ref, qry, 7, threads=8)
    assert_knn_matches(gotc.idx, gotc.d0, gotc.d1, oidx, od0, od1, False)
    # rows are queries, columns refs (tests/distance.rs:465-490)
    line = str(gotc).split("\n")[0].split("\t")
    assert line[0] == qry.names[0] and line[1] in ref.names and len(line) == 4
    assert (gotc.d0 < 1.0).mean() > 0.5  # the query set really has relatives among the references
    # knn >= n_ref is clamped to n_ref and returns every ref (tests/distance.rs:665-687)
    small = synth_db(5, 16, [17, 21, 25, 29], seed=24, n_families=1)
    gs = up(sk, ctx, small)
    allr = sk.cross_dists_knn(gs, gq, knn=50, dist_type=sk.set_k(gs, 21))
    assert allr.knn == 5 and allr.idx.shape == (90, 5)
    assert all(sorted(r.tolist()) == [0, 1, 2, 3, 4] for r in allr.idx)


def test_knn_many_exact_ties(sk, ctx):
    # identical sketches in groups: lots of exactly tied keys; ordering must still be total and valid
    base = synth_db(12, 4, [21, 25, 29], seed=30, n_families=12)
    reps = np.repeat(np.arange(12), 20)
    bins = base.bins.reshape(12, -1)[reps].reshape(-1)
    db = oracle.Db(bins, 240, base.kmers, 4)
    g = up(sk, ctx, db)
    got = sk.self_dists_knn(g, knn=19, dist_type=sk.set_k(g, 21))
    assert (got.d0 == 0.0).all()  # the 19 other copies
    assert all(set(got.idx[r].tolist()) == set(np.nonzero(reps == reps[r])[0].tolist()) - {r} for r in range(240))
    oidx, od0, _ = oracle.self_dists_knn(db, 25, k_idx=0, threads=4)
    got25 = sk.self_dists_knn(g, knn=25, dist_type=sk.set_k(g, 21))
    assert np.array_equal(got25.d0, od0)


def test_knn_errors(sk, ctx):
    db = synth_db(10, 16, [17, 21, 25], seed=31)
    g = up(sk, ctx, db)
    with pytest.raises(sk.SketchDistError):
        sk.self_dists_knn(g, knn=10, dist_type=sk.set_k(g))  # knn must be <= n-1 (CLI clamps, lib.rs:379-382)
    got = sk.self_dists_knn(g, knn=9, dist_type=sk.set_k(g, 21))
    assert all(sorted(r.tolist()) == [x for x in range(10) if x != i] for i, r in enumerate(got.idx))
    empty = sk.SketchDb.upload(ctx, np.zeros(0, np.uint64), 0, [17, 21, 25], 16)
    with pytest.raises(sk.SketchDistError, match="Reference database has no loaded samples"):
        sk.cross_dists_knn(empty, g, knn=3, dist_type=sk.set_k(g, 21))
    with pytest.raises(sk.SketchDistError, match="Query database has no loaded samples"):
        sk.cross_dists_knn(g, empty, knn=3, dist_type=sk.set_k(g, 21))
    other = up(sk, ctx, synth_db(10, 8, [17, 21, 25], seed=32))
    with pytest.raises(sk.SketchDistError, match="differ in sketch size"):
        sk.cross_dists_all(g, other, dist_type=sk.set_k(g, 21))


# ------------------------------------------------------------------ loader
def test_load_skd_file_and_subset(sk, ctx, tmp_path):
    db = synth_db(150, 16, [17, 21, 25], seed=40)
    path = str(tmp_path / "db.skd")
    db.bins.astype("<u8").tofile(path)
    g_file = sk.SketchDb.load_skd(ctx, path, 150, db.kmers, 16)
    g_mem = up(sk, ctx, db)
    dt = sk.set_k(g_mem)
    assert np.array_equal(sk.self_dists_all(g_file, dist_type=dt).distances, sk.self_dists_all(g_mem, dist_type=dt).distances)
    order = [149, 3, 77, 5, 64, 0]
    g_sub = sk.SketchDb.load_skd(ctx, path, len(order), db.kmers, 16, sample_indices=order)
    sub = subset_db(db, [db.names[i] for i in order])
    check_close(sk.self_dists_all(g_sub, dist_type=dt).distances, oracle.self_dists_all(sub))
    with pytest.raises(sk.SketchDistError):
        sk.SketchDb.load_skd(ctx, path, 151, db.kmers, 16)
    with pytest.raises(sk.SketchDistError, match="Could not read from"):
        sk.SketchDb.load_skd(ctx, str(tmp_path / "missing.skd"), 1, db.kmers, 16)


def test_fixture_skd_files_direct(sk, ctx):
    ms = oracle.MultiSketch(os.path.join(FIXTURES, "sketches2")).read_sketch_data()
    g = sk.SketchDb.load_skd(ctx, os.path.join(FIXTURES, "sketches2.skd"), 4, ms.kmer_lengths, ms.sketchsize64,
                             ms.sketch_size, ms.sample_stride, ms.kmer_stride, ms.names())
    got = sk.self_dists_all(g, dist_type=sk.set_k(g, 31)).distances
    assert np.array_equal(got, oracle.self_dists_all(oracle.Db.from_multisketch(ms), k_idx=0))
    # tests/distance.rs:206-217 expects roughly [0.40755, 1, 1, 1, 1, 0.37049] (+-0.05) for this database
    assert np.abs(got - np.array([0.40755, 1, 1, 1, 1, 0.37049], np.float32)).max() < 0.05


# ------------------------------------------------------------------ scale: size-independent properties
def test_scale_properties(sk, ctx):
    n = 3000
    db = synth_db(n, 16, [15, 17, 19, 21, 23, 25, 27, 29], seed=50, n_families=40)
    g = up(sk, ctx, db)
    dm = sk.self_dists_all(g, dist_type=sk.set_k(g))
    d = dm.distances.reshape(-1, 2)
    assert d.shape[0] == n * (n - 1) // 2 and np.isfinite(d).all() and d.min() >= 0.0 and d.max() <= 1.0
    # spot rows against the oracle
    rng = np.random.default_rng(1)
    for i in rng.integers(0, n - 1, size=6):
        first = int(oracle.lib().oracle_square_to_condensed(int(i), int(i) + 1, n))
        cnt = n - 1 - int(i)
        ref = oracle.self_dists_all(db, first=first - first % 1000, count=cnt + first % 1000, threads=8)
        ref = ref[2 * (first % 1000):]
        check_close(d[first:first + cnt].ravel(), ref)
    # kNN agrees with the dense matrix at this size
    knn = sk.self_dists_knn(g, knn=50, dist_type=sk.set_k(g))
    full = np.full((n, n), np.inf, dtype=np.float32)
    iu = np.triu_indices(n, 1)
    full[iu] = d[:, 0]
    full.T[iu] = d[:, 0]
    assert np.array_equal(np.sort(full, axis=1)[:, :50], knn.d0)


def test_full_size_config2_properties(sk, ctx):
    """BASELINE.json configs[1] at FULL size (n=10 000, 1024 bins, 8 k-mer lengths; 49 995 000 pairs): the oracle
    cannot do the whole matrix in seconds, so the result is checked through size-independent properties -- a checksum
    of the bit patterns over rank-style row shards, symmetry through the cross driver, kNN against the dense matrix --
    plus whole oracle rows at sampled positions."""
    n, kmers = 10000, [15, 17, 19, 21, 23, 25, 27, 29]
    db = synth_db(n, 16, kmers, seed=2, n_families=200)
    g = up(sk, ctx, db)
    dt = sk.set_k(g)
    d = sk.self_dists_all(g, dist_type=dt).distances
    assert d.size == n * (n - 1) and np.isfinite(d).all() and d.min() >= 0.0 and d.max() <= 1.0
    bits = d.view(np.uint32).astype(np.uint64)
    whole = int(bits.sum())
    # (1) checksum of checksums over the row blocks 3 and 8 ranks would own
    for world in (3, 8):
        assert sum(int(sk.self_dists_all(g, dist_type=dt, rows=sk.triangle_row_partition(n, r, world)).distances
                       .view(np.uint32).astype(np.uint64).sum()) for r in range(world)) == whole
    # (2) oracle rows: first, last, block edges, random
    d2 = d.reshape(-1, 2)
    for i in (0, 63, 64, 4999, 9936, 9997, 9998):
        first = int(oracle.lib().oracle_square_to_condensed(i, i + 1, n))
        cnt = n - 1 - i
        ref = oracle.self_dists_all(db, first=first - first % 1000, count=cnt + first % 1000, threads=8)[2 * (first % 1000):]
        check_close(d2[first:first + cnt].ravel(), ref)
    # (3) the cross driver on (db, db) is the symmetric completion of the self result (sampled reference rows)
    for r0 in (0, 5120, 9936):
        x = sk.cross_dists_all(g, g, dist_type=dt, rows=(r0, r0 + 64)).distances.reshape(64, n, 2)
        for i in range(r0, r0 + 64):
            lo = int(oracle.lib().oracle_square_to_condensed(i, i + 1, n)) if i < n - 1 else 0
            assert np.array_equal(x[i - r0, i + 1:], d2[lo:lo + n - 1 - i])
            assert np.array_equal(x[i - r0, i], np.zeros(2, np.float32))
    # (4) kNN-50 core distances are the 50 smallest of every row of the dense matrix (ties: values only)
    knn = sk.self_dists_knn(g, knn=50, dist_type=dt)
    full = np.full((n, n), np.inf, dtype=np.float32)
    iu = np.triu_indices(n, 1)
    full[iu] = d2[:, 0]
    full.T[iu] = d2[:, 0]
    del iu
    full.partition(49, axis=1)
    assert np.array_equal(np.sort(full[:, :50], axis=1), knn.d0)
    # single-k Jaccard at full size: integer counts, so bit-exact rows and an idempotent second run
    dj = sk.self_dists_all(g, dist_type=sk.set_k(g, 21)).distances
    for i in (0, 7777):
        first = int(oracle.lib().oracle_square_to_condensed(i, i + 1, n))
        ref = oracle.self_dists_all(db, k_idx=3, first=first - first % 1000, count=n - 1 - i + first % 1000, threads=8)
        assert np.array_equal(dj[first:first + n - 1 - i], ref[first % 1000:])
    assert np.array_equal(dj, sk.self_dists_all(g, dist_type=sk.set_k(g, 21)).distances)


# ------------------------------------------------------------------ symmetric self kNN, partial lists, merges
def fresh_ctx(sk, **env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return sk.Context(0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("kind", ["ca", "jaccard", "ani"])
def test_symmetric_knn_equals_row_owned_knn(sk, ctx, kind):
    db = synth_db(700, 16, [15, 19, 23, 27], seed=70, n_families=9, unrelated_frac=0.05)
    plain = fresh_ctx(sk, SKD_KNN_SYMMETRIC=0)
    tiny = fresh_ctx(sk, SKD_SYM_CHUNK_MB=1)  # many row groups: exercises the list carry-over between groups
    try:
        outs = []
        for c in (ctx, plain, tiny):
            g = up(sk, c, db)
            dt = sk.set_k(g) if kind == "ca" else sk.set_k(g, 23, ani=(kind == "ani"))
            outs.append(sk.self_dists_knn(g, knn=20, dist_type=dt))
        for o in outs[1:]:  # same deterministic tie-break (neighbour index): identical, not just equivalent
            assert np.array_equal(o.idx, outs[0].idx) and np.array_equal(o.d0, outs[0].d0)
            if kind == "ca":
                assert np.array_equal(o.d1, outs[0].d1)
        oidx, od0, od1 = oracle.self_dists_knn(db, 20, k_idx=-1 if kind == "ca" else 2, ani=(kind == "ani"), threads=8)
        assert_knn_matches(outs[0].idx, outs[0].d0, outs[0].d1, oidx, od0, od1, kind != "ca")
    finally:
        os.environ.pop("SKD_SYM_CHUNK_MB", None)
        fresh_ctx(sk, SKD_SYM_CHUNK_MB=2048).close()  # restore the library-wide chunk size
        plain.close()
        tiny.close()


def test_self_knn_partials_merge_like_ranks(sk, ctx):
    db = synth_db(650, 16, [17, 21, 25, 29], seed=71, n_families=7)
    g = up(sk, ctx, db)
    for dt in (sk.set_k(g), sk.set_k(g, 21)):
        whole = sk.self_dists_knn(g, knn=30, dist_type=dt)
        for world in (2, 3):
            shards = [sk.triangle_row_partition(650, r, world) for r in range(world)]
            parts = [sk.self_knn_partial(g, knn=30, dist_type=dt, rows=s) for s in shards]
            merged = sk.knn_merge(ctx, [p[0] for p in parts], [p[1] for p in parts], dt, db.names)
            assert np.array_equal(merged.idx, whole.idx) and np.array_equal(merged.d0, whole.d0)
            if dt.kind == sk.CORE_ACC:
                assert np.array_equal(merged.d1, whole.d1)
    # beyond the symmetric fold's list size (64) the row blocks are evaluated against all columns: same merged rows
    dt = sk.set_k(g)
    whole = sk.self_dists_knn(g, knn=100, dist_type=dt)
    parts = [sk.self_knn_partial(g, knn=100, dist_type=dt, rows=sk.triangle_row_partition(650, r, 3)) for r in range(3)]
    merged = sk.knn_merge(ctx, [p[0] for p in parts], [p[1] for p in parts], dt, db.names)
    assert np.array_equal(merged.d0, whole.d0) and np.array_equal(merged.d1, whole.d1)


# ------------------------------------------------------------------ early exit of the dense core/accessory kernel
def test_early_exit_gives_the_same_results_and_skips_work(sk, ctx):
    """The dense kernel stops counting a warp's pair block after the third k-mer length when (almost) all of its
    pairs are final and finishes the few open ones pair by pair (defer_kernel); with the exit switched off every
    k-mer length of every pair is counted.  Both must give bit-identical outputs -- dense self, dense cross, rect keys
    under the symmetric kNN, with and without completeness vectors -- and equal the oracle."""
    kmers = [15, 17, 19, 21, 23, 25, 27, 29]
    db = synth_db(1500, 16, kmers, seed=90, n_families=40, unrelated_frac=0.3)
    ref, qry = synth_pair(900, 300, 16, kmers, seed=91, n_families=25, unrelated_frac=0.3)
    g, gr, gq = up(sk, ctx, db), up(sk, ctx, ref), up(sk, ctx, qry)
    rng = np.random.default_rng(3)
    cv = rng.uniform(0.7, 1.0, size=1500)

    def run_all():
        ctx.work_counters(reset=True)
        out = [sk.self_dists_all(g, dist_type=sk.set_k(g)).distances,
               sk.self_dists_all(g, dist_type=sk.set_k(g), completeness_vec=cv).distances,
               sk.cross_dists_all(gr, gq, dist_type=sk.set_k(gr)).distances]
        knn = sk.self_dists_knn(g, knn=30, dist_type=sk.set_k(g))
        return out + [knn.idx, knn.d0, knn.d1], ctx.work_counters()

    try:
        ctx.set_early_exit(False)
        full, (done_f, total_f, deferred_f) = run_all()
        ctx.set_early_exit(True)
        fast, (done, total, deferred) = run_all()
    finally:
        ctx.set_early_exit(True)
    for a, b in zip(full, fast):
        assert np.array_equal(a, b)
    assert done_f == total_f and deferred_f == 0          # nothing skipped with the exit off
    assert total == total_f and done < 0.8 * total         # ... a good part of the counting skipped with it on
    assert 0 < deferred < 0.05 * (1500 * 1499 // 2 * 3)   # a few per cent of the pairs went through defer_kernel
    check_close(fast[0], oracle.self_dists_all(db, threads=16))
    check_close(fast[1], oracle.self_dists_all(db, completeness=cv, threads=16))
    check_close(fast[2], oracle.cross_dists_all(ref, qry, threads=16))
    # a deferred-pair list that is far too small: warps that find it full keep counting, the result does not change
    tiny = fresh_ctx(sk, SKD_DEFER_CAP=50)
    gt = up(sk, tiny, db)
    assert np.array_equal(sk.self_dists_all(gt, dist_type=sk.set_k(gt)).distances, full[0])
    knn_t = sk.self_dists_knn(gt, knn=30, dist_type=sk.set_k(gt))
    assert np.array_equal(knn_t.idx, full[3]) and np.array_equal(knn_t.d0, full[4]) and np.array_equal(knn_t.d1, full[5])
    done_t, total_t, deferred_t = tiny.work_counters()
    assert done < done_t <= total_t and deferred_t <= 2 * 50 + 64   # (two launches, lists of 50)
    tiny.close()
    # one species: every pair is related, so only warps at the diagonal / the ragged edge (few valid pairs) leave early
    same = synth_db(700, 16, kmers, seed=92, n_families=1)
    gs = up(sk, ctx, same)
    ctx.work_counters(reset=True)
    d = sk.self_dists_all(gs, dist_type=sk.set_k(gs)).distances
    done, total, deferred = ctx.work_counters()
    assert done > 0.8 * total and deferred < 0.02 * (700 * 699 // 2)
    check_close(d, oracle.self_dists_all(same, threads=16))


# ------------------------------------------------------------------ the reference's envelope (no knn or k-mer limits)
def test_knn_beyond_list_capacity_and_up_to_n_minus_1(sk, ctx):
    """--knn is clamped to n - 1 and runs (src/lib.rs:379-382): knn = 2000 (more than the row-owned kernel's 1984
    shared-memory candidates) and knn = n - 1 (every other sample, sorted) go through dense keys + a full row sort."""
    n = 2300
    db = synth_db(n, 16, [15, 19, 23, 27], seed=80, n_families=20, unrelated_frac=0.05)
    g = up(sk, ctx, db)
    for knn in (2000, n - 1):
        got = sk.self_dists_knn(g, knn=knn, dist_type=sk.set_k(g, 19))
        assert got.idx.shape == (n, knn) and (np.diff(got.d0, axis=1) >= 0).all()
        assert (got.idx != np.arange(n)[:, None]).all()
        oidx, od0, _ = oracle.self_dists_knn(db, knn, k_idx=1, threads=16, rows=(0, 40))
        assert_knn_matches(got.idx[:40], got.d0[:40], None, oidx, od0, None, True)
        oidx, od0, _ = oracle.self_dists_knn(db, knn, k_idx=1, threads=16, rows=(n - 30, n))
        assert_knn_matches(got.idx[n - 30:], got.d0[n - 30:], None, oidx, od0, None, True)
    assert all(sorted(r.tolist()) == [x for x in range(n) if x != i] for i, r in enumerate(got.idx[:50]))
    # core/accessory, row range, and the ANI round trip on the same path
    got = sk.self_dists_knn(g, knn=2100, dist_type=sk.set_k(g), rows=(100, 164))
    oidx, od0, od1 = oracle.self_dists_knn(db, 2100, threads=16, rows=(100, 164))
    assert_knn_matches(got.idx, got.d0, got.d1, oidx, od0, od1, False)
    got = sk.self_dists_knn(g, knn=2000, dist_type=sk.set_k(g, 19, ani=True), rows=(0, 16))
    _, od0, _ = oracle.self_dists_knn(db, 2000, k_idx=1, ani=True, threads=16, rows=(0, 16))
    assert np.array_equal(got.d0, od0)
    # cross: knn larger than the list capacity, clamped to the number of references
    ref, qry = synth_pair(2100, 70, 16, [15, 19, 23, 27], seed=81, n_families=15)
    gr, gq = up(sk, ctx, ref), up(sk, ctx, qry)
    got = sk.cross_dists_knn(gr, gq, knn=5000, dist_type=sk.set_k(gr))
    assert got.knn == 2100 and got.idx.shape == (70, 2100)
    oidx, od0, od1 = oracle.cross_dists_knn(ref, qry, 5000, threads=16)
    assert_knn_matches(got.idx, got.d0, got.d1, oidx, od0, od1, False)


def test_thirty_kmer_lengths(sk, ctx):
    """The reference's regression loops over any number of k-mer lengths (src/distances/jaccard.rs:77).  30 of them:
    a 64 x 64 tile's counts no longer fit shared memory, so dense runs keep them in L2 (dense_kernel) and kNN runs go
    through the symmetric fold / dense keys."""
    kmers = list(range(5, 65, 2))
    assert len(kmers) == 30
    n = 400
    db = synth_db(n, 16, kmers, seed=82, n_families=5, mu_range=(0.0005, 0.004))
    g = up(sk, ctx, db)
    dt = sk.set_k(g)
    check_close(sk.self_dists_all(g, dist_type=dt).distances, oracle.self_dists_all(db, threads=16))
    got = sk.self_dists_knn(g, knn=20, dist_type=dt)
    oidx, od0, od1 = oracle.self_dists_knn(db, 20, threads=16)
    assert_knn_matches(got.idx, got.d0, got.d1, oidx, od0, od1, False)
    rows = sk.self_dists_knn(g, knn=20, dist_type=dt, rows=(64, 200))   # row-owned call: the sorted path
    assert np.array_equal(rows.d0, got.d0[64:200])
    ref, qry = synth_pair(300, 90, 16, kmers, seed=83, n_families=4, mu_range=(0.0005, 0.004))
    gr, gq = up(sk, ctx, ref), up(sk, ctx, qry)
    check_close(sk.cross_dists_all(gr, gq, dist_type=sk.set_k(gr)).distances, oracle.cross_dists_all(ref, qry, threads=16))
    gotc = sk.cross_dists_knn(gr, gq, knn=15, dist_type=sk.set_k(gr))
    oidx, od0, od1 = oracle.cross_dists_knn(ref, qry, 15, threads=16)
    assert_knn_matches(gotc.idx, gotc.d0, gotc.d1, oidx, od0, od1, False)
    # a handful of queries: too many k-mer lengths for the narrow shared-memory variants as well
    few = oracle.Db(qry.bins[:5 * qry.sample_stride], 5, kmers, 16)
    gf = up(sk, ctx, few)
    check_close(sk.cross_dists_all(gr, gf, dist_type=sk.set_k(gr)).distances, oracle.cross_dists_all(ref, few, threads=16))
    gotf = sk.cross_dists_knn(gr, gf, knn=15, dist_type=sk.set_k(gr))
    oidx, od0, od1 = oracle.cross_dists_knn(ref, few, 15, threads=16)
    assert_knn_matches(gotf.idx, gotf.d0, gotf.d1, oidx, od0, od1, False)
    # single k of the 30 and the raw counts
    c = sk.match_counts(g, g, pairs=[(0, 1), (5, 399)])
    assert np.array_equal(c, oracle.match_counts(db, db, [(0, 1), (5, 399)]))


def test_cross_knn_ref_shards_merge(sk, ctx):
    ref, qry = synth_pair(500, 130, 16, [17, 21, 25, 29], seed=72, n_families=6)
    gq = up(sk, ctx, qry)
    whole_ref = up(sk, ctx, ref)
    bounds = [0, 100, 101, 320, 500]  # ragged shards, one with a single reference
    for kidx in (-1, 1):
        dt = sk.set_k(whole_ref) if kidx < 0 else sk.set_k(whole_ref, 21)
        want = sk.cross_dists_knn(whole_ref, gq, knn=12, dist_type=dt)
        parts = []
        for a, b in zip(bounds, bounds[1:]):
            shard = oracle.Db(ref.bins[a * ref.sample_stride:b * ref.sample_stride], b - a, ref.kmers, ref.W)
            parts.append(sk.cross_knn_partial(up(sk, ctx, shard), gq, knn=12, dist_type=dt, ref_index_offset=a))
        merged = sk.knn_merge(ctx, [p[0] for p in parts], [p[1] for p in parts], dt, ref.names, qry.names)
        assert np.array_equal(merged.idx, want.idx) and np.array_equal(merged.d0, want.d0)
        if kidx < 0:
            assert np.array_equal(merged.d1, want.d1)


# ------------------------------------------------------------------ precluster (candidate-list kNN)
def test_precluster_goldens_and_modes(sk, ctx):
    db = fixture_db("sketches3")
    g = up(sk, ctx, db)
    skq = np.fromfile(os.path.join(FIXTURES, "inverted.skq"), dtype="<u2").reshape(4, 10)
    for knn in (1, 50):  # tests/inverted.rs:244-347: knn 50 prints the same rows (padding is skipped)
        got = sk.self_dists_knn_precluster(g, skq, db.names, knn, sk.set_k(g, 21))
        assert str(got) == golden("inverted_precluster")
        got = sk.self_dists_knn_precluster(g, skq, db.names, knn, sk.set_k(g, 21, ani=True))
        assert str(got) == golden("inverted_precluster_ani")
    # .ski order differs from .skd order: names drive the index remap (mod.rs:412-438)
    perm = [2, 0, 3, 1]
    got = sk.self_dists_knn_precluster(g, skq[perm], [db.names[i] for i in perm], 1, sk.set_k(g, 21))
    assert str(got) == golden("inverted_precluster")
    with pytest.raises(sk.SketchDistError, match="could not be found in the .ski"):
        sk.self_dists_knn_precluster(g, skq[:3], db.names[:3], 1, sk.set_k(g, 21))
    leg = up(sk, ctx, fixture_db("legacy_db"))
    with pytest.raises(sk.SketchDistError, match="Prefilter only available for single k-mer distances"):
        sk.self_dists_knn_precluster(leg, skq[:2], ["R6.fa.gz", "TIGR4.fa.gz"], 1, sk.set_k(leg))


@pytest.mark.parametrize("retain", [None, "singleton", "bruteforce"])
def test_precluster_vs_oracle(sk, ctx, retain):
    db = synth_db(300, 16, [17, 21, 25], seed=80, n_families=25, unrelated_frac=0.2)
    g = up(sk, ctx, db)
    rng = np.random.default_rng(5)
    # a coarse 8-bin "inverted" sketch per sample: family members share bins, loners share none
    fam = rng.integers(0, 40, size=300)
    skq = (fam[:, None] * 131 + np.arange(8)[None, :] * 7).astype(np.uint16)
    lone = rng.random(300) < 0.15
    skq[lone] = rng.integers(0, 1 << 16, size=(int(lone.sum()), 8), dtype=np.uint16)
    mut = rng.random((300, 8)) < 0.3
    skq[mut] = rng.integers(0, 1 << 16, size=int(mut.sum()), dtype=np.uint16)
    off, idx = oracle.any_shared_bins_csr(skq)
    for ani in (False, True):
        dt = sk.set_k(g, 21, ani=ani)
        got = sk.self_dists_knn_precluster(g, skq, db.names, 10, dt, retain_unmatched=retain)
        oidx, od0 = oracle.self_dists_knn_precluster(db, off, idx, 10, 1, ani=ani,
                                                      retain={None: 0, "singleton": 1, "bruteforce": 2}[retain])
        assert np.array_equal(got.d0, od0)
        pad = od0 == 1.0  # padding rows point at the row itself in both
        assert np.array_equal(got.idx[pad & (oidx == np.arange(300)[:, None])], oidx[pad & (oidx == np.arange(300)[:, None])])
        for r in range(300):
            strict = od0[r] != od0[r, -1]
            assert set(got.idx[r][strict].tolist()) == set(oidx[r][strict].tolist())
    cv = rng.uniform(0.7, 1.0, size=300)
    got = sk.self_dists_knn_precluster(g, skq, db.names, 5, sk.set_k(g, 21), completeness_vec=cv)
    oidx, od0 = oracle.self_dists_knn_precluster(db, off, idx, 5, 1, completeness=cv)
    assert np.abs(got.d0.astype(np.float64) - od0).max() <= TOL


# ------------------------------------------------------------------ few queries: narrow-column kernels, HBM-bound path
@pytest.mark.parametrize("nq", [1, 7, 8, 9, 16, 17, 32, 33, 64])
def test_cross_with_few_queries(sk, ctx, nq):
    ref, qry = synth_pair(330, nq, 16, [17, 21, 25, 29], seed=90, n_families=5)
    gr, gq = up(sk, ctx, ref), up(sk, ctx, qry)
    # raw counts through the (narrow) tiled kernel, refs x queries x K
    want = oracle.match_counts(ref, qry, all_pairs(330, nq)).reshape(330, nq, 4)
    assert np.array_equal(sk.match_counts(gr, gq), want)
    check_close(sk.cross_dists_all(gr, gq, dist_type=sk.set_k(gr)).distances, oracle.cross_dists_all(ref, qry, threads=4))
    assert np.array_equal(sk.cross_dists_all(gr, gq, dist_type=sk.set_k(gr, 21, ani=True)).distances,
                          oracle.cross_dists_all(ref, qry, k_idx=1, ani=True, threads=4))
    # kNN through dense keys + column select
    got = sk.cross_dists_knn(gr, gq, knn=11, dist_type=sk.set_k(gr, 21))
    oidx, od0, _ = oracle.cross_dists_knn(ref, qry, 11, k_idx=1, threads=4)
    assert_knn_matches(got.idx, got.d0, None, oidx, od0, None, True)
    gotc = sk.cross_dists_knn(gr, gq, knn=11, dist_type=sk.set_k(gr))
    oidx, od0, od1 = oracle.cross_dists_knn(ref, qry, 11, threads=4)
    assert_knn_matches(gotc.idx, gotc.d0, gotc.d1, oidx, od0, od1, False)


@pytest.mark.parametrize("n_ref,nq,knn", [(900, 40, 25), (20000, 3, 64), (5000, 8, 1), (3000, 1, 50), (6000, 70, 25),
                                          (3000, 600, 10)])
def test_few_queries_knn_equals_row_owned_path(sk, ctx, n_ref, nq, knn):
    """Whole-range cross kNN (knn <= 64) = dense keys in HBM + select_few_kernel (stream kernel up to 16 queries, tile
    kernel above) against the row-owned kernel; the 1 MB key buffer of the `tiny` context cuts the last two cases into
    4 and 14 reference chunks whose lists continue from chunk to chunk."""
    ref, qry = synth_pair(n_ref, nq, 16, [17, 21, 25, 29], seed=91, n_families=12)
    plain = fresh_ctx(sk, SKD_KNN_SYMMETRIC=0)  # also disables the few-queries route
    tiny = fresh_ctx(sk, SKD_SYM_CHUNK_MB=1)    # several reference chunks and row slices
    try:
        outs = []
        for c in (ctx, plain, tiny):
            gr, gq = up(sk, c, ref), up(sk, c, qry)
            outs.append((sk.cross_dists_knn(gr, gq, knn=knn, dist_type=sk.set_k(gr)),
                         sk.cross_dists_knn(gr, gq, knn=knn, dist_type=sk.set_k(gr, 21, ani=True))))
        for o in outs[1:]:
            for a, b in zip(o, outs[0]):
                # ties are broken by the neighbour index on every path (composite keys): exact equality
                assert np.array_equal(a.idx, b.idx) and np.array_equal(a.d0, b.d0)
                assert a.d1 is None or np.array_equal(a.d1, b.d1)
    finally:
        fresh_ctx(sk, SKD_SYM_CHUNK_MB=2048).close()
        plain.close()
        tiny.close()


@pytest.mark.parametrize("n_ref,nq,W,kmers", [(700, 3, 160, [21]), (700, 8, 160, [21]), (450, 2, 40, list(range(13, 31, 2))),
                                              (450, 5, 40, list(range(13, 31, 2))), (300, 1, 1, [11, 15]), (1000, 13, 16, [15, 19, 23, 27]),
                                              (95, 16, 3, [17, 21, 25])])
def test_stream_kernel_shapes(sk, ctx, n_ref, nq, W, kmers):
    """The streaming few-queries kernel over its envelope: large sketches (config 5 shape), a query side that does not
    fit in shared memory (9 k x 40 words x 5 queries: falls back to the tile kernel), K = 2, two passes (9-16
    queries), ragged last blocks, row ranges, completeness; counts bit-exact, distances against the oracle."""
    ref, qry = synth_pair(n_ref, nq, W, kmers, seed=123 + nq, n_families=4)
    gr, gq = up(sk, ctx, ref), up(sk, ctx, qry)
    K = len(kmers)
    want = oracle.match_counts(ref, qry, all_pairs(n_ref, nq)).reshape(n_ref, nq, K)
    assert np.array_equal(sk.match_counts(gr, gq), want)
    mid = K // 2
    wj = oracle.cross_dists_all(ref, qry, k_idx=mid, threads=4)
    assert np.array_equal(sk.cross_dists_all(gr, gq, dist_type=sk.set_k(gr, kmers[mid])).distances, wj)
    lo, hi = n_ref // 3 + 1, n_ref - 5   # a row range that starts and ends inside 32-sample half blocks
    part = sk.cross_dists_all(gr, gq, dist_type=sk.set_k(gr, kmers[mid]), rows=(lo, hi)).distances
    assert np.array_equal(part, wj.reshape(n_ref, nq)[lo:hi].ravel())
    if K >= 2:
        check_close(sk.cross_dists_all(gr, gq, dist_type=sk.set_k(gr)).distances, oracle.cross_dists_all(ref, qry, threads=4))
        rng = np.random.default_rng(nq)
        rc, qc = rng.uniform(0.7, 1.0, size=n_ref), rng.uniform(0.7, 1.0, size=nq)
        got = sk.cross_dists_all(gr, gq, dist_type=sk.set_k(gr), ref_completeness_vec=rc, query_completeness_vec=qc)
        check_close(got.distances, oracle.cross_dists_all(ref, qry, ref_c=rc, query_c=qc, threads=4))
    knn = min(7, n_ref)
    got = sk.cross_dists_knn(gr, gq, knn=knn, dist_type=sk.set_k(gr, kmers[mid], ani=True))
    oidx, od0, _ = oracle.cross_dists_knn(ref, qry, knn, k_idx=mid, ani=True, threads=4)
    assert_knn_matches(got.idx, got.d0, None, oidx, od0, None, True)


# ------------------------------------------------------------------ randomized differential sweep
def test_random_shapes_against_oracle(sk, ctx):
    rng = np.random.default_rng(2026)
    for trial in range(36):
        n = int(rng.integers(2, 180))
        nq = int(rng.integers(1, 90))
        W = int(rng.choice([1, 2, 3, 5, 16, 17]))
        K = int(rng.integers(1, 10))
        kmers = sorted(rng.choice(np.arange(5, 61), size=K, replace=False).tolist())
        a = synth_db(n, W, kmers, seed=1000 + trial, n_families=max(1, n // 12), unrelated_frac=0.1)
        b = synth_db(nq, W, kmers, seed=2000 + trial, n_families=max(1, n // 12))
        # make some queries related to refs: copy a few reference sketches into the query set
        take = rng.integers(0, n, size=max(1, nq // 3))
        bb = b.bins.reshape(nq, -1).copy()
        bb[: len(take)] = a.bins.reshape(n, -1)[take]
        b = oracle.Db(bb.reshape(-1), nq, kmers, W)
        ga, gb = up(sk, ctx, a), up(sk, ctx, b)
        tag = f"trial {trial}: n={n} nq={nq} W={W} kmers={kmers}"
        # integer parity, dense tile path
        assert np.array_equal(sk.match_counts(ga, gb), oracle.match_counts(a, b, all_pairs(n, nq)).reshape(n, nq, K)), tag
        kidx = int(rng.integers(0, K))
        ani = bool(rng.integers(0, 2))
        dtj = sk.set_k(ga, kmers[kidx], ani=ani)
        assert np.array_equal(sk.self_dists_all(ga, dist_type=dtj).distances, oracle.self_dists_all(a, k_idx=kidx, ani=ani)), tag
        assert np.array_equal(sk.cross_dists_all(ga, gb, dist_type=dtj).distances,
                              oracle.cross_dists_all(a, b, k_idx=kidx, ani=ani)), tag
        knn = int(rng.integers(1, min(n - 1, 70) + 1))
        got = sk.self_dists_knn(ga, knn=knn, dist_type=dtj)
        oidx, od0, _ = oracle.self_dists_knn(a, knn, k_idx=kidx, ani=ani)
        assert np.array_equal(got.d0, od0), tag
        gotx = sk.cross_dists_knn(ga, gb, knn=knn, dist_type=dtj)
        oidx, od0, _ = oracle.cross_dists_knn(a, b, knn, k_idx=kidx, ani=ani)
        assert np.array_equal(gotx.d0, od0), tag
        if K >= 2:
            dt = sk.set_k(ga)
            check_close(sk.self_dists_all(ga, dist_type=dt).distances, oracle.self_dists_all(a))
            check_close(sk.cross_dists_all(ga, gb, dist_type=dt).distances, oracle.cross_dists_all(a, b))
            got = sk.self_dists_knn(ga, knn=knn, dist_type=dt)
            oidx, od0, od1 = oracle.self_dists_knn(a, knn)
            assert_knn_matches(got.idx, got.d0, got.d1, oidx, od0, od1, False)
        ga.free()
        gb.free()


def test_knn_paths_with_completeness(sk, ctx):
    ref, qry = synth_pair(420, 50, 16, [17, 21, 25, 29], seed=97, n_families=6)
    rng = np.random.default_rng(3)
    rc, qc = rng.uniform(0.65, 1.0, size=420), rng.uniform(0.65, 1.0, size=50)
    gr, gq = up(sk, ctx, ref), up(sk, ctx, qry)
    # symmetric self kNN (single GPU) and its row-owned twin
    plain = fresh_ctx(sk, SKD_KNN_SYMMETRIC=0)
    try:
        gp = up(sk, plain, ref)
        for dtf in (lambda g: sk.set_k(g), lambda g: sk.set_k(g, 21), lambda g: sk.set_k(g, 21, ani=True)):
            a = sk.self_dists_knn(gr, knn=15, dist_type=dtf(gr), completeness_vec=rc)
            b = sk.self_dists_knn(gp, knn=15, dist_type=dtf(gp), completeness_vec=rc)
            assert np.array_equal(a.idx, b.idx) and np.array_equal(a.d0, b.d0)
            kidx, ani = (-1, False) if a.dist_type.kind == sk.CORE_ACC else (1, a.dist_type.kind == sk.ANI)
            oidx, od0, od1 = oracle.self_dists_knn(ref, 15, k_idx=kidx, ani=ani, completeness=rc, threads=4)
            assert_knn_matches(a.idx, a.d0, a.d1, oidx, od0, od1, False)
    finally:
        plain.close()
    # cross kNN, few-queries route: separate vectors per side (query = rows)
    got = sk.cross_dists_knn(gr, gq, knn=9, dist_type=sk.set_k(gr), ref_completeness_vec=rc, query_completeness_vec=qc)
    oidx, od0, od1 = oracle.cross_dists_knn(ref, qry, 9, ref_c=rc, query_c=qc, threads=4)
    assert_knn_matches(got.idx, got.d0, got.d1, oidx, od0, od1, False)
    nocomp = sk.cross_dists_knn(gr, gq, knn=9, dist_type=sk.set_k(gr))
    # a constant factor on J shifts ln J for every k alike: the slope (core) stays, the intercept (accessory) moves
    assert not np.array_equal(nocomp.d1, got.d1)


def test_async_upload_overlapped_dense_and_other_calls(sk, ctx):
    db = synth_db(1500, 16, [17, 21, 25, 29], seed=98, n_families=20)
    pin = ctx.pinned((db.bins.size,), np.uint64)
    pin.array[:] = db.bins
    want = sk.self_dists_all(up(sk, ctx, db), dist_type=sk.DistType(sk.CORE_ACC)).distances
    for rows in (None, (0, 700), (130, 1500), (900, 901)):
        g = sk.SketchDb.upload(ctx, pin.array, db.n, db.kmers, db.W, async_upload=True)
        got = sk.self_dists_all(g, dist_type=sk.set_k(g), rows=rows).distances  # pinned result: interleaved bands
        lo, hi = (0, 1500) if rows is None else rows
        first = int(ctx.lib.skd_self_pairs_in_rows(1500, 0, lo))
        cnt = int(ctx.lib.skd_self_pairs_in_rows(1500, lo, hi))
        assert np.array_equal(got, want[2 * first:2 * (first + cnt)])
        # the database is complete afterwards: any other call works on it
        assert np.array_equal(sk.self_dists_all(g, dist_type=sk.set_k(g, 21)).distances,
                              oracle.self_dists_all(db, k_idx=1, threads=8))
        g.free()
    # first call is not a dense self run: the planes are re-tiled up front
    g = sk.SketchDb.upload(ctx, pin.array, db.n, db.kmers, db.W, async_upload=True)
    a = sk.self_dists_knn(g, knn=10, dist_type=sk.set_k(g, 21))
    oidx, od0, _ = oracle.self_dists_knn(db, 10, k_idx=1, threads=8)
    assert np.array_equal(a.d0, od0)
    g.free()
    g = sk.SketchDb.upload(ctx, pin.array, db.n, db.kmers, db.W, async_upload=True)
    g.free()  # never used: freeing waits for the copies
    # pageable memory silently takes the synchronous path
    g = sk.SketchDb.upload(ctx, db.bins, db.n, db.kmers, db.W, async_upload=True)
    assert np.array_equal(sk.self_dists_all(g, dist_type=sk.set_k(g), pinned=False).distances, want)
    pin.free()


def test_python_file_level_api_reproduces_goldens(sk, ctx, tmp_path):
    """The reference's flow in Python: MultiSketch.load_metadata -> read_sketch_data[_block] -> distances -> Display,
    and Inverted.load -> precluster, on files, against the reference's golden stdout files."""
    from helpers import write_skm_skd
    prefix = write_skm_skd(str(tmp_path / "db"), load_json_db("s10048"))
    ms = sk.MultiSketch.load_metadata(prefix + ".skm")
    g = ms.read_sketch_data(ctx)
    assert g.names == ms.names() and g.number_samples_loaded() == 4
    assert str(sk.self_dists_knn(g, knn=1, dist_type=sk.set_k(g))) == golden("dists_knn_ca")
    assert str(sk.self_dists_knn(g, knn=1, dist_type=sk.set_k(g, 21))) == golden("dists_knn_jaccard")
    sub = sk.MultiSketch(prefix)
    gs = sub.read_sketch_data_block(ctx, sk.read_subset_names(os.path.join(FIXTURES, "subset.txt")))
    assert str(sk.self_dists_all(gs, dist_type=sk.set_k(gs))) == golden("dists_subset")
    for ski in ("inverted", "reordered_inv"):
        for ani, gold in ((False, "inverted_precluster"), (True, "inverted_precluster_ani")):
            got = sk.precluster_from_files(ctx, os.path.join(FIXTURES, ski + ".ski"), os.path.join(FIXTURES, "sketches3"),
                                           knn=1, ani=ani)
            assert sorted(str(got).splitlines()) == sorted(golden(gold).splitlines())
    with pytest.raises(sk.HostError, match="could not find|Could not find"):
        sk.MultiSketch(prefix).read_sketch_data_block(ctx, ["not_a_sample"])

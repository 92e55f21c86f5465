"""Pins the CPU oracle (oracle/) against every golden the reference's own tests hold for the
distance path.  CPU only."""
import os

import numpy as np
import pytest

import oracle
from helpers import FIXTURES, fixture_db, load_json_db, render_dense_self, render_sparse, subset_db


def golden(name):
    return open(os.path.join(FIXTURES, name + ".stdout")).read()


def test_skm_fixtures_decode():
    for nm, ss, w, ks in (("sketches1", 1024, 16, [31]), ("sketches2", 10048, 157, [31]),
                          ("sketches3", 1024, 16, [21]), ("legacy_db", 128, 2, [17, 21, 25])):
        ms = oracle.MultiSketch(os.path.join(FIXTURES, nm)).read_sketch_data()
        assert (ms.sketch_size, ms.sketchsize64, ms.kmer_lengths) == (ss, w, ks)
        assert ms.bins.size == ms.number_samples_loaded() * len(ks) * w * 14


def test_known_answer_match_counts():
    # SURVEY.md header: per-k counts found from the reference's fixtures
    db3 = fixture_db("sketches3")
    assert oracle.match_counts(db3, db3, [(0, 1), (2, 3)]).ravel().tolist() == [660, 722]
    db1 = fixture_db("sketches1")
    assert oracle.match_counts(db1, db1, [(0, 1), (2, 3)]).ravel().tolist() == [597, 678]
    big = load_json_db("s10048")
    c = oracle.match_counts(big, big, [(0, 1), (2, 3), (1, 3), (0, 2)])
    assert c[0].tolist() == [6811, 6547, 6310, 6127]
    assert c[1].tolist() == [7167, 6871, 6623, 6342]
    assert c[2].tolist() == [7, 1, 1, 3]
    assert c[3].tolist()[:2] == [12, 0]
    leg = fixture_db("legacy_db")
    assert oracle.match_counts(leg, leg, [(0, 1)]).ravel().tolist() == [98, 100, 82]


def test_inverted_precluster_goldens():
    db = fixture_db("sketches3")  # k=21, 1024 bins; same values as the dense kNN path
    idx, d0, _ = oracle.self_dists_knn(db, 1, k_idx=0)
    assert render_sparse(db.names, db.names, idx, d0) == golden("inverted_precluster")
    idx, d0, _ = oracle.self_dists_knn(db, 1, k_idx=0, ani=True)
    assert render_sparse(db.names, db.names, idx, d0) == golden("inverted_precluster_ani")


def test_dists_knn_goldens():
    db = load_json_db("s10048")  # --k-seq 17,31,4 -s 10000, tests/distance.rs:281-301
    idx, d0, d1 = oracle.self_dists_knn(db, 1)
    assert render_sparse(db.names, db.names, idx, d0, d1) == golden("dists_knn_ca")
    idx, d0, _ = oracle.self_dists_knn(db, 1, k_idx=db.k_idx(21))
    assert render_sparse(db.names, db.names, idx, d0) == golden("dists_knn_jaccard")
    idx, d0, _ = oracle.self_dists_knn(db, 1, k_idx=db.k_idx(21), ani=True)
    assert render_sparse(db.names, db.names, idx, d0) == golden("dists_knn_ani")


def test_dists_subset_golden():
    db = load_json_db("s10048")
    names = [l.strip() for l in open(os.path.join(FIXTURES, "subset.txt")) if l.strip()]
    sub = subset_db(db, names)
    out = oracle.self_dists_all(sub)
    assert render_dense_self(sub.names, out, 2) == golden("dists_subset")


def test_index_maths_roundtrip():
    L = oracle.lib()
    for n in (2, 3, 4, 17, 1000, 10000):
        total = n * (n - 1) // 2
        for d in sorted({0, 1, total // 3, total // 2, total - 2, total - 1} & set(range(total))):
            i = L.oracle_calc_row_idx(d, n)
            j = L.oracle_calc_col_idx(d, i, n)
            assert i < j < n
            assert L.oracle_square_to_condensed(i, j, n) == d


def test_threads_agree_and_layouts():
    db = load_json_db("s1024")
    a = oracle.self_dists_all(db, threads=1)
    b = oracle.self_dists_all(db, threads=4)
    assert np.array_equal(a, b) and a.size == 12
    # cross (ref-major) against self: out[i*nq + j]
    x = oracle.cross_dists_all(db, db).reshape(4, 4, 2)
    d = 0
    for i in range(4):
        for j in range(i + 1, 4):
            assert np.array_equal(x[i, j], a[2 * d:2 * d + 2])
            assert np.array_equal(x[j, i], a[2 * d:2 * d + 2])
            d += 1
    # identical sketches: every y is exactly 0 -> r = NaN -> (0, 0)   (jaccard.rs:127-140)
    assert np.array_equal(x[0, 0], [0.0, 0.0])


def test_degenerate_cases():
    # K = 2: n < 3 -> (1, 1); K = 1 core/acc is an error (reference panics, jaccard.rs:70-72)
    leg = fixture_db("legacy_db")
    two = oracle.Db(leg.bins, 2, [17, 21], 2, 128, leg.sample_stride, leg.kmer_stride)
    assert oracle.self_dists_all(two).tolist() == [1.0, 1.0]
    one = fixture_db("sketches3")
    with pytest.raises(RuntimeError):
        oracle.self_dists_all(one)
    # completeness: correction applied above the cutoff, capped at 1, skipped below it
    j = oracle.lib().oracle_jaccard_from_samebits
    base = j(660, 16, 0, 0.0, 0.0, 0.64)
    assert base == 660 / 1024
    c1, c2 = 0.8, 0.9
    assert j(660, 16, 1, c1, c2, 0.64) == pytest.approx(base / (c1 * c2 / (c1 + c2 - c1 * c2)), abs=1e-15)
    assert j(660, 16, 1, 0.5, 0.5, 0.64) == base
    assert j(1000, 16, 1, 0.7, 0.95, 0.64) == 1.0
    # expected_samebits > 0 needs >= 16384 bins (W = 256 -> expected = 1)
    assert j(1, 256, 0, 0, 0, 0.64) == 0.0
    assert j(16384, 256, 0, 0, 0, 0.64) == 1.0


def test_precluster_driver_goldens():
    # tests/inverted.rs:244-347: index built with -k 21 -s 10 (inverted.skq), distances from the standard
    # k=21 / 1024-bin database (= the shipped sketches3), knn 1 and knn 50 print the same rows (padding skipped)
    db = fixture_db("sketches3")
    skq = np.fromfile(os.path.join(FIXTURES, "inverted.skq"), dtype="<u2").reshape(4, 10)
    off, idx = oracle.any_shared_bins_csr(skq)
    cand = [idx[off[i]:off[i + 1]].tolist() for i in range(4)]
    assert cand == [[0, 1], [0, 1], [2, 3], [2, 3]]  # "Identified 2 prefilter pairs from a max of 6"
    for knn in (1, 3):
        nidx, d0 = oracle.self_dists_knn_precluster(db, off, idx, knn, 0)
        assert render_sparse(db.names, db.names, nidx, d0) == golden("inverted_precluster")
        nidx, d0 = oracle.self_dists_knn_precluster(db, off, idx, knn, 0, ani=True)
        assert render_sparse(db.names, db.names, nidx, d0) == golden("inverted_precluster_ani")
    # no candidates: padded with (i, 1.0) and not printed; singleton keeps (i, 0); bruteforce = plain kNN
    none_off, none_idx = np.arange(5, dtype=np.uint64), np.arange(4, dtype=np.uint64)  # only itself
    nidx, d0 = oracle.self_dists_knn_precluster(db, none_off, none_idx, 2, 0)
    assert (nidx == np.arange(4)[:, None]).all() and (d0 == 1.0).all()
    assert render_sparse(db.names, db.names, nidx, d0) == ""
    nidx, d0 = oracle.self_dists_knn_precluster(db, none_off, none_idx, 2, 0, retain=1)
    assert d0.tolist() == [[0.0, 1.0]] * 4 and (nidx == np.arange(4)[:, None]).all()
    nidx, d0 = oracle.self_dists_knn_precluster(db, none_off, none_idx, 2, 0, retain=2)
    kidx, kd0, _ = oracle.self_dists_knn(db, 2, k_idx=0)
    assert np.array_equal(nidx, kidx) and np.array_equal(d0, kd0)

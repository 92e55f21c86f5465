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


# ---------------------------------------------------------------- second opinion: an independent restatement
def _py_bins(db, s, k_idx):
    """14-bit bin values of sample s at k index k_idx, rebuilt from the bit planes (inverse of fill_usigs,
    src/sketch/mod.rs:215-223)."""
    sl = db.bins[s * db.sample_stride + k_idx * db.kmer_stride:][:db.W * 14].reshape(db.W, 14)
    bits = (sl[:, :, None] >> np.arange(64, dtype=np.uint64)[None, None, :]) & np.uint64(1)    # [w][plane][bin]
    return (bits * (np.uint64(1) << np.arange(14, dtype=np.uint64))[None, :, None]).sum(axis=1).reshape(-1)


def _py_jaccard(a_bins, b_bins, W, c1=None, c2=None, cutoff=0.64):
    """jaccard.rs:6-45 on bin VALUES (equal bins <=> all 14 planes equal)."""
    samebits = int((a_bins == b_bins).sum())
    maxnbits = 64 * W
    expected = maxnbits >> 14
    diff = max(0, samebits - expected)
    intersize = diff * maxnbits / (maxnbits - expected)
    j = intersize / (64 * W)
    if c1 is not None and c2 is not None and c1 * c2 >= cutoff:
        j = min(1.0, j / (c1 * c2 / (c1 + c2 - c1 * c2)))
    return j


def _py_core_acc(db, i, j, comp=None):
    """core_acc_dist + simple_linear_regression, jaccard.rs:61-142, in Python floats (IEEE double, same libm)."""
    import math
    tol = math.log(2.0 / (db.sketch_size * 64.0))
    xsum = ysum = xysum = xsq = ysq = 0.0
    n = 0.0
    for k_idx, k in enumerate(db.kmers.tolist()):
        c1, c2 = (None, None) if comp is None else (comp[i], comp[j])
        jac = _py_jaccard(_py_bins(db, i, k_idx), _py_bins(db, j, k_idx), db.W, c1, c2)
        y = math.log(jac) if jac > 0 else -math.inf
        if y < tol:
            break
        x = float(k)
        xsum += x
        ysum += y
        xysum += x * y
        xsq += x * x
        ysq += y * y
        n += 1.0
    if math.isnan(ysum) or ysum == -math.inf or n < 3:
        return 1.0, 1.0
    xbar, ybar = xsum / n, ysum / n
    x_diff, y_diff = xsq - xsum * xsum / n, ysq - ysum * ysum / n
    xstddev, ystddev = math.sqrt(x_diff / n), math.sqrt(y_diff / n)
    den = math.sqrt(x_diff * y_diff)
    num = xysum - xsum * ysum / n
    r = num / den if den != 0 else (math.nan if num == 0 else math.copysign(math.inf, num))
    beta = r * ystddev / xstddev
    alpha = -beta * xbar + ybar
    if math.isnan(r):
        core = 0.0          # `beta < 0` and `r > 0` are both false for NaN
    else:
        core = 1.0 - math.exp(beta) if beta < 0 else (1.0 if r > 0 else 0.0)
    acc = 1.0 - math.exp(alpha) if alpha < 0 else 0.0
    return core, acc


@pytest.mark.parametrize("n,W,kmers,seed", [(9, 2, [13, 17, 21, 25, 29], 1), (6, 16, [15, 19, 23], 2), (4, 256, [17, 21, 25, 29], 3),
                                            (7, 1, [11, 13, 15, 17], 4)])
def test_oracle_agrees_with_independent_python_restatement(n, W, kmers, seed):
    """The C oracle against a second, structurally different restatement (bin values rebuilt from the planes, Python
    float arithmetic): bit-identical f32 distances for core/accessory, Jaccard, ANI and completeness, including
    >= 16 384 bins (expected_samebits = 1), early breaks, n < 3 and constant-y pairs."""
    import math
    from helpers import synth_db
    db = synth_db(n, W, kmers, seed=seed, n_families=2)
    rows = db.bins.reshape(n, -1).copy()
    rows[1] = rows[0]                       # identical pair: y = 0 at every k -> NaN r -> (0, 0)
    rows[2, :W * 14] ^= np.uint64(0xFFFFFFFFFFFFFFFF)   # sample 2 shares nothing at the first k -> break at once -> (1, 1)
    db = oracle.Db(rows.reshape(-1), n, kmers, W)
    comp = np.linspace(0.7, 1.0, n)
    got = oracle.self_dists_all(db).reshape(-1, 2)
    gotc = oracle.self_dists_all(db, completeness=comp, cutoff=0.64).reshape(-1, 2)
    mid = len(kmers) // 2
    gotj = oracle.self_dists_all(db, k_idx=mid)
    gota = oracle.self_dists_all(db, k_idx=mid, ani=True)
    d = 0
    for i in range(n):
        for j in range(i + 1, n):
            assert tuple(got[d]) == tuple(np.float32(v) for v in _py_core_acc(db, i, j)), (i, j)
            assert tuple(gotc[d]) == tuple(np.float32(v) for v in _py_core_acc(db, i, j, comp)), (i, j)
            jac = _py_jaccard(_py_bins(db, i, mid), _py_bins(db, j, mid), W)
            assert gotj[d] == np.float32(1.0 - jac)
            ani = max(0.0, 1.0 + (1.0 / kmers[mid]) * math.log(2.0 * jac / (1.0 + jac))) if jac > 0 else 0.0
            assert gota[d] == np.float32(ani)
            d += 1
    assert tuple(got[0]) == (0.0, 0.0) and tuple(got[1]) == (1.0, 1.0)

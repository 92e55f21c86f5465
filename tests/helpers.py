"""Shared test helpers: fixture loading, synthetic sketch generation, text rendering."""
import json
import os

import numpy as np

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
FIXTURES = os.path.join(GOLDEN, "ref_fixtures")


def load_json_db(tag):
    """Load tests/golden/rfile_k17-29_<tag>.{skd,json} as an oracle.Db."""
    base = os.path.join(GOLDEN, f"rfile_k17-29_{tag}")
    meta = json.load(open(base + ".json"))
    bins = np.fromfile(base + ".skd", dtype="<u8")
    return oracle.Db(bins, len(meta["names"]), meta["kmer_lengths"], meta["sketchsize64"], meta["sketch_size"],
                     meta["sample_stride"], meta["kmer_stride"], meta["names"])


def subset_db(db, names):
    """read_sketch_data_block semantics: rows in the order of `names`."""
    rows = [db.names.index(n) for n in names]
    bins = np.concatenate([db.bins[r * db.sample_stride:(r + 1) * db.sample_stride] for r in rows])
    return oracle.Db(bins, len(rows), db.kmers, db.W, db.sketch_size, db.sample_stride, db.kmer_stride, list(names))


def fixture_db(name):
    return oracle.Db.from_multisketch(oracle.MultiSketch(os.path.join(FIXTURES, name)).read_sketch_data())


def synth_db(n, W, kmers, seed=0, n_families=None, mu_range=(0.001, 0.03), acc_range=(0.0, 0.3), unrelated_frac=0.0):
    """Sketch-level synthetic families: per family and k an ancestor draws its
    bins uniformly in [0, 2^14); each member keeps a bin with probability (1-a)*exp(-mu*k), else redraws."""
    rng = np.random.default_rng(seed)
    kmers = list(kmers)
    nbins = 64 * W
    n_families = n_families or max(1, n // 50)
    fam = rng.integers(0, n_families, size=n)
    anc = rng.integers(0, 1 << 14, size=(n_families, len(kmers), nbins), dtype=np.uint16)
    mu = rng.uniform(*mu_range, size=n)
    acc = rng.uniform(*acc_range, size=n)
    bins = np.empty((n, len(kmers), nbins), dtype=np.uint16)
    for ki, k in enumerate(kmers):
        keep_p = (1.0 - acc) * np.exp(-mu * k)
        keep = rng.random((n, nbins)) < keep_p[:, None]
        redraw = rng.integers(0, 1 << 14, size=(n, nbins), dtype=np.uint16)
        bins[:, ki, :] = np.where(keep, anc[fam, ki, :], redraw)
    if unrelated_frac > 0:
        lone = rng.random(n) < unrelated_frac
        bins[lone] = rng.integers(0, 1 << 14, size=(int(lone.sum()), len(kmers), nbins), dtype=np.uint16)
    usigs = oracle.fill_usigs(bins)  # (n, K, W*14)
    return oracle.Db(usigs.reshape(-1), n, kmers, W)


def render_dense_self(names, dists, ncols):
    lines, d = [], 0
    for i in range(len(names)):
        for j in range(i + 1, len(names)):
            vals = "\t".join(oracle.fmt_f32(dists[d * ncols + c]) for c in range(ncols))
            lines.append(f"{names[i]}\t{names[j]}\t{vals}\n")
            d += 1
    return "".join(lines)


def render_sparse(row_names, col_names, idx, d0, d1=None):
    lines = []
    for r, rn in enumerate(row_names):
        for t in range(idx.shape[1]):
            cn = col_names[int(idx[r, t])]
            if d1 is None:
                if d0[r, t] < 1.0 or cn != rn:  # padding rule, distance_matrix.rs:376-381
                    lines.append(f"{rn}\t{cn}\t{oracle.fmt_f32(d0[r, t])}\n")
            else:
                lines.append(f"{rn}\t{cn}\t{oracle.fmt_f32(d0[r, t])}\t{oracle.fmt_f32(d1[r, t])}\n")
    return "".join(lines)

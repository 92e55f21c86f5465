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


def synth_pair(n_ref, n_query, W, kmers, seed=0, n_families=None, **kw):
    """A reference and a query database drawn from the SAME families (one synth_db split in two), so that cross
    comparisons see related pairs, not only unrelated ones."""
    both = synth_db(n_ref + n_query, W, kmers, seed=seed, n_families=n_families, **kw)
    rows = both.bins.reshape(n_ref + n_query, -1)
    rng = np.random.default_rng(seed + 12345)
    perm = rng.permutation(n_ref + n_query)  # family members end up on both sides
    ref = oracle.Db(rows[perm[:n_ref]].reshape(-1), n_ref, kmers, W, names=[f"ref{i}" for i in range(n_ref)])
    qry = oracle.Db(rows[perm[n_ref:]].reshape(-1), n_query, kmers, W, names=[f"qry{i}" for i in range(n_query)])
    return ref, qry


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


# ---------------------------------------------------------------- .skm writer (tests only)
def _cbor(obj):
    def head(major, val):
        if val < 24:
            return bytes([(major << 5) | val])
        for info, nb in ((24, 1), (25, 2), (26, 4), (27, 8)):
            if val < (1 << (8 * nb)):
                return bytes([(major << 5) | info]) + val.to_bytes(nb, "big")
        raise ValueError(val)
    if isinstance(obj, bool):
        return bytes([0xF5 if obj else 0xF4])
    if isinstance(obj, int):
        return head(0, obj) if obj >= 0 else head(1, -1 - obj)
    if isinstance(obj, str):
        b = obj.encode()
        return head(3, len(b)) + b
    if isinstance(obj, (list, tuple)):
        return head(4, len(obj)) + b"".join(_cbor(x) for x in obj)
    if isinstance(obj, dict):
        return head(5, len(obj)) + b"".join(_cbor(k) + _cbor(v) for k, v in obj.items())
    raise TypeError(type(obj))


def _crc32c(data: bytes) -> int:
    table = getattr(_crc32c, "table", None)
    if table is None:
        table = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            table.append(c)
        _crc32c.table = table
    crc = 0xFFFFFFFF
    for b in data:
        crc = table[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def write_skm_skd(prefix, db, version="0.3.0", name_map=None, indices=None):
    """Writes <prefix>.skm/.skd for an oracle.Db in the reference's container format (snappy frame with
    uncompressed chunks + CBOR map with the serde field names of MultiSketch / Sketch)."""
    meta = {
        "sketch_size": db.sketch_size, "sketchsize64": db.W, "kmer_lengths": [int(k) for k in db.kmers],
        "sketch_metadata": [
            {"name": nm, "index": i if indices is None else int(indices[i]), "rc": True, "reads": False, "seq_length": 2000000, "densified": False,
             "acgt": [1, 1, 1, 1], "non_acgt": 0} for i, nm in enumerate(db.names)],
        "name_map": name_map if name_map is not None else {nm: i for i, nm in enumerate(db.names)},
        "bin_stride": 1, "kmer_stride": db.kmer_stride, "sample_stride": db.sample_stride,
        "sketch_version": version, "hash_type": "DNA",
    }
    payload = _cbor(meta)
    out = bytearray(b"\xff\x06\x00\x00sNaPpY")
    for off in range(0, len(payload), 65536):
        chunk = payload[off:off + 65536]
        crc = _crc32c(chunk)
        masked = (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF
        body = masked.to_bytes(4, "little") + chunk
        out += b"\x01" + len(body).to_bytes(3, "little") + body
    with open(prefix + ".skm", "wb") as fh:
        fh.write(bytes(out))
    db.bins[: db.n * db.sample_stride].astype("<u8").tofile(prefix + ".skd")
    return prefix

"""TEST INFRASTRUCTURE ONLY: ctypes front-end to the CPU oracle (oracle/*.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
legs may import this package.  The product (sketchlib.rust_b200) never does.
"""
import ctypes as C
import gzip
import os
import subprocess

import numpy as np

from .skm import MultiSketch  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def _cpu_stamp() -> str:
    """The oracle is compiled with -march=native: a library built on another CPU (the build container vs the
    GPU box) must be rebuilt, or it may use instructions this host lacks."""
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("flags"):
                    import hashlib
                    return hashlib.sha1(line.encode()).hexdigest()
    except OSError:
        pass
    return "unknown"


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("sketchdist_oracle.c", "sketcher_oracle.c", "Makefile")]
    stamp_path = _LIB_PATH + ".cpu"
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs
    )
    try:
        other_cpu = open(stamp_path).read().strip() != _cpu_stamp()
    except OSError:
        other_cpu = True
    if force or stale or other_cpu:
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "liboracle.so"])
        with open(stamp_path, "w") as fh:
            fh.write(_cpu_stamp())
    return _LIB_PATH


class _Db(C.Structure):
    _fields_ = [
        ("bins", C.c_void_p),
        ("n_samples", C.c_uint64),
        ("sample_stride", C.c_uint64),
        ("kmer_stride", C.c_uint64),
        ("sketchsize64", C.c_uint64),
        ("sketch_size", C.c_uint64),
        ("kmers", C.c_void_p),
        ("n_k", C.c_uint64),
    ]


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.oracle_jaccard_from_samebits.restype = C.c_double
        _lib.oracle_jaccard_from_samebits.argtypes = [C.c_uint32, C.c_uint64, C.c_int, C.c_double, C.c_double, C.c_double]
        _lib.oracle_ani_pois.restype = C.c_double
        _lib.oracle_ani_pois.argtypes = [C.c_double, C.c_double]
        _lib.oracle_samebits.restype = C.c_uint32
        for f in ("oracle_square_to_condensed", "oracle_calc_row_idx", "oracle_calc_col_idx"):
            getattr(_lib, f).restype = C.c_uint64
        _lib.oracle_square_to_condensed.argtypes = [C.c_uint64] * 3
        _lib.oracle_calc_row_idx.argtypes = [C.c_uint64] * 2
        _lib.oracle_calc_col_idx.argtypes = [C.c_uint64] * 3
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Db:
    """A sketch database in the reference's in-RAM layout (flat u64, sample-major)."""

    def __init__(self, bins, n_samples, kmers, sketchsize64, sketch_size=None, sample_stride=None, kmer_stride=None,
                 names=None):
        self.bins = np.ascontiguousarray(bins, dtype=np.uint64)
        self.kmers = np.ascontiguousarray(kmers, dtype=np.uint64)
        self.n = int(n_samples)
        self.W = int(sketchsize64)
        self.sketch_size = int(sketch_size if sketch_size is not None else 64 * self.W)
        self.kmer_stride = int(kmer_stride if kmer_stride is not None else self.W * 14)
        self.sample_stride = int(sample_stride if sample_stride is not None else self.kmer_stride * len(self.kmers))
        self.names = names if names is not None else [f"s{i}" for i in range(self.n)]
        assert self.bins.size >= self.n * self.sample_stride
        self.c = _Db(_p(self.bins), self.n, self.sample_stride, self.kmer_stride, self.W, self.sketch_size,
                     _p(self.kmers), len(self.kmers))

    @classmethod
    def from_multisketch(cls, ms):
        return cls(ms.bins, ms.number_samples_loaded(), ms.kmer_lengths, ms.sketchsize64, ms.sketch_size,
                   ms.sample_stride, ms.kmer_stride, ms.names())

    @property
    def K(self):
        return len(self.kmers)

    def k_idx(self, k):  # get_k_idx, multisketch.rs:115-121
        hits = np.nonzero(self.kmers == k)[0]
        if len(hits) == 0:
            raise KeyError(f"K-mer size {k} not found in file")
        return int(hits[0])


def _c_or_none(c):
    return None if c is None else np.ascontiguousarray(c, dtype=np.float64)


def match_counts(ref: Db, query: Db, pairs) -> np.ndarray:
    pairs = np.ascontiguousarray(pairs, dtype=np.uint64).reshape(-1, 2)
    out = np.zeros((len(pairs), ref.K), dtype=np.uint32)
    lib().oracle_match_counts(C.byref(ref.c), C.byref(query.c), _p(pairs), C.c_uint64(len(pairs)), _p(out))
    return out


def core_acc_from_counts(counts, kmers, sketchsize64, sketch_size, c1=None, c2=None, cutoff=0.64):
    counts = np.ascontiguousarray(counts, dtype=np.uint32)
    kmers = np.ascontiguousarray(kmers, dtype=np.uint64)
    core, acc = C.c_float(), C.c_float()
    has_c = c1 is not None and c2 is not None
    lib().oracle_core_acc_from_counts(_p(counts), _p(kmers), C.c_uint64(len(kmers)), C.c_uint64(sketchsize64),
                                      C.c_uint64(sketch_size), C.c_int(has_c), C.c_double(c1 or 0.0),
                                      C.c_double(c2 or 0.0), C.c_double(cutoff), C.byref(core), C.byref(acc))
    return core.value, acc.value


def self_dists_all(db: Db, k_idx=-1, ani=False, completeness=None, cutoff=0.64, threads=1, first=0, count=None):
    n = db.n
    total = n * (n - 1) // 2
    count = total - first if count is None else count
    ncols = 1 if k_idx >= 0 else 2
    out = np.zeros(count * ncols, dtype=np.float32)
    cv = _c_or_none(completeness)
    rc = lib().oracle_self_dists_all(C.byref(db.c), C.c_int64(k_idx), C.c_int(ani), _p(cv), C.c_double(cutoff),
                                     C.c_uint64(first), C.c_uint64(count), _p(out), C.c_int(threads))
    if rc != 0:
        raise RuntimeError("Need at least two k-mer lengths to calculate core/accessory distances")
    return out


def cross_dists_all(ref: Db, query: Db, k_idx=-1, ani=False, ref_c=None, query_c=None, cutoff=0.64, threads=1,
                    first=0, count=None):
    total = ref.n * query.n
    count = total - first if count is None else count
    ncols = 1 if k_idx >= 0 else 2
    out = np.zeros(count * ncols, dtype=np.float32)
    rcv, qcv = _c_or_none(ref_c), _c_or_none(query_c)
    rc = lib().oracle_cross_dists_all(C.byref(ref.c), C.byref(query.c), C.c_int64(k_idx), C.c_int(ani), _p(rcv),
                                      _p(qcv), C.c_double(cutoff), C.c_uint64(first), C.c_uint64(count), _p(out),
                                      C.c_int(threads))
    if rc != 0:
        raise RuntimeError("Need at least two k-mer lengths to calculate core/accessory distances")
    return out


def self_dists_knn(db: Db, knn, k_idx=-1, ani=False, completeness=None, cutoff=0.64, threads=1, rows=None):
    r0, r1 = (0, db.n) if rows is None else rows
    idx = np.zeros((r1 - r0, knn), dtype=np.uint64)
    d0 = np.zeros((r1 - r0, knn), dtype=np.float32)
    d1 = np.zeros((r1 - r0, knn), dtype=np.float32) if k_idx < 0 else None
    cv = _c_or_none(completeness)
    rc = lib().oracle_self_dists_knn(C.byref(db.c), C.c_uint64(knn), C.c_int64(k_idx), C.c_int(ani), _p(cv),
                                     C.c_double(cutoff), C.c_uint64(r0), C.c_uint64(r1), _p(idx), _p(d0), _p(d1),
                                     C.c_int(threads))
    if rc != 0:
        raise RuntimeError(f"oracle_self_dists_knn failed ({rc})")
    return idx, d0, d1


def cross_dists_knn(ref: Db, query: Db, knn, k_idx=-1, ani=False, ref_c=None, query_c=None, cutoff=0.64, threads=1,
                    rows=None):
    knn = min(knn, ref.n)
    r0, r1 = (0, query.n) if rows is None else rows
    idx = np.zeros((r1 - r0, knn), dtype=np.uint64)
    d0 = np.zeros((r1 - r0, knn), dtype=np.float32)
    d1 = np.zeros((r1 - r0, knn), dtype=np.float32) if k_idx < 0 else None
    used = C.c_uint64()
    rcv, qcv = _c_or_none(ref_c), _c_or_none(query_c)
    rc = lib().oracle_cross_dists_knn(C.byref(ref.c), C.byref(query.c), C.c_uint64(knn), C.c_int64(k_idx),
                                      C.c_int(ani), _p(rcv), _p(qcv), C.c_double(cutoff), C.c_uint64(r0),
                                      C.c_uint64(r1), _p(idx), _p(d0), _p(d1), C.byref(used), C.c_int(threads))
    if rc != 0:
        raise RuntimeError(f"oracle_cross_dists_knn failed ({rc})")
    return idx, d0, d1


# ---------------------------------------------------------------- sketching (input generation only)
def read_fasta_records(path):
    opener = gzip.open if path.endswith(".gz") else open
    recs, cur = [], []
    with opener(path, "rb") as fh:
        for line in fh:
            if line.startswith(b">"):
                if cur:
                    recs.append(b"".join(cur))
                cur = []
            else:
                cur.append(line.strip())
    if cur:
        recs.append(b"".join(cur))
    return recs


def sketch_dna(seq_records, kmers, sketch_size, rc=True):
    """seq_records: list of bytes (one per FASTA record) -> (usigs u64[K*W*14], densified)."""
    W = (sketch_size + 63) // 64
    kmers = np.ascontiguousarray(sorted(kmers), dtype=np.uint64)
    seq = np.frombuffer(b"\n".join(seq_records), dtype=np.uint8)
    out = np.zeros(len(kmers) * W * 14, dtype=np.uint64)
    dens = lib().oracle_sketch_dna(_p(seq), C.c_uint64(len(seq)), _p(kmers), C.c_uint64(len(kmers)),
                                   C.c_uint64(sketch_size), C.c_int(rc), _p(out))
    return out, bool(dens)


def bin_minima(seq_records, k, nbins, rc=True) -> np.ndarray:
    """Densified bin minima (u64[nbins]) of one sample at one k; nbins need not be a multiple of 64 (the inverted
    index sketches with exactly --sketch-size bins, src/lib.rs:583)."""
    seq = np.frombuffer(b"\n".join(seq_records), dtype=np.uint8)
    out = np.zeros(nbins, dtype=np.uint64)
    lib().oracle_bin_minima(_p(seq), C.c_uint64(len(seq)), C.c_uint64(k), C.c_uint64(nbins), C.c_int(rc), _p(out))
    return out


def fill_usigs(bins16) -> np.ndarray:
    """bins16: uint16[..., nbins] (nbins % 64 == 0) -> u64[..., nbins/64*14] bit planes (fill_usigs)."""
    bins16 = np.ascontiguousarray(bins16, dtype=np.uint16)
    nbins = bins16.shape[-1]
    assert nbins % 64 == 0
    flat = bins16.reshape(-1, nbins)
    out = np.zeros((flat.shape[0], nbins // 64 * 14), dtype=np.uint64)
    for r in range(flat.shape[0]):
        lib().oracle_fill_usigs(_p(flat[r]), C.c_uint64(nbins), _p(out[r]))
    return out.reshape(bins16.shape[:-1] + (nbins // 64 * 14,))


def fmt_f32(x) -> str:
    """Rust `{}` for f32: shortest round-trip digits, positional, `1.0` -> `1`."""
    return np.format_float_positional(np.float32(x), unique=True, trim="-")


# ---------------------------------------------------------------- precluster (candidate-list kNN)
def any_shared_bins_csr(skq_bins, ski_to_skd=None):
    """skq_bins: uint16 [n_ski, stride] (untransposed bins, .skq layout, src/inverted.rs:92-100).  For each
    sample (in .skd order) the ascending-.ski-index list of samples sharing at least one (position, value) bin
    with it -- what Inverted::any_shared_bins returns (src/inverted.rs:259-268) -- mapped to .skd indices.
    Pure-Python/numpy, O(n^2 * stride / 8): test sizes only."""
    skq = np.ascontiguousarray(skq_bins, dtype=np.uint16)
    n = skq.shape[0]
    ski_to_skd = np.arange(n) if ski_to_skd is None else np.asarray(ski_to_skd)
    skd_to_ski = np.empty(n, dtype=np.int64)
    skd_to_ski[ski_to_skd] = np.arange(n)
    off, idx = [0], []
    for skd_i in range(n):
        row = skq[skd_to_ski[skd_i]]
        hits = np.nonzero((skq == row[None, :]).any(axis=1))[0]  # ascending ski index, includes itself
        idx.extend(int(ski_to_skd[h]) for h in hits)
        off.append(len(idx))
    return np.array(off, dtype=np.uint64), np.array(idx, dtype=np.uint64)


def self_dists_knn_precluster(db: Db, cand_off, cand_idx, knn, k_idx, ani=False, completeness=None, cutoff=0.64,
                              retain=0):
    cand_off = np.ascontiguousarray(cand_off, dtype=np.uint64)
    cand_idx = np.ascontiguousarray(cand_idx, dtype=np.uint64)
    idx = np.zeros((db.n, knn), dtype=np.uint64)
    d0 = np.zeros((db.n, knn), dtype=np.float32)
    cv = _c_or_none(completeness)
    rc = lib().oracle_self_dists_knn_precluster(C.byref(db.c), _p(cand_off), _p(cand_idx), C.c_uint64(knn),
                                                C.c_int64(k_idx), C.c_int(ani), _p(cv), C.c_double(cutoff),
                                                C.c_int(retain), _p(idx), _p(d0))
    if rc != 0:
        raise RuntimeError("Prefilter only available for single k-mer distances")
    return idx, d0

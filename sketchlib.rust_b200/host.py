"""Host-side mirrors of the reference's file-level types, bound to the C++ readers of the CLI (libsketchhost.so, no CUDA):

* `MultiSketch` -- `.skm` metadata + `.skd` loading (src/sketch/multisketch.rs:23-43, 90-210): `load_metadata`,
  `read_sketch_data`, `read_sketch_data_block` (subset order), `number_samples_loaded`, `sketch_name`, `get_k_idx`.
  The sketch data itself goes straight from the file into HBM (`skd_db_load_skd`): `read_sketch_data*` return the
  `SketchDb` the distance functions take.
* `read_completeness_file` (src/io.rs:240-324).
* `Inverted` -- `.ski` index (src/inverted.rs:46-57, 200-268) and `read_skq`; `precluster_from_files` runs the
  reference's `inverted precluster` flow on them.
"""
import ctypes as C
import os

import numpy as np

_LIB = None
_ERR = 1024


def load_library():
    global _LIB
    if _LIB is None:
        here = os.path.dirname(os.path.abspath(__file__))
        path = os.environ.get("SKH_LIB_OVERRIDE") or os.path.join(here, "lib", "libsketchhost.so")
        if not os.path.exists(path):
            from . import build
            path = build.build_hostlib()
        lib = C.CDLL(path)
        vp, u64, cp = C.c_void_p, C.c_uint64, C.c_char_p
        sigs = {
            "skh_skm_load": (C.c_int, [cp, C.POINTER(vp), cp, C.c_size_t]),
            "skh_skm_free": (None, [vp]),
            "skh_skm_n_loaded": (u64, [vp]), "skh_skm_n_total": (u64, [vp]), "skh_skm_n_kmers": (C.c_uint32, [vp]),
            "skh_skm_kmers": (None, [vp, vp]),
            "skh_skm_sketch_size": (u64, [vp]), "skh_skm_sketchsize64": (u64, [vp]),
            "skh_skm_sample_stride": (u64, [vp]), "skh_skm_kmer_stride": (u64, [vp]),
            "skh_skm_sketch_version": (cp, [vp]), "skh_skm_hash_type": (cp, [vp]),
            "skh_skm_sketch_name": (cp, [vp, u64]),
            "skh_skm_file_indices": (None, [vp, vp]),
            "skh_skm_select_subset": (C.c_int, [vp, C.POINTER(cp), u64, cp, C.c_size_t]),
            "skh_read_completeness_file": (C.c_int, [cp, vp, vp, cp, C.c_size_t, cp, C.c_size_t]),
            "skh_ski_load": (C.c_int, [cp, C.POINTER(vp), cp, C.c_size_t]),
            "skh_ski_free": (None, [vp]),
            "skh_ski_n_samples": (u64, [vp]), "skh_ski_sketch_size": (u64, [vp]), "skh_ski_kmer": (u64, [vp]),
            "skh_ski_rc": (C.c_int, [vp]),
            "skh_ski_sample_name": (cp, [vp, u64]),
            "skh_ski_any_shared_bins": (u64, [vp, vp, vp, u64]),
            "skh_ski_match_counts": (None, [vp, vp, vp]),
            "skh_ski_any_shared_bin_pairs": (u64, [vp]),
            "skh_write_dense": (C.c_int, [cp, C.c_int, C.POINTER(cp), u64, C.POINTER(cp), u64, vp, C.c_int, u64, u64, C.c_int,
                                          cp, C.c_size_t]),
            "skh_write_sparse": (C.c_int, [cp, C.c_int, C.POINTER(cp), C.POINTER(cp), u64, u64, u64, vp, vp, vp, C.c_int, cp,
                                           C.c_size_t]),
        }
        for name, (res, args) in sigs.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _LIB = lib
    return _LIB


class HostError(RuntimeError):
    pass


def _call(fn, *args):
    err = C.create_string_buffer(_ERR)
    if fn(*args, err, _ERR) != 0:
        raise HostError(err.value.decode(errors="replace"))


class MultiSketch:
    """`.skm` metadata of a sketch database; see the module docstring."""

    def __init__(self, file_prefix):
        self.lib = load_library()
        self.prefix = file_prefix[:-4] if file_prefix.endswith((".skm", ".skd")) else file_prefix  # utils.rs:9-15
        h = C.c_void_p()
        _call(self.lib.skh_skm_load, os.fsencode(self.prefix), C.byref(h))
        self.handle = h
        k = np.zeros(self.lib.skh_skm_n_kmers(h), dtype=np.uint64)
        self.lib.skh_skm_kmers(h, k.ctypes.data_as(C.c_void_p))
        self.kmer_lengths = [int(x) for x in k]
        self.sketch_size = int(self.lib.skh_skm_sketch_size(h))
        self.sketchsize64 = int(self.lib.skh_skm_sketchsize64(h))
        self.sample_stride = int(self.lib.skh_skm_sample_stride(h))
        self.kmer_stride = int(self.lib.skh_skm_kmer_stride(h))
        self.sketch_version = self.lib.skh_skm_sketch_version(h).decode()
        self.hash_type = self.lib.skh_skm_hash_type(h).decode()

    @classmethod
    def load_metadata(cls, file_prefix):  # multisketch.rs:90-103
        return cls(file_prefix)

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            self.lib.skh_skm_free(h)

    def number_samples_loaded(self):
        return int(self.lib.skh_skm_n_loaded(self.handle))

    def sketch_name(self, idx):
        return self.lib.skh_skm_sketch_name(self.handle, int(idx)).decode()

    def names(self):
        return [self.sketch_name(i) for i in range(self.number_samples_loaded())]

    def get_k_idx(self, k):
        return self.kmer_lengths.index(k) if k in self.kmer_lengths else None

    def file_indices(self):
        """File positions of the loaded samples: 0..n-1 for a full load (read_sketch_data reads by position,
        multisketch.rs:167-184), Sketch::index of the selected samples after select_subset."""
        out = np.zeros(self.number_samples_loaded(), dtype=np.uint64)
        self.lib.skh_skm_file_indices(self.handle, out.ctypes.data_as(C.c_void_p))
        return out

    def _to_device(self, ctx):
        from . import SketchDb
        return SketchDb.load_skd(ctx, self.prefix + ".skd", self.number_samples_loaded(), self.kmer_lengths,
                                 self.sketchsize64, self.sketch_size, self.sample_stride, self.kmer_stride, self.names(),
                                 sample_indices=self.file_indices() if self.has_subset else None)

    def read_sketch_data(self, ctx):
        """multisketch.rs:167-184: every sample of the file, into HBM."""
        return self._to_device(ctx)

    has_subset = False

    def select_subset(self, names):
        """The name lookup of read_sketch_data_block (multisketch.rs:187-200): afterwards the loaded samples are
        `names`, in that order.  Raises HostError("Could not find requested sample ...") like the reference panics."""
        arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
        _call(self.lib.skh_skm_select_subset, self.handle, arr, len(names))
        self.has_subset = True

    def read_sketch_data_block(self, ctx, names):
        """multisketch.rs:187-210: only `names`, in that order (the --subset list)."""
        self.select_subset(names)
        return self._to_device(ctx)


def read_subset_names(subset_file):
    """src/io.rs:227-236: every line verbatim (BufRead::lines: the trailing "\n" / "\r\n" is stripped, nothing else),
    so names may contain spaces and a blank line is a name that will not be found -- like the C++ reader."""
    with open(subset_file, newline="") as fh:
        lines = fh.read().split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    return [ln[:-1] if ln.endswith("\r") else ln for ln in lines]


def read_completeness_file(path, sketches: MultiSketch):
    """src/io.rs:240-324 -> (values per loaded sample, list of warnings)."""
    out = np.zeros(sketches.number_samples_loaded(), dtype=np.float64)
    warn = C.create_string_buffer(1 << 16)
    _call(sketches.lib.skh_read_completeness_file, os.fsencode(path), sketches.handle, out.ctypes.data_as(C.c_void_p), warn,
          1 << 16)
    text = warn.value.decode(errors="replace")
    return out, (text.split("\n") if text else [])


class Inverted:
    """`.ski` inverted index (candidate generator of the precluster mode)."""

    def __init__(self, file_prefix):
        self.lib = load_library()
        self.prefix = file_prefix[:-4] if file_prefix.endswith(".ski") else file_prefix
        h = C.c_void_p()
        _call(self.lib.skh_ski_load, os.fsencode(self.prefix), C.byref(h))
        self.handle = h
        self.n = int(self.lib.skh_ski_n_samples(h))
        self.stride = int(self.lib.skh_ski_sketch_size(h))
        self.kmer = int(self.lib.skh_ski_kmer(h))
        self.rc = bool(self.lib.skh_ski_rc(h))
        self.sample_names = [self.lib.skh_ski_sample_name(h, i).decode() for i in range(self.n)]

    load = classmethod(lambda cls, file_prefix: cls(file_prefix))

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            self.lib.skh_ski_free(h)

    def sketch_size(self):
        return self.stride

    def _sigs(self, sigs):
        s = np.ascontiguousarray(sigs, dtype=np.uint16)
        if s.size != self.stride:
            raise ValueError(f"query has {s.size} bins, the index {self.stride}")
        return s

    def any_shared_bins(self, sigs):
        s = self._sigs(sigs)
        out = np.zeros(self.n, dtype=np.uint32)
        cnt = self.lib.skh_ski_any_shared_bins(self.handle, s.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), self.n)
        return out[:cnt]

    def query_against_inverted_index(self, sigs):
        s = self._sigs(sigs)
        out = np.zeros(self.n, dtype=np.uint32)
        self.lib.skh_ski_match_counts(self.handle, s.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        return out

    def any_shared_bin_pairs(self):
        return int(self.lib.skh_ski_any_shared_bin_pairs(self.handle))

    def read_skq(self):
        """The `.skq` written next to the index by `inverted build --write-skq`: n x sketch_size u16, index order."""
        from .precluster import read_skq
        return read_skq(self.prefix + ".skq", self.n, self.stride)


def precluster_from_files(ctx, ski_prefix, skd_prefix, knn, ani=False, retain_unmatched=None, ref_completeness_file=None,
                          completeness_cutoff=0.64):
    """The reference's `inverted precluster IDX.ski --skd DB` arm (src/lib.rs:682-792) on the GPU: candidates from the
    index over the .skq bins, distances and per-row top-k through skd_self_dists_knn_candidates."""
    import sketchlib.rust_b200 as sk
    inv = Inverted(ski_prefix)
    skq = inv.read_skq()
    refs = MultiSketch(skd_prefix)
    db = refs.read_sketch_data(ctx)
    n = refs.number_samples_loaded()
    knn = min(int(knn), n - 1)
    if inv.kmer not in refs.kmer_lengths:
        raise HostError(f"K-mer size {inv.kmer} used for .ski not found in .skd: K-mer size {inv.kmer} not found in file")
    dt = sk.set_k(db, inv.kmer, ani=ani)
    comp = None
    if ref_completeness_file:
        comp, _ = read_completeness_file(ref_completeness_file, refs)
    return sk.self_dists_knn_precluster(db, skq, inv.sample_names, knn, dt, retain_unmatched, comp, completeness_cutoff)


def _names(names):
    return (C.c_char_p * len(names))(*[n.encode() for n in names])


def write_dense(path, ref_names, query_names, distances, ncols, row_range, threads=1, append=False):
    """Display of DistanceMatrix (distance_matrix.rs:175-209) written to `path` by `threads` host threads."""
    lib = load_library()
    d = np.ascontiguousarray(distances, dtype=np.float32)
    q = None if query_names is None else _names(query_names)
    _call(lib.skh_write_dense, os.fsencode(path), int(append), _names(ref_names), len(ref_names), q,
          0 if query_names is None else len(query_names), d.ctypes.data_as(C.c_void_p), int(ncols), int(row_range[0]),
          int(row_range[1]), int(threads))


def write_sparse(path, row_names, col_names, row_begin, idx, d0, d1, threads=1, append=False):
    """Display of SparseDistanceMatrix (distance_matrix.rs:360-401) written to `path`."""
    lib = load_library()
    idx = np.ascontiguousarray(idx, dtype=np.uint64)
    d0 = np.ascontiguousarray(d0, dtype=np.float32)
    d1 = None if d1 is None else np.ascontiguousarray(d1, dtype=np.float32)
    n_rows, knn = idx.shape
    _call(lib.skh_write_sparse, os.fsencode(path), int(append), _names(row_names), _names(col_names), int(row_begin), n_rows,
          knn, idx.ctypes.data_as(C.c_void_p), d0.ctypes.data_as(C.c_void_p),
          None if d1 is None else d1.ctypes.data_as(C.c_void_p), int(threads))

"""sketchlib.rust_b200 -- B200 (sm_100a) implementation of sketchlib.rust's pairwise sketch-distance path.

This module is the Python host layer above the C ABI (include/sketchdist.h, lib/libsketchdist.so).  It
mirrors the reference's `sketchlib::distances` API -- `set_k`, `self_dists_all`, `self_dists_knn`,
`cross_dists_all`, `cross_dists_knn`, `DistanceMatrix`, `SparseDistanceMatrix`
(reference src/distances/mod.rs, src/distances/distance_matrix.rs) -- with the same argument meaning,
output layouts, text format and error behaviour.  All compute happens in the CUDA library; there is no
CPU fallback: importing works anywhere (so host logic can be tested), but creating a `Context` raises
unless libsketchdist.so is built and a B200 is visible.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

__all__ = [
    "Context", "SketchDb", "DistType", "DistanceMatrix", "SparseDistanceMatrix", "SketchDistError", "set_k",
    "self_dists_all", "self_dists_knn", "cross_dists_all", "cross_dists_knn", "pair_dists", "match_counts",
    "self_knn_partial", "cross_knn_partial", "knn_merge", "self_dists_knn_precluster",
    "row_partition", "triangle_row_partition", "slice_rows", "PeerGroup", "load_library", "library_path",
]

SKD_HOST, SKD_DEVICE = 0, 1
CORE_ACC, JACCARD, ANI = 0, 1, 2
_STATUS = {0: "ok", 1: "invalid argument", 2: "CUDA error", 3: "need two k-mer lengths", 4: "k-mer not found",
           5: "empty database", 6: "incompatible databases", 7: "unsupported", 8: "I/O error", 9: "out of memory"}

_lib = None


class SketchDistError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"{_STATUS.get(status, status)}: {message}")
        self.status = status


class _DistType(C.Structure):
    _fields_ = [("kind", C.c_int32), ("k_idx", C.c_int32)]


def library_path():
    return _build.LIB


def load_library(build=True):
    """Load lib/libsketchdist.so, building it with nvcc first if it is missing or stale."""
    global _lib
    if _lib is not None:
        return _lib
    if build:
        _build.build()
    if not os.path.exists(_build.LIB):
        raise SketchDistError(2, f"{_build.LIB} is missing: the CUDA extension has not been built "
                                 "(python sketchlib.rust_b200/build.py); there is no CPU fallback")
    L = C.CDLL(os.environ.get("SKD_LIB_OVERRIDE", _build.LIB))  # override: ablation builds (tools/)
    vp, u64, u32, i64, dbl, f32p = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int64, C.c_double, C.c_void_p
    sig = {
        "skd_device_count": (C.c_int, []),
        "skd_ctx_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
        "skd_ctx_destroy": (None, [vp]),
        "skd_last_error": (C.c_char_p, [vp]),
        "skd_ctx_device": (C.c_int, [vp]),
        "skd_ctx_launch_count": (u64, [vp]),
        "skd_ctx_set_early_exit": (C.c_int, [vp, C.c_int]),
        "skd_ctx_work_counters": (C.c_int, [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_int]),
        "skd_ctx_last_kernel_ms": (C.c_float, [vp]),
        "skd_ctx_last_tile_kernel_ms": (C.c_float, [vp]),
        "skd_ctx_synchronize": (C.c_int, [vp]),
        "skd_ctx_stream": (vp, [vp]),
        "skd_host_alloc": (C.c_int, [vp, C.c_size_t, C.POINTER(vp)]),
        "skd_host_free": (None, [vp, vp]),
        "skd_device_alloc": (C.c_int, [vp, C.c_size_t, C.POINTER(vp)]),
        "skd_device_free": (None, [vp, vp]),
        "skd_memcpy": (C.c_int, [vp, vp, C.c_int, vp, C.c_int, C.c_size_t]),
        "skd_db_upload": (C.c_int, [vp, vp, C.c_int, u64, u64, u64, u64, u64, vp, u32, C.POINTER(vp)]),
        "skd_db_upload_async": (C.c_int, [vp, vp, u64, u64, u64, u64, u64, vp, u32, C.POINTER(vp)]),
        "skd_db_load_skd": (C.c_int, [vp, C.c_char_p, vp, u64, u64, u64, u64, u64, vp, u32, C.POINTER(vp)]),
        "skd_db_free": (None, [vp]),
        "skd_db_adopt": (C.c_int, [vp, vp]),
        "skd_db_n_samples": (u64, [vp]),
        "skd_db_n_kmers": (u32, [vp]),
        "skd_db_sketch_size": (u64, [vp]),
        "skd_db_hbm_bytes": (u64, [vp]),
        "skd_db_compatible": (C.c_int, [vp, vp]),
        "skd_set_k": (C.c_int, [vp, vp, i64, C.c_int, C.POINTER(_DistType)]),
        "skd_self_dists_all": (C.c_int, [vp, vp, _DistType, vp, dbl, u64, u64, f32p, C.c_int]),
        "skd_self_pairs_in_rows": (u64, [u64, u64, u64]),
        "skd_cross_dists_all": (C.c_int, [vp, vp, vp, _DistType, vp, vp, dbl, u64, u64, f32p, C.c_int]),
        "skd_self_dists_knn": (C.c_int, [vp, vp, u64, _DistType, vp, dbl, u64, u64, vp, vp, vp, C.c_int]),
        "skd_cross_dists_knn": (C.c_int, [vp, vp, vp, u64, _DistType, vp, vp, dbl, u64, u64, vp, vp, vp, C.c_int,
                                          C.POINTER(u64)]),
        "skd_self_knn_partial": (C.c_int, [vp, vp, u64, _DistType, vp, dbl, u64, u64, vp, vp, C.c_int]),
        "skd_cross_knn_partial": (C.c_int, [vp, vp, vp, u64, _DistType, vp, vp, dbl, u64, vp, vp, C.c_int]),
        "skd_knn_merge": (C.c_int, [vp, vp, vp, u32, u64, u64, _DistType, C.c_int, vp, vp, vp, C.c_int]),
        "skd_self_dists_knn_candidates": (C.c_int, [vp, vp, u64, _DistType, vp, dbl, vp, vp, C.c_int, vp, vp]),
        "skd_pair_dists": (C.c_int, [vp, vp, vp, _DistType, vp, vp, dbl, vp, u64, vp]),
        "skd_match_counts": (C.c_int, [vp, vp, vp, vp, u64, vp]),
        "skd_match_counts_dense": (C.c_int, [vp, vp, vp, u64, u64, vp]),
        "skd_comm_create": (C.c_int, [vp, C.c_int, C.c_int, C.c_size_t, C.POINTER(vp)]),
        "skd_comm_destroy": (None, [vp]),
        "skd_comm_export": (C.c_int, [vp, vp]),
        "skd_comm_connect": (C.c_int, [vp, vp]),
        "skd_comm_rank": (C.c_int, [vp]),
        "skd_comm_world": (C.c_int, [vp]),
        "skd_comm_last_op_ms": (C.c_float, [vp]),
        "skd_comm_last_exchange_ms": (C.c_float, [vp]),
        "skd_comm_last_row_block": (None, [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
        "skd_comm_window_bytes_knn": (C.c_size_t, [u64, u64, C.c_int]),
        "skd_comm_window_bytes_db": (C.c_size_t, [u64, u64, u32, C.c_int]),
        "skd_triangle_rows": (None, [u64, C.c_int, C.c_int, C.POINTER(u64), C.POINTER(u64)]),
        "skd_slice_rows": (None, [u64, C.c_int, C.c_int, u64, C.POINTER(u64), C.POINTER(u64)]),
        "skd_triangle_rows_weighted": (None, [u64, C.c_int, C.c_void_p, C.c_void_p]),
        "skd_comm_db_upload": (C.c_int, [vp, vp, C.c_int, u64, u64, u64, u64, u64, vp, u32, C.POINTER(vp)]),
        "skd_comm_db_load_skd": (C.c_int, [vp, C.c_char_p, vp, u64, u64, u64, u64, u64, vp, u32, C.POINTER(vp)]),
        "skd_comm_self_dists_knn": (C.c_int, [vp, vp, u64, _DistType, vp, dbl, C.c_int, vp, vp, vp, C.c_int]),
        "skd_comm_cross_dists_knn": (C.c_int, [vp, vp, vp, u64, _DistType, vp, vp, dbl, u64, u64, C.c_int, vp, vp, vp,
                                               C.c_int, C.POINTER(u64)]),
        "skd_comm_knn_merge": (C.c_int, [vp, vp, vp, u32, C.c_int, u64, u64, _DistType, C.c_int, vp, vp, vp, C.c_int]),
        "skd_comm_barrier": (C.c_int, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    L._signatures = sig
    _lib = L
    return L


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(int(a))


def _f64(v):
    return None if v is None else np.ascontiguousarray(v, dtype=np.float64)


class PinnedArray:
    """numpy view of page-locked host memory owned by the library (skd_host_alloc)."""

    def __init__(self, ctx, shape, dtype):
        self.ctx = ctx
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) if len(shape) else 1
        p = C.c_void_p()
        ctx._check(ctx.lib.skd_host_alloc(ctx.handle, max(1, n * dtype.itemsize), C.byref(p)))
        self.ptr = p.value
        buf = (C.c_char * max(1, n * dtype.itemsize)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            self.ctx.lib.skd_host_free(self.ctx.handle, C.c_void_p(self.ptr))
            self.ptr = None


class Context:
    """One per GPU per process (skd_ctx)."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = C.c_void_p()
        st = self.lib.skd_ctx_create(int(device), C.byref(h))
        if st != 0:
            raise SketchDistError(st, self.lib.skd_last_error(None).decode())
        self.handle = h
        self.device = int(device)

    def _check(self, st):
        if st != 0:
            raise SketchDistError(st, self.lib.skd_last_error(self.handle).decode())

    def close(self):
        if getattr(self, "handle", None):
            self.lib.skd_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self):
        return int(self.lib.skd_ctx_launch_count(self.handle))

    def set_early_exit(self, enabled):
        """skd_ctx_set_early_exit: the dense core/accessory kernel's early exit (on by default; same results)."""
        self._check(self.lib.skd_ctx_set_early_exit(self.handle, int(bool(enabled))))

    def work_counters(self, reset=False):
        """(steps_computed, steps_total, pairs_deferred) of the dense core/accessory kernel (skd_ctx_work_counters)."""
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(self.lib.skd_ctx_work_counters(self.handle, C.byref(a), C.byref(b), C.byref(c), int(bool(reset))))
        return int(a.value), int(b.value), int(c.value)

    @property
    def last_kernel_ms(self):
        return float(self.lib.skd_ctx_last_kernel_ms(self.handle))

    @property
    def last_tile_kernel_ms(self):
        return float(self.lib.skd_ctx_last_tile_kernel_ms(self.handle))

    @property
    def stream(self):
        return self.lib.skd_ctx_stream(self.handle)

    def synchronize(self):
        self._check(self.lib.skd_ctx_synchronize(self.handle))

    def pinned(self, shape, dtype):
        return PinnedArray(self, tuple(np.atleast_1d(shape)), dtype)

    def device_alloc(self, nbytes):
        p = C.c_void_p()
        self._check(self.lib.skd_device_alloc(self.handle, int(nbytes), C.byref(p)))
        return p.value

    def device_free(self, ptr):
        self.lib.skd_device_free(self.handle, C.c_void_p(ptr))

    def memcpy(self, dst, dst_loc, src, src_loc, nbytes):
        self._check(self.lib.skd_memcpy(self.handle, _ptr(dst), dst_loc, _ptr(src), src_loc, int(nbytes)))


class SketchDb:
    """A sketch database resident in HBM: the hot path's view of the reference's MultiSketch
    (src/sketch/multisketch.rs:23-43; sketch_size, sketchsize64, kmer_lengths, names, bins)."""

    def __init__(self, ctx, handle, names, kmer_lengths, sketchsize64, sketch_size):
        self.ctx, self.handle = ctx, handle
        self._names = None if names is None else list(names)
        self.kmer_lengths = [int(k) for k in kmer_lengths]
        self.sketchsize64, self.sketch_size = int(sketchsize64), int(sketch_size)

    @classmethod
    def upload(cls, ctx, bins, n_samples, kmer_lengths, sketchsize64, sketch_size=None, sample_stride=None,
               kmer_stride=None, names=None, bins_loc=SKD_HOST, async_upload=False):
        """bins: numpy u64 array in the reference's layout, or an int device pointer with bins_loc=SKD_DEVICE.
        async_upload=True (pinned host arrays from Context.pinned): return once the copies are queued; the array
        must stay unchanged until the first compute call on the database has returned (skd_db_upload_async)."""
        kmers = np.ascontiguousarray(kmer_lengths, dtype=np.uint64)
        W = int(sketchsize64)
        kmer_stride = int(kmer_stride if kmer_stride is not None else W * 14)
        sample_stride = int(sample_stride if sample_stride is not None else kmer_stride * len(kmers))
        sketch_size = int(sketch_size if sketch_size is not None else 64 * W)
        if isinstance(bins, np.ndarray):
            bins = np.ascontiguousarray(bins, dtype=np.uint64)
            if bins.size < int(n_samples) * sample_stride:
                raise ValueError("bins array shorter than n_samples * sample_stride")
        h = C.c_void_p()
        if async_upload and bins_loc == SKD_HOST:
            ctx._check(ctx.lib.skd_db_upload_async(ctx.handle, _ptr(bins), int(n_samples), sample_stride, kmer_stride,
                                                   W, sketch_size, _ptr(kmers), len(kmers), C.byref(h)))
        else:
            ctx._check(ctx.lib.skd_db_upload(ctx.handle, _ptr(bins), bins_loc, int(n_samples), sample_stride,
                                             kmer_stride, W, sketch_size, _ptr(kmers), len(kmers), C.byref(h)))
        return cls(ctx, h, names, kmers.tolist(), W, sketch_size)

    @classmethod
    def load_skd(cls, ctx, skd_path, n_samples, kmer_lengths, sketchsize64, sketch_size=None, sample_stride=None,
                 kmer_stride=None, names=None, sample_indices=None):
        kmers = np.ascontiguousarray(kmer_lengths, dtype=np.uint64)
        W = int(sketchsize64)
        kmer_stride = int(kmer_stride if kmer_stride is not None else W * 14)
        sample_stride = int(sample_stride if sample_stride is not None else kmer_stride * len(kmers))
        sketch_size = int(sketch_size if sketch_size is not None else 64 * W)
        idx = None if sample_indices is None else np.ascontiguousarray(sample_indices, dtype=np.uint64)
        h = C.c_void_p()
        ctx._check(ctx.lib.skd_db_load_skd(ctx.handle, os.fsencode(skd_path), _ptr(idx), int(n_samples), sample_stride,
                                           kmer_stride, W, sketch_size, _ptr(kmers), len(kmers), C.byref(h)))
        return cls(ctx, h, names, kmers.tolist(), W, sketch_size)

    @property
    def names(self):
        """Sample names; default names s0, s1, ... are made on first use (10 000 f-strings cost more than a launch)."""
        if self._names is None:
            self._names = [f"s{i}" for i in range(self.number_samples_loaded())]
        return self._names

    def number_samples_loaded(self):
        return int(self.ctx.lib.skd_db_n_samples(self.handle))

    def kmer_lengths_(self):
        return self.kmer_lengths

    def sketch_name(self, idx):
        return self.names[idx]

    def get_k_idx(self, k):
        return self.kmer_lengths.index(k) if k in self.kmer_lengths else None

    def is_compatible_with(self, other):
        return bool(self.ctx.lib.skd_db_compatible(self.handle, other.handle))

    @property
    def hbm_bytes(self):
        return int(self.ctx.lib.skd_db_hbm_bytes(self.handle))

    def free(self):
        if self.handle:
            self.ctx.lib.skd_db_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            if self.ctx.handle:
                self.free()
        except Exception:
            pass


class DistType:
    """DistType::{Jaccard(k_idx, k, ani), CoreAcc}: src/distances/distance_matrix.rs:55-76."""

    def __init__(self, kind, k_idx=0, k=None):
        self.kind, self.k_idx, self.k = kind, k_idx, k

    @property
    def c(self):
        return _DistType(self.kind, self.k_idx)

    @property
    def n_dist_cols(self):
        return 2 if self.kind == CORE_ACC else 1

    @property
    def ani(self):
        return self.kind == ANI

    def __str__(self):  # Display, distance_matrix.rs:62-76
        if self.kind == CORE_ACC:
            return "Distances: core/accessory regression"
        return f"Distances: ANI at k={self.k}" if self.kind == ANI else f"Distances: Jaccard distances at k={self.k}"


def set_k(sketches, kmer=None, ani=False):
    """set_k: src/distances/mod.rs:25-37.  Raises SketchDistError("K-mer size {k} not found in file")."""
    out = _DistType()
    ctx = sketches.ctx
    ctx._check(ctx.lib.skd_set_k(ctx.handle, sketches.handle, -1 if kmer is None else int(kmer), int(bool(ani)),
                                 C.byref(out)))
    return DistType(out.kind, out.k_idx, kmer)


def fmt_f32(x):
    """Rust `{}` for f32: shortest round-trip digits, positional notation, 1.0 -> "1"."""
    return np.format_float_positional(np.float32(x), unique=True, trim="-")


class DistanceMatrix:
    """Dense long-form matrix: src/distances/distance_matrix.rs:120-209.  `distances` is the f32 vector of the
    reference (n_distances x n_dist_cols, core/accessory interleaved); for a row-sharded call it holds only
    this shard's slice, `row_range` says which rows."""

    def __init__(self, ref_names, query_names, dist_type, distances, row_range, pinned=None):
        self.ref_names, self.query_names = ref_names, query_names
        self.dist_type, self.distances, self.row_range = dist_type, distances, row_range
        self.n_distances = distances.size // dist_type.n_dist_cols
        self._pinned = pinned

    def __str__(self):  # Display, distance_matrix.rs:175-209
        nc, d, out = self.dist_type.n_dist_cols, 0, []
        r0, r1 = self.row_range
        for i in range(r0, r1):
            cols = self.query_names if self.query_names is not None else self.ref_names[i + 1:]
            for cn in cols:
                vals = "\t".join(fmt_f32(self.distances[d * nc + c]) for c in range(nc))
                out.append(f"{self.ref_names[i]}\t{cn}\t{vals}\n")
                d += 1
        return "".join(out)

    def write(self, path, threads=1, append=False):
        """The same text as str(self), formatted by `threads` host threads in C++ (large matrices)."""
        from . import host
        host.write_dense(path, self.ref_names, self.query_names, self.distances, self.dist_type.n_dist_cols, self.row_range,
                         threads, append)


class SparseDistanceMatrix:
    """kNN rows: src/distances/distance_matrix.rs:285-401."""

    def __init__(self, ref_names, query_names, dist_type, knn, idx, d0, d1, row_range):
        self.ref_names, self.query_names, self.dist_type = ref_names, query_names, dist_type
        self.knn, self.idx, self.d0, self.d1, self.row_range = knn, idx, d0, d1, row_range
        self.n_distances = idx.size

    def __str__(self):  # Display, distance_matrix.rs:360-401
        rows = self.query_names if self.query_names is not None else self.ref_names
        out = []
        for r in range(self.idx.shape[0]):
            rn = rows[self.row_range[0] + r]
            for t in range(self.knn):
                cn = self.ref_names[int(self.idx[r, t])]
                if self.dist_type.kind == CORE_ACC:
                    out.append(f"{rn}\t{cn}\t{fmt_f32(self.d0[r, t])}\t{fmt_f32(self.d1[r, t])}\n")
                elif self.d0[r, t] < 1.0 or cn != rn:  # padding entries skipped, :376-381
                    out.append(f"{rn}\t{cn}\t{fmt_f32(self.d0[r, t])}\n")
        return "".join(out)

    def write(self, path, threads=1, append=False):
        """The same text as str(self), formatted in C++."""
        from . import host
        rows = self.query_names if self.query_names is not None else self.ref_names
        host.write_sparse(path, rows, self.ref_names, self.row_range[0], self.idx, self.d0,
                          self.d1 if self.dist_type.kind == CORE_ACC else None, threads, append)


# ---- multi-GPU partitioning (one process per GPU; no data-path collective) ---------------------------
def row_partition(n_rows, rank, world):
    """Even contiguous split (kNN rows / cross refs / cross queries)."""
    base, rem = divmod(n_rows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def triangle_row_partition(n, rank, world, align=64):
    """Row blocks of the upper triangle with about equal pair counts: row i holds n-1-i pairs, so the
    boundary after a fraction f of the pairs is n * (1 - sqrt(1 - f)); rounded to the 64-sample tile.
    Same rule as the library's skd_triangle_rows (which the peer-group calls use)."""
    def bound(r):
        if r <= 0:
            return 0
        if r >= world:
            return n
        x = n * (1.0 - (1.0 - r / world) ** 0.5)
        return min(n, int(x / align + 0.5) * align)   # llround: halves away from zero, like the C library
    return bound(rank), bound(rank + 1)


def triangle_rows_weighted(n, shares):
    """skd_triangle_rows_weighted: row boundaries [world + 1] of the triangle split for the given pair shares."""
    sh = np.ascontiguousarray(shares, dtype=np.float64)
    out = np.zeros(len(sh) + 1, dtype=np.uint64)
    load_library().skd_triangle_rows_weighted(int(n), len(sh), sh.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return [int(x) for x in out]


def slice_rows(n, rank, world, align=1):
    """skd_slice_rows: contiguous slices of ceil(n / world) items rounded up to `align` (sample slices of the
    distributed loader and reference shards: align 64; owned result rows: align 1)."""
    per = -(-(-(-n // world)) // align) * align
    a = min(n, per * rank)
    return a, min(n, a + per)


class PeerGroup:
    """skd_comm: this rank's membership in a group of `world` GPUs of one box working on one problem (one process
    per GPU, or several threads of one process).  Create on every rank, exchange the `export()` blobs with whatever
    the launcher offers (torch.distributed.all_gather_object, a file, a pipe), `connect()` with all of them in rank
    order, then call the collective methods on every rank in the same order."""

    def __init__(self, ctx, rank, world, window_bytes):
        self.ctx, self.rank, self.world = ctx, int(rank), int(world)
        h = C.c_void_p()
        ctx._check(ctx.lib.skd_comm_create(ctx.handle, self.rank, self.world, int(window_bytes), C.byref(h)))
        self.handle = h

    @staticmethod
    def window_bytes_knn(n_rows, knn, world):
        return int(load_library().skd_comm_window_bytes_knn(int(n_rows), int(knn), int(world)))

    @staticmethod
    def window_bytes_db(n_samples, sketchsize64, n_kmers, world):
        return int(load_library().skd_comm_window_bytes_db(int(n_samples), int(sketchsize64), int(n_kmers), int(world)))

    def export(self):
        buf = C.create_string_buffer(128)
        self.ctx._check(self.ctx.lib.skd_comm_export(self.handle, buf))
        return buf.raw

    def connect(self, handles):
        blob = b"".join(handles) if not isinstance(handles, (bytes, bytearray)) else bytes(handles)
        if len(blob) != 128 * self.world:
            raise ValueError(f"need {self.world} handles of 128 bytes")
        self.ctx._check(self.ctx.lib.skd_comm_connect(self.handle, blob))

    def close(self):
        if getattr(self, "handle", None) and self.ctx.handle:
            self.ctx.lib.skd_comm_destroy(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def last_op_ms(self):
        return float(self.ctx.lib.skd_comm_last_op_ms(self.handle))

    @property
    def last_exchange_ms(self):
        return float(self.ctx.lib.skd_comm_last_exchange_ms(self.handle))

    @property
    def last_row_block(self):
        """Triangle rows this rank evaluated in its last self_dists_knn call (the split adapts to measured times)."""
        a, b = C.c_uint64(), C.c_uint64()
        self.ctx.lib.skd_comm_last_row_block(self.handle, C.byref(a), C.byref(b))
        return int(a.value), int(b.value)

    def barrier(self):
        self.ctx._check(self.ctx.lib.skd_comm_barrier(self.handle))

    def my_sample_slice(self, n_total):
        return slice_rows(n_total, self.rank, self.world, 64)

    def upload(self, bins_slice, n_total, kmer_lengths, sketchsize64, sketch_size=None, names=None, bins_loc=SKD_HOST):
        """Distributed loader (skd_comm_db_upload): `bins_slice` = the reference-layout planes of this rank's sample
        slice `my_sample_slice(n_total)` (numpy u64 array or device pointer); returns the WHOLE database in HBM."""
        ctx = self.ctx
        kmers = np.ascontiguousarray(kmer_lengths, dtype=np.uint64)
        W = int(sketchsize64)
        kmer_stride, sample_stride = W * 14, W * 14 * len(kmers)
        sketch_size = int(sketch_size if sketch_size is not None else 64 * W)
        h = C.c_void_p()
        ctx._check(ctx.lib.skd_comm_db_upload(self.handle, _ptr(bins_slice), bins_loc, int(n_total), sample_stride,
                                              kmer_stride, W, sketch_size, _ptr(kmers), len(kmers), C.byref(h)))
        return SketchDb(ctx, h, names, kmers.tolist(), W, sketch_size)

    def load_skd(self, skd_path, n_total, kmer_lengths, sketchsize64, sketch_size=None, sample_stride=None,
                 kmer_stride=None, names=None, sample_indices=None):
        ctx = self.ctx
        kmers = np.ascontiguousarray(kmer_lengths, dtype=np.uint64)
        W = int(sketchsize64)
        kmer_stride = int(kmer_stride if kmer_stride is not None else W * 14)
        sample_stride = int(sample_stride if sample_stride is not None else kmer_stride * len(kmers))
        sketch_size = int(sketch_size if sketch_size is not None else 64 * W)
        idx = None if sample_indices is None else np.ascontiguousarray(sample_indices, dtype=np.uint64)
        h = C.c_void_p()
        ctx._check(ctx.lib.skd_comm_db_load_skd(self.handle, os.fsencode(skd_path), _ptr(idx), int(n_total),
                                                sample_stride, kmer_stride, W, sketch_size, _ptr(kmers), len(kmers),
                                                C.byref(h)))
        return SketchDb(ctx, h, names, kmers.tolist(), W, sketch_size)

    def _outputs(self, n_rows, knn, dist_type, root, out):
        mine = root < 0 or root == self.rank
        if out is not None:
            return out
        if not mine:
            return None, None, None
        idx = np.zeros((n_rows, knn), dtype=np.uint64)
        d0 = np.zeros((n_rows, knn), dtype=np.float32)
        d1 = np.zeros((n_rows, knn), dtype=np.float32) if dist_type.kind == CORE_ACC else None
        return idx, d0, d1

    def self_dists_knn(self, sketches, knn=50, dist_type=None, completeness_vec=None, completeness_cutoff=0.64, root=0,
                       out=None, out_loc=SKD_HOST):
        """self_dists_knn (src/distances/mod.rs:133-224) over the group; the SparseDistanceMatrix on `root` (every
        rank if root < 0), None elsewhere.  `out` = (idx, d0, d1) preallocated arrays / device pointers."""
        ctx = self.ctx
        n = sketches.number_samples_loaded()
        dist_type = dist_type or DistType(CORE_ACC)
        idx, d0, d1 = self._outputs(n, knn, dist_type, root, out)
        ctx._check(ctx.lib.skd_comm_self_dists_knn(self.handle, sketches.handle, int(knn), dist_type.c,
                                                   _ptr(_f64(completeness_vec)), float(completeness_cutoff), int(root),
                                                   _ptr(idx), _ptr(d0), _ptr(d1), out_loc))
        if idx is None or out is not None:
            return None
        return SparseDistanceMatrix(sketches.names, None, dist_type, knn, idx, d0, d1, (0, n))

    def knn_merge(self, keys_parts, aux_parts, dist_type, ref_names, query_names=None, root=0):
        """skd_comm_knn_merge: this rank's partial lists (one or several [n_rows, knn] sets from self_knn_partial /
        cross_knn_partial) merged across the group; the SparseDistanceMatrix on `root`."""
        ctx = self.ctx
        keys = np.ascontiguousarray(np.stack(keys_parts), dtype=np.uint64)
        n_parts, n_rows, knn = keys.shape
        aux = np.ascontiguousarray(np.stack(aux_parts), dtype=np.float32) if dist_type.kind == CORE_ACC else None
        idx, d0, d1 = self._outputs(n_rows, knn, dist_type, root, None)
        ctx._check(ctx.lib.skd_comm_knn_merge(self.handle, _ptr(keys), _ptr(aux), n_parts, SKD_HOST, n_rows, knn,
                                              dist_type.c, int(root), _ptr(idx), _ptr(d0), _ptr(d1), SKD_HOST))
        if idx is None:
            return None
        return SparseDistanceMatrix(ref_names, query_names, dist_type, knn, idx, d0, d1, (0, n_rows))

    def cross_dists_knn(self, ref_shard, query_sketches, knn, dist_type, ref_index_offset, n_ref_total, ref_names=None,
                        ref_completeness_vec=None, query_completeness_vec=None, completeness_cutoff=0.64, root=0,
                        out=None, out_loc=SKD_HOST):
        """cross_dists_knn (src/distances/mod.rs:306-395) with the references sharded over the group."""
        ctx = self.ctx
        nq = query_sketches.number_samples_loaded()
        knn_c = max(1, min(int(knn), int(n_ref_total)))
        idx, d0, d1 = self._outputs(nq, knn_c, dist_type, root, out)
        used = C.c_uint64()
        ctx._check(ctx.lib.skd_comm_cross_dists_knn(self.handle, ref_shard.handle, query_sketches.handle, int(knn),
                                                    dist_type.c, _ptr(_f64(ref_completeness_vec)),
                                                    _ptr(_f64(query_completeness_vec)), float(completeness_cutoff),
                                                    int(ref_index_offset), int(n_ref_total), int(root), _ptr(idx),
                                                    _ptr(d0), _ptr(d1), out_loc, C.byref(used)))
        if idx is None or out is not None:
            return None
        return SparseDistanceMatrix(ref_names, query_sketches.names, dist_type, int(used.value), idx, d0, d1, (0, nq))


# ---- the reference's distance drivers ------------------------------------------------------------------
def _out_buffer(ctx, n_elems, dtype, pinned):
    if pinned:
        pa = ctx.pinned((n_elems,), dtype)
        return pa.array, pa
    return np.empty(n_elems, dtype=dtype), None


def self_dists_all(sketches, n=None, dist_type=None, quiet=True, completeness_vec=None, completeness_cutoff=0.64,
                   rows=None, pinned=True):
    """self_dists_all: src/distances/mod.rs:58-130.  `rows=(begin, end)` restricts to a row shard."""
    ctx = sketches.ctx
    n = sketches.number_samples_loaded() if n is None else n
    dist_type = dist_type or DistType(CORE_ACC)
    r0, r1 = (0, n) if rows is None else rows
    npairs = int(ctx.lib.skd_self_pairs_in_rows(n, r0, r1))
    out, pa = _out_buffer(ctx, npairs * dist_type.n_dist_cols, np.float32, pinned)
    ctx._check(ctx.lib.skd_self_dists_all(ctx.handle, sketches.handle, dist_type.c, _ptr(_f64(completeness_vec)),
                                          float(completeness_cutoff), r0, r1, _ptr(out), SKD_HOST))
    return DistanceMatrix(sketches.names, None, dist_type, out, (r0, min(r1, n)), pa)


def cross_dists_all(ref_sketches, query_sketches, n=None, n_query=None, dist_type=None, quiet=True,
                    ref_completeness_vec=None, query_completeness_vec=None, completeness_cutoff=0.64, rows=None,
                    pinned=True):
    """cross_dists_all: src/distances/mod.rs:227-297 (ref-major output)."""
    ctx = ref_sketches.ctx
    n = ref_sketches.number_samples_loaded() if n is None else n
    nq = query_sketches.number_samples_loaded() if n_query is None else n_query
    dist_type = dist_type or DistType(CORE_ACC)
    r0, r1 = (0, n) if rows is None else rows
    out, pa = _out_buffer(ctx, (r1 - r0) * nq * dist_type.n_dist_cols, np.float32, pinned)
    ctx._check(ctx.lib.skd_cross_dists_all(ctx.handle, ref_sketches.handle, query_sketches.handle, dist_type.c,
                                           _ptr(_f64(ref_completeness_vec)), _ptr(_f64(query_completeness_vec)),
                                           float(completeness_cutoff), r0, r1, _ptr(out), SKD_HOST))
    return DistanceMatrix(ref_sketches.names, query_sketches.names, dist_type, out, (r0, r1), pa)


def self_dists_knn(sketches, n=None, knn=50, dist_type=None, quiet=True, completeness_vec=None,
                   completeness_cutoff=0.64, rows=None):
    """self_dists_knn: src/distances/mod.rs:133-224."""
    ctx = sketches.ctx
    n = sketches.number_samples_loaded() if n is None else n
    dist_type = dist_type or DistType(CORE_ACC)
    r0, r1 = (0, n) if rows is None else rows
    idx = np.zeros((r1 - r0, knn), dtype=np.uint64)
    d0 = np.zeros((r1 - r0, knn), dtype=np.float32)
    d1 = np.zeros((r1 - r0, knn), dtype=np.float32) if dist_type.kind == CORE_ACC else None
    ctx._check(ctx.lib.skd_self_dists_knn(ctx.handle, sketches.handle, int(knn), dist_type.c,
                                          _ptr(_f64(completeness_vec)), float(completeness_cutoff), r0, r1, _ptr(idx),
                                          _ptr(d0), _ptr(d1), SKD_HOST))
    return SparseDistanceMatrix(sketches.names, None, dist_type, knn, idx, d0, d1, (r0, r1))


def cross_dists_knn(ref_sketches, query_sketches, n=None, n_query=None, knn=50, dist_type=None, quiet=True,
                    ref_completeness_vec=None, query_completeness_vec=None, completeness_cutoff=0.64, rows=None):
    """cross_dists_knn: src/distances/mod.rs:306-395 (one row per query; knn clamped to n_ref)."""
    ctx = ref_sketches.ctx
    n = ref_sketches.number_samples_loaded() if n is None else n
    nq = query_sketches.number_samples_loaded() if n_query is None else n_query
    dist_type = dist_type or DistType(CORE_ACC)
    r0, r1 = (0, nq) if rows is None else rows
    knn_c = max(1, min(int(knn), max(n, 1)))
    idx = np.zeros((r1 - r0, knn_c), dtype=np.uint64)
    d0 = np.zeros((r1 - r0, knn_c), dtype=np.float32)
    d1 = np.zeros((r1 - r0, knn_c), dtype=np.float32) if dist_type.kind == CORE_ACC else None
    used = C.c_uint64()
    ctx._check(ctx.lib.skd_cross_dists_knn(ctx.handle, ref_sketches.handle, query_sketches.handle, int(knn),
                                           dist_type.c, _ptr(_f64(ref_completeness_vec)),
                                           _ptr(_f64(query_completeness_vec)), float(completeness_cutoff), r0, r1,
                                           _ptr(idx), _ptr(d0), _ptr(d1), SKD_HOST, C.byref(used)))
    return SparseDistanceMatrix(ref_sketches.names, query_sketches.names, dist_type, int(used.value), idx, d0, d1,
                                (r0, r1))


def self_knn_partial(sketches, knn=50, dist_type=None, rows=None, completeness_vec=None, completeness_cutoff=0.64):
    """Symmetric self kNN over triangle rows `rows` (skd_self_knn_partial): partial lists for ALL n rows as
    (keys uint64 [n, knn], aux float32 [n, knn] or None).  Merge the shards' lists with knn_merge."""
    ctx = sketches.ctx
    n = sketches.number_samples_loaded()
    dist_type = dist_type or DistType(CORE_ACC)
    r0, r1 = (0, n) if rows is None else rows
    keys = np.zeros((n, knn), dtype=np.uint64)
    aux = np.zeros((n, knn), dtype=np.float32) if dist_type.kind == CORE_ACC else None
    ctx._check(ctx.lib.skd_self_knn_partial(ctx.handle, sketches.handle, int(knn), dist_type.c,
                                            _ptr(_f64(completeness_vec)), float(completeness_cutoff), r0, r1, _ptr(keys),
                                            _ptr(aux), SKD_HOST))
    return keys, aux


def cross_knn_partial(ref_shard, query_sketches, knn=50, dist_type=None, ref_index_offset=0, ref_completeness_vec=None,
                      query_completeness_vec=None, completeness_cutoff=0.64):
    """Best knn of one reference shard for every query (skd_cross_knn_partial): (keys [n_query, knn], aux)."""
    ctx = ref_shard.ctx
    nq = query_sketches.number_samples_loaded()
    dist_type = dist_type or DistType(CORE_ACC)
    keys = np.zeros((nq, knn), dtype=np.uint64)
    aux = np.zeros((nq, knn), dtype=np.float32) if dist_type.kind == CORE_ACC else None
    ctx._check(ctx.lib.skd_cross_knn_partial(ctx.handle, ref_shard.handle, query_sketches.handle, int(knn), dist_type.c,
                                             _ptr(_f64(ref_completeness_vec)), _ptr(_f64(query_completeness_vec)),
                                             float(completeness_cutoff), int(ref_index_offset), _ptr(keys), _ptr(aux),
                                             SKD_HOST))
    return keys, aux


def knn_merge(ctx, keys_parts, aux_parts, dist_type, ref_names, query_names=None):
    """Merge partial lists (one per rank / shard) into the reference's sparse rows (skd_knn_merge)."""
    keys = np.ascontiguousarray(np.stack(keys_parts), dtype=np.uint64)
    n_parts, n_rows, knn = keys.shape
    aux = None
    if dist_type.kind == CORE_ACC:
        aux = np.ascontiguousarray(np.stack(aux_parts), dtype=np.float32)
    idx = np.zeros((n_rows, knn), dtype=np.uint64)
    d0 = np.zeros((n_rows, knn), dtype=np.float32)
    d1 = np.zeros((n_rows, knn), dtype=np.float32) if dist_type.kind == CORE_ACC else None
    ctx._check(ctx.lib.skd_knn_merge(ctx.handle, _ptr(keys), _ptr(aux), n_parts, n_rows, knn, dist_type.c, SKD_HOST,
                                     _ptr(idx), _ptr(d0), _ptr(d1), SKD_HOST))
    return SparseDistanceMatrix(ref_names, query_names, dist_type, knn, idx, d0, d1, (0, n_rows))


def self_dists_knn_precluster(sketches, skq_bins, skq_names, knn, dist_type, retain_unmatched=None,
                              completeness_vec=None, completeness_cutoff=0.64):
    """self_dists_knn_precluster: src/distances/mod.rs:399-553 (see precluster.py)."""
    import sys
    from . import precluster
    return precluster.self_dists_knn_precluster(sys.modules[__name__], sketches, skq_bins, skq_names, knn, dist_type,
                                                retain_unmatched, completeness_vec, completeness_cutoff)


def pair_dists(ref_sketches, query_sketches, pairs, dist_type=None, ref_completeness_vec=None,
               query_completeness_vec=None, completeness_cutoff=0.64):
    ctx = ref_sketches.ctx
    dist_type = dist_type or DistType(CORE_ACC)
    pairs = np.ascontiguousarray(pairs, dtype=np.uint64).reshape(-1, 2)
    out = np.zeros((len(pairs), dist_type.n_dist_cols), dtype=np.float32)
    ctx._check(ctx.lib.skd_pair_dists(ctx.handle, ref_sketches.handle, query_sketches.handle, dist_type.c,
                                      _ptr(_f64(ref_completeness_vec)), _ptr(_f64(query_completeness_vec)),
                                      float(completeness_cutoff), _ptr(pairs), len(pairs), _ptr(out)))
    return out


def match_counts(ref_sketches, query_sketches, pairs=None, rows=None):
    """Raw per-k bin-match counts: for explicit pairs, or dense [ref rows, n_query, K] from the tiled kernel."""
    ctx = ref_sketches.ctx
    K = len(ref_sketches.kmer_lengths)
    if pairs is not None:
        pairs = np.ascontiguousarray(pairs, dtype=np.uint64).reshape(-1, 2)
        out = np.zeros((len(pairs), K), dtype=np.uint32)
        ctx._check(ctx.lib.skd_match_counts(ctx.handle, ref_sketches.handle, query_sketches.handle, _ptr(pairs),
                                            len(pairs), _ptr(out)))
        return out
    n, nq = ref_sketches.number_samples_loaded(), query_sketches.number_samples_loaded()
    r0, r1 = (0, n) if rows is None else rows
    out = np.zeros((r1 - r0, nq, K), dtype=np.uint32)
    ctx._check(ctx.lib.skd_match_counts_dense(ctx.handle, ref_sketches.handle, query_sketches.handle, r0, r1,
                                              _ptr(out)))
    return out


# host-side mirrors of the reference's file-level types (.skm / .skd / .ski / .skq), bound to the C++ readers
from .host import (HostError, Inverted, MultiSketch, precluster_from_files, read_completeness_file,  # noqa: E402,F401
                   read_subset_names)

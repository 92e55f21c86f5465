"""Candidate-list ("precluster") kNN: host side of the reference's self_dists_knn_precluster
(src/distances/mod.rs:399-553).  The reference gets its candidates from an inverted index (`.ski`,
src/inverted.rs; read and written by the C++ CLI, csrc/host/inverted.cpp); the same candidate sets follow from the
untransposed bins the reference also requires for this mode (`.skq`, u16 per bin, src/inverted.rs:92-100):
sample j is a candidate of sample i iff they hold the same value in at least one bin position
(Inverted::any_shared_bins, src/inverted.rs:259-268).  Distances and the per-row top-k run on the GPU
(skd_self_dists_knn_candidates)."""
import ctypes as C

import numpy as np

RETAIN = {None: 0, "none": 0, "singleton": 1, "bruteforce": 2}


def read_skq(path, n_samples, stride):
    """`.skq`: n_samples x stride little-endian u16 in .ski sample order (src/sketch/sketch_datafile.rs:147-156)."""
    data = np.fromfile(path, dtype="<u2")
    if data.size < n_samples * stride:
        raise ValueError(f"{path} holds {data.size} bins, expected {n_samples * stride}")
    return data[: n_samples * stride].reshape(n_samples, stride)


def any_shared_bins_csr(skq_bins):
    """CSR (offsets, indices) over .ski indices: for every sample the ascending list of samples sharing at
    least one (bin position, value) with it -- itself included, like the reference's bitmap union."""
    skq = np.ascontiguousarray(skq_bins, dtype=np.uint16)
    n, stride = skq.shape
    keys = [np.arange(n, dtype=np.int64) * n + np.arange(n, dtype=np.int64)]  # (i, i)
    for b in range(stride):
        col = skq[:, b]
        order = np.argsort(col, kind="stable")
        vals = col[order]
        starts = np.flatnonzero(np.r_[True, vals[1:] != vals[:-1]])
        ends = np.r_[starts[1:], n]
        for s, e in zip(starts[ends - starts > 1], ends[ends - starts > 1]):
            members = order[s:e].astype(np.int64)
            keys.append((members[:, None] * n + members[None, :]).ravel())
    allk = np.unique(np.concatenate(keys))
    rows, cols = allk // n, allk % n
    off = np.zeros(n + 1, dtype=np.uint64)
    np.add.at(off, rows + 1, 1)
    return np.cumsum(off).astype(np.uint64), cols.astype(np.uint64)


def self_dists_knn_precluster(sk, sketches, skq_bins, skq_names, knn, dist_type, retain_unmatched=None,
                              completeness_vec=None, completeness_cutoff=0.64):
    """Mirror of self_dists_knn_precluster.  `sk` is the sketchlib.rust_b200 module (avoids a circular import),
    `skq_names` the sample names in .ski/.skq order."""
    ctx = sketches.ctx
    n = sketches.number_samples_loaded()
    lookup = {name: i for i, name in enumerate(skq_names)}
    missing = [nm for nm in sketches.names[:n] if nm not in lookup]
    if missing:  # mod.rs:429-431
        raise sk.SketchDistError(1, f"The following samples in the .skd could not be found in the .ski:\n{missing}")
    skd_to_ski = np.array([lookup[nm] for nm in sketches.names[:n]], dtype=np.int64)
    ski_to_skd = np.full(len(skq_names), -1, dtype=np.int64)
    ski_to_skd[skd_to_ski] = np.arange(n)
    off_ski, idx_ski = any_shared_bins_csr(skq_bins)
    retain = RETAIN[retain_unmatched]
    offs, idxs = [0], []
    everyone = np.arange(n, dtype=np.uint64)
    for i in range(n):
        s = skd_to_ski[i]
        cand = ski_to_skd[idx_ski[int(off_ski[s]):int(off_ski[s + 1])].astype(np.int64)]
        cand = np.sort(cand[cand >= 0]).astype(np.uint64)
        if retain == 2 and not (cand != i).any():
            cand = everyone  # Bruteforce: mod.rs:501-524
        idxs.append(cand)
        offs.append(offs[-1] + len(cand))
    cand_off = np.array(offs, dtype=np.uint64)
    cand_idx = np.concatenate(idxs) if idxs else np.zeros(0, dtype=np.uint64)
    idx = np.zeros((n, knn), dtype=np.uint64)
    d0 = np.zeros((n, knn), dtype=np.float32)
    cv = None if completeness_vec is None else np.ascontiguousarray(completeness_vec, dtype=np.float64)
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    ctx._check(ctx.lib.skd_self_dists_knn_candidates(ctx.handle, sketches.handle, int(knn), dist_type.c, p(cv),
                                                     float(completeness_cutoff), p(cand_off), p(cand_idx), retain,
                                                     p(idx), p(d0)))
    return sk.SparseDistanceMatrix(sketches.names, None, dist_type, knn, idx, d0, None, (0, n))

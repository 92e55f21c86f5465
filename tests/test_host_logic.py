"""CPU-side tests: host logic of the package, the C-ABI library's exported symbols, the synthetic data
generator, and the multi-rank sharding logic over gloo (world_size 2).  No GPU needed."""
import os
import re
import socket

import numpy as np
import pytest

import oracle
from helpers import ROOT, synth_db


def test_library_builds_loads_and_exports_every_declared_symbol():
    import sketchlib.rust_b200 as sk
    lib = sk.load_library()
    header = open(os.path.join(ROOT, "include", "sketchdist.h")).read()
    declared = set(re.findall(r"\b(skd_[a-z0-9_]+)\s*\(", header))
    declared -= {"skd_status", "skd_loc", "skd_dist_kind", "skd_dist_type"}
    assert len(declared) >= 28
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/sketchdist.h but not exported"
    assert set(lib._signatures) == declared  # the ctypes layer binds exactly the declared ABI


def test_no_cpu_fallback_without_gpu():
    import sketchlib.rust_b200 as sk
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    with pytest.raises(sk.SketchDistError, match="no CPU fallback"):
        sk.Context(0)


def test_pure_helpers_need_no_device():
    import sketchlib.rust_b200 as sk
    lib = sk.load_library()
    for n in (2, 5, 64, 65, 1000, 10000):
        assert lib.skd_self_pairs_in_rows(n, 0, n) == n * (n - 1) // 2
        for r0, r1 in ((0, 1), (1, n), (n // 2, n // 2 + 1), (3, 2)):
            want = sum(n - 1 - i for i in range(r0, min(r1, n)))
            assert lib.skd_self_pairs_in_rows(n, r0, r1) == want
        for world in (1, 2, 3, 4, 8):
            tri = [sk.triangle_row_partition(n, r, world) for r in range(world)]
            assert tri[0][0] == 0 and tri[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(tri, tri[1:]))
            ev = [sk.row_partition(n, r, world) for r in range(world)]
            assert ev[0][0] == 0 and ev[-1][1] == n and all(a[1] == b[0] for a, b in zip(ev, ev[1:]))
            assert max(b - a for a, b in ev) - min(b - a for a, b in ev) <= 1
    # equal pair counts to within one 64-row tile band
    n, world = 100000, 8
    counts = [lib.skd_self_pairs_in_rows(n, *sk.triangle_row_partition(n, r, world)) for r in range(world)]
    assert max(counts) / min(counts) < 1.02


def test_text_rendering_matches_rust_display():
    import sketchlib.rust_b200 as sk
    assert [sk.fmt_f32(x) for x in (1.0, 0.0, 0.008820182, 0.2137754, 0.35546875, 0.99009615, 1e-7)] == \
        ["1", "0", "0.008820182", "0.2137754", "0.35546875", "0.99009615", "0.0000001"]
    names = ["a", "b", "c"]
    dm = sk.DistanceMatrix(names, None, sk.DistType(sk.CORE_ACC), np.array([0.5, 1, 0.25, 0, 1, 1], np.float32), (0, 3))
    assert str(dm) == "a\tb\t0.5\t1\na\tc\t0.25\t0\nb\tc\t1\t1\n"
    dmx = sk.DistanceMatrix(names, ["q"], sk.DistType(sk.JACCARD, 0, 21), np.array([0.5, 1, 0.25], np.float32), (0, 3))
    assert str(dmx) == "a\tq\t0.5\nb\tq\t1\nc\tq\t0.25\n"
    sp = sk.SparseDistanceMatrix(names, None, sk.DistType(sk.JACCARD, 0, 21), 2, np.array([[1, 0], [0, 2], [1, 2]], np.uint64),
                                 np.array([[0.5, 1.0], [0.5, 0.75], [0.75, 1.0]], np.float32), None, (0, 3))
    # padding entries (dist 1, self) are skipped: distance_matrix.rs:376-381
    assert str(sp) == "a\tb\t0.5\nb\ta\t0.5\nb\tc\t0.75\nc\tb\t0.75\n"
    assert str(sk.DistType(sk.CORE_ACC)) == "Distances: core/accessory regression"
    assert str(sk.DistType(sk.ANI, 0, 21)) == "Distances: ANI at k=21"


def test_synth_bit_slice_matches_fill_usigs():
    import torch
    from sketchlib.rust_b200 import synth
    rng = np.random.default_rng(3)
    bins = rng.integers(0, 1 << 14, size=(5, 3, 128), dtype=np.uint16)
    got = synth.as_u64_numpy(synth.bit_slice(torch.from_numpy(bins.astype(np.int16))))
    assert np.array_equal(got, oracle.fill_usigs(bins))
    flat = synth.as_u64_numpy(synth.make_bins(130, 2, [17, 21, 25], seed=1, family_mean=10))
    assert flat.size == 130 * 3 * 2 * 14
    db = oracle.Db(flat, 130, [17, 21, 25], 2)
    d = oracle.self_dists_all(db, k_idx=0)
    assert (d < 0.9).any() and (d > 0.9).any()  # families are visible: both related and unrelated pairs


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _shard_worker(rank, world, port, tmp):
    import torch.distributed as dist
    import sketchlib.rust_b200 as sk
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    db = synth_db(150, 2, [17, 21, 25], seed=4, n_families=3)
    lib = sk.load_library()
    # dense: this rank's triangle row block, computed by the CPU oracle standing in for the device
    r0, r1 = sk.triangle_row_partition(db.n, rank, world, align=16)
    first = int(lib.skd_self_pairs_in_rows(db.n, 0, r0))
    count = int(lib.skd_self_pairs_in_rows(db.n, r0, r1))
    chunk0 = first - first % 1000
    mine = oracle.self_dists_all(db, first=chunk0, count=count + first - chunk0)[2 * (first - chunk0):]
    # kNN: even row split
    k0, k1 = sk.row_partition(db.n, rank, world)
    idx, d0, _ = oracle.self_dists_knn(db, 5, k_idx=1, rows=(k0, k1))
    # symmetric kNN the way the peer group does it: triangle rows per rank (skd_triangle_rows) -> partial lists for ALL
    # rows (each unordered pair evaluated once, credited to both rows) -> the owner of a row (skd_slice_rows) merges
    import ctypes as C
    a, b = C.c_uint64(), C.c_uint64()
    lib.skd_triangle_rows(db.n, rank, world, C.byref(a), C.byref(b))
    t0, t1 = a.value, b.value
    knn = 5
    tfirst, tcount = int(lib.skd_self_pairs_in_rows(db.n, 0, t0)), int(lib.skd_self_pairs_in_rows(db.n, t0, t1))
    c0 = tfirst - tfirst % 1000
    vals = oracle.self_dists_all(db, k_idx=1, first=c0, count=tcount + tfirst - c0)[tfirst - c0:]
    cand = [[] for _ in range(db.n)]
    p = 0
    for i in range(t0, t1):
        for j in range(i + 1, db.n):
            cand[i].append((float(vals[p]), j))
            cand[j].append((float(vals[p]), i))
            p += 1
    assert p == tcount
    partial = [sorted(c)[:knn] for c in cand]
    lists = [None] * world
    dist.all_gather_object(lists, partial)
    lib.skd_slice_rows(db.n, rank, world, 1, C.byref(a), C.byref(b))
    merged = {r: sorted(sum((lists[q][r] for q in range(world)), []))[:knn] for r in range(a.value, b.value)}
    gathered = [None] * world
    dist.all_gather_object(gathered, (r0, r1, mine, k0, k1, idx, d0, merged))
    if rank == 0:
        sym = {}
        for g in gathered:
            sym.update(g[7])
        np.save(os.path.join(tmp, "sym_d0.npy"), np.array([[d for d, _ in sym[r]] for r in range(db.n)], dtype=np.float32))
        np.save(os.path.join(tmp, "dense.npy"), np.concatenate([g[2] for g in gathered]))
        np.save(os.path.join(tmp, "knn_idx.npy"), np.concatenate([g[5] for g in gathered]))
        assert gathered[0][0] == 0 and gathered[-1][1] == db.n and gathered[0][3] == 0 and gathered[-1][4] == db.n
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_over_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_shard_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    db = synth_db(150, 2, [17, 21, 25], seed=4, n_families=3)
    assert np.array_equal(np.load(tmp_path / "dense.npy"), oracle.self_dists_all(db))
    assert np.array_equal(np.load(tmp_path / "knn_idx.npy"), oracle.self_dists_knn(db, 5, k_idx=1)[0])
    assert np.array_equal(np.load(tmp_path / "sym_d0.npy"), oracle.self_dists_knn(db, 5, k_idx=1)[1])


def test_bench_workload_sizes():
    import bench
    import sketchlib.rust_b200 as sk
    assert bench.N_KNN == 100000 and bench.KNN == 50                      # BASELINE.json configs[2], every N
    assert bench.dense_n_for_gpus(1) == 10000 and bench.LOP3_PER_PAIR == 3584 and bench.POPC_PER_PAIR == 256
    for g in (2, 4, 8):
        n = bench.dense_n_for_gpus(g)
        assert n % 64 == 0 and abs(n * (n - 1) / 2 / g - 49995000) / 49995000 < 0.02
        # strong-scaled kNN job: the ranks' triangle blocks tile the rows and hold about equal pair counts
        lib = sk.load_library()
        parts = [sk.triangle_row_partition(bench.N_KNN, r, g) for r in range(g)]
        assert parts[0][0] == 0 and parts[-1][1] == bench.N_KNN and all(parts[i][1] == parts[i + 1][0] for i in range(g - 1))
        pairs = [int(lib.skd_self_pairs_in_rows(bench.N_KNN, a, b)) for a, b in parts]
        assert sum(pairs) == bench.N_KNN * (bench.N_KNN - 1) // 2 and max(pairs) / (sum(pairs) / g) < 1.006


def test_python_multisketch_and_inverted_mirror_the_reference_types(tmp_path):
    """sketchlib.rust_b200.MultiSketch / Inverted (libsketchhost.so: the CLI's C++ readers behind ctypes) against the
    oracle's independent Python reader on the reference's fixtures: metadata, legacy fix-up, subset order, completeness
    files, index queries.  No GPU involved."""
    import oracle
    import sketchlib.rust_b200 as sk
    from helpers import FIXTURES
    for name in ("sketches1", "sketches2", "sketches3", "legacy_db"):
        ms = sk.MultiSketch.load_metadata(os.path.join(FIXTURES, name + ".skm"))
        want = oracle.MultiSketch(os.path.join(FIXTURES, name))
        assert ms.kmer_lengths == list(want.kmer_lengths) and ms.sketch_size == want.sketch_size
        assert (ms.sketchsize64, ms.sample_stride, ms.kmer_stride) == (want.sketchsize64, want.sample_stride, want.kmer_stride)
        assert ms.names() == want.names() and ms.number_samples_loaded() == len(want.names())
        assert ms.file_indices().tolist() == [m["index"] for m in want.sketch_metadata]
        assert ms.hash_type == "DNA" and ms.get_k_idx(ms.kmer_lengths[-1]) == len(ms.kmer_lengths) - 1 and ms.get_k_idx(4) is None
    # --subset: loaded order = list order (read_sketch_data_block), unknown names are the reference's panic
    ms = sk.MultiSketch(os.path.join(FIXTURES, "sketches3"))
    names = sk.read_subset_names(os.path.join(FIXTURES, "subset.txt"))
    ms.select_subset(names)
    assert ms.names() == names and ms.number_samples_loaded() == len(names)
    all_names = sk.MultiSketch(os.path.join(FIXTURES, "sketches3")).names()
    assert ms.file_indices().tolist() == [all_names.index(n) for n in names]
    with pytest.raises(sk.HostError, match="Could not find requested sample"):
        ms.select_subset(["nope"])
    with pytest.raises(sk.HostError):
        sk.MultiSketch(str(tmp_path / "missing"))
    # completeness file of the reference's tests: values by sample name, default 1.0 with a warning
    full = sk.MultiSketch(os.path.join(FIXTURES, "sketches3"))
    vals, warns = sk.read_completeness_file(os.path.join(FIXTURES, "completeness.txt"), full)
    listed = dict(l.split("\t") for l in open(os.path.join(FIXTURES, "completeness.txt")).read().splitlines() if "\t" in l)
    assert [v for v in vals] == [float(listed.get(n, 1.0)) for n in full.names()]
    # .ski: candidates and match counts equal the definition on the .skq bins
    inv = sk.Inverted.load(os.path.join(FIXTURES, "reordered_inv.ski"))
    skq = inv.read_skq()
    assert (inv.n, inv.stride, inv.kmer, inv.rc) == (4, 10, 21, True)
    assert inv.sample_names == ["TIGR4.fa.gz", "R6.fa.gz", "14412_3#82.contigs_velvet.fa.gz", "14412_3#84.contigs_velvet.fa.gz"]
    for s in range(4):
        counts = (skq == skq[s][None, :]).sum(axis=1)
        assert inv.query_against_inverted_index(skq[s]).tolist() == counts.tolist()
        assert inv.any_shared_bins(skq[s]).tolist() == np.flatnonzero(counts > 0).tolist()
    assert inv.any_shared_bin_pairs() == 2


def test_matrix_writers_match_python_display(tmp_path):
    """DistanceMatrix.write / SparseDistanceMatrix.write (C++, threaded) produce exactly str(matrix): self / cross / row
    shards, one and two columns, sparse rows with padding entries, tricky floats."""
    import sketchlib.rust_b200 as sk
    rng = np.random.default_rng(3)
    n, nq = 77, 5
    names, qnames = [f"s{i}#x" for i in range(n)], [f"q{i}" for i in range(nq)]
    tricky = np.array([0.0, 1.0, 1e-7, 0.1, 1 / 3, 0.008820182, 123456.7, 3.4e38, 1e-45, 0.99999994], dtype=np.float32)
    for ncols, kind in ((2, sk.CORE_ACC), (1, sk.JACCARD)):
        dt = sk.DistType(kind, 0)
        for r0, r1 in ((0, n), (13, 50), (76, 77)):
            cnt = sum(n - 1 - i for i in range(r0, r1))
            d = rng.random(cnt * ncols).astype(np.float32)
            d[: min(d.size, tricky.size)] = tricky[: min(d.size, tricky.size)]
            m = sk.DistanceMatrix(names, None, dt, d, (r0, r1))
            m.write(str(tmp_path / "a.tsv"), threads=3)
            assert open(tmp_path / "a.tsv").read() == str(m)
        d = rng.random((40 - 7) * nq * ncols).astype(np.float32)
        m = sk.DistanceMatrix(names, qnames, dt, d, (7, 40))
        m.write(str(tmp_path / "b.tsv"), threads=2)
        assert open(tmp_path / "b.tsv").read() == str(m)
        m.write(str(tmp_path / "b.tsv"), append=True)
        assert open(tmp_path / "b.tsv").read() == str(m) * 2
        knn = 4
        idx = rng.integers(0, n, size=(30, knn)).astype(np.uint64)
        d0 = rng.random((30, knn)).astype(np.float32)
        idx[3, 2:], d0[3, 2:] = 10 + 3, 1.0          # padding entries of row 3 (its own index, distance 1)
        d1 = rng.random((30, knn)).astype(np.float32) if ncols == 2 else None
        sp = sk.SparseDistanceMatrix(names, None, dt, knn, idx, d0, d1, (10, 40))
        sp.write(str(tmp_path / "c.tsv"), threads=4)
        assert open(tmp_path / "c.tsv").read() == str(sp)
        spx = sk.SparseDistanceMatrix(names, qnames, dt, knn, idx[:5], d0[:5], None if d1 is None else d1[:5], (0, 5))
        spx.write(str(tmp_path / "d.tsv"))
        assert open(tmp_path / "d.tsv").read() == str(spx)
    with pytest.raises(sk.HostError):
        sp.write(str(tmp_path / "no_such_dir" / "x.tsv"))


def test_weighted_triangle_split():
    """The split skd_comm_self_dists_knn uses when it re-balances the ranks from their measured times: equal shares give
    skd_triangle_rows; any shares give a partition of [0, n) on tile boundaries whose pair counts follow the shares."""
    import sketchlib.rust_b200 as sk
    lib = sk.load_library()
    for n in (1, 63, 64, 1000, 10_000, 100_000, 1_234_567):
        for world in (1, 2, 3, 8, 16):
            b = sk.triangle_rows_weighted(n, [1.0 / world] * world)
            assert b[0] == 0 and b[-1] == n and all(x <= y for x, y in zip(b, b[1:]))
            for r in range(world):
                assert (b[r], b[r + 1]) == sk.triangle_row_partition(n, r, world)
    rng = np.random.default_rng(5)
    n = 100_000
    total = n * (n - 1) // 2
    for world in (2, 5, 8):
        w = rng.uniform(0.5, 1.5, size=world)
        shares = w / w.sum()
        b = sk.triangle_rows_weighted(n, shares)
        assert b[0] == 0 and b[-1] == n and all(x < y for x, y in zip(b, b[1:]))
        assert all(x % 64 == 0 for x in b[1:-1])
        for r in range(world):
            pairs = int(lib.skd_self_pairs_in_rows(n, b[r], b[r + 1]))
            assert abs(pairs / total - shares[r]) < 2e-3   # off by at most a tile row or two

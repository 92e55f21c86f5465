"""Peer groups (skd_comm_*): several ranks working on one problem, exchanging partial kNN lists and database slices
through each other's device memory.  The ranks run as host threads of this process -- on different GPUs when the box
has several, otherwise sharing GPU 0 (same code path: peers' windows are plain device pointers) -- and as separate
processes that map each other's windows with CUDA IPC.  Results must equal the single-context calls bit for bit and
the oracle within the path's tolerance."""
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

import oracle
from helpers import synth_db, synth_pair, write_skm_skd
from test_gpu_parity import assert_knn_matches

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def sk():
    import sketchlib.rust_b200 as sk
    return sk


def n_devices():
    out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout
    return max(1, sum(1 for ln in out.splitlines() if ln.startswith("GPU ")))


def run_group(sk, world, window_bytes, body):
    """Runs body(rank, ctx, grp) on `world` threads after connecting the group; returns the per-rank results.
    Windows are only destroyed after every rank has finished (a failing rank must not pull its memory from under
    the peers), and every rank's error is reported."""
    os.environ.setdefault("SKD_PEER_TIMEOUT_S", "30")
    ndev = n_devices()
    handles, results, errors = [None] * world, [None] * world, [None] * world
    owned = [None] * world
    gate = threading.Barrier(world)

    def worker(rank):
        try:
            ctx = sk.Context(rank % ndev)
            owned[rank] = [ctx, None]
            grp = sk.PeerGroup(ctx, rank, world, window_bytes)
            owned[rank][1] = grp
            handles[rank] = grp.export()
            gate.wait(timeout=120)
            grp.connect(handles)
            results[rank] = body(rank, ctx, grp)
            grp.barrier()
        except Exception as e:  # noqa: BLE001
            errors[rank] = e
            gate.abort()

    threads = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=600)
    for o in owned:
        if o is not None:
            if o[1] is not None:
                o[1].close()
            o[0].close()
    if any(e is not None for e in errors):
        raise RuntimeError("; ".join(f"rank {r}: {type(e).__name__}: {e}" for r, e in enumerate(errors) if e is not None))
    return results


@pytest.mark.parametrize("world", [2, 3])
def test_peer_group_self_knn_equals_single_device(sk, world):
    db = synth_db(1100, 16, [15, 17, 19, 21, 23, 25, 27, 29], seed=31, n_families=12, unrelated_frac=0.05)
    per = db.sample_stride
    win = max(sk.PeerGroup.window_bytes_knn(db.n, 70, world), sk.PeerGroup.window_bytes_db(db.n, db.W, len(db.kmers), world))

    def body(rank, ctx, grp):
        s0, s1 = grp.my_sample_slice(db.n)
        g = grp.upload(np.ascontiguousarray(db.bins[s0 * per:s1 * per]), db.n, db.kmers, db.W, names=db.names)
        # the gathered database is the uploaded one: raw match counts of a few pairs from every slice
        pairs = np.array([(0, db.n - 1), (s0, (s1 - 1) % db.n), (db.n // 2, 3), (db.n - 1, db.n - 2)], dtype=np.uint64)
        assert np.array_equal(sk.match_counts(g, g, pairs=pairs), oracle.match_counts(db, db, pairs))
        out = {}
        out["ca"] = grp.self_dists_knn(g, knn=50, dist_type=sk.set_k(g), root=0)          # symmetric partial lists
        out["jac"] = grp.self_dists_knn(g, knn=9, dist_type=sk.set_k(g, 21), root=-1)      # every rank gets the rows
        out["ani70"] = grp.self_dists_knn(g, knn=70, dist_type=sk.set_k(g, 21, ani=True), root=world - 1)  # row split
        if rank == 0:
            out["single_ca"] = sk.self_dists_knn(g, knn=50, dist_type=sk.set_k(g))
            out["single_jac"] = sk.self_dists_knn(g, knn=9, dist_type=sk.set_k(g, 21))
            out["single_ani70"] = sk.self_dists_knn(g, knn=70, dist_type=sk.set_k(g, 21, ani=True))
        out["xchg_ms"] = grp.last_exchange_ms
        # repeated calls of one problem shape re-split the triangle from the ranks' measured times: the blocks may
        # move (and stay a partition of the rows), the rows must not change
        blocks = []
        for _ in range(3):
            again = grp.self_dists_knn(g, knn=50, dist_type=sk.set_k(g), root=0)
            blocks.append(grp.last_row_block)
            if rank == 0:
                assert np.array_equal(again.idx, out["ca"].idx) and np.array_equal(again.d0, out["ca"].d0)
                assert np.array_equal(again.d1, out["ca"].d1)
        out["blocks"] = blocks
        g.free()
        return out

    res = run_group(sk, world, win, body)
    r0 = res[0]
    for call in range(3):   # every call's blocks tile [0, n) in rank order, on 64-sample boundaries
        edges = [res[r]["blocks"][call] for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == db.n
        assert all(edges[r][1] == edges[r + 1][0] and edges[r][1] % 64 == 0 for r in range(world - 1))
    assert all(r["ca"] is None for r in res[1:]) and all(r["ani70"] is None for r in res[:-1])
    for a, b in ((r0["ca"], r0["single_ca"]), (r0["jac"], r0["single_jac"]), (res[-1]["ani70"], r0["single_ani70"])):
        assert np.array_equal(a.idx, b.idx) and np.array_equal(a.d0, b.d0)
        assert (a.d1 is None) == (b.d1 is None) and (a.d1 is None or np.array_equal(a.d1, b.d1))
    for r in res[1:]:
        assert np.array_equal(r["jac"].idx, r0["jac"].idx) and np.array_equal(r["jac"].d0, r0["jac"].d0)
    oidx, od0, od1 = oracle.self_dists_knn(db, 50, threads=8)
    assert_knn_matches(r0["ca"].idx, r0["ca"].d0, r0["ca"].d1, oidx, od0, od1, False)
    oidx, od0, _ = oracle.self_dists_knn(db, 9, k_idx=3, threads=8)
    assert_knn_matches(r0["jac"].idx, r0["jac"].d0, None, oidx, od0, None, True)


def test_peer_group_cross_knn_reference_shards(sk):
    world = 3
    ref, qry = synth_pair(1300, 150, 16, [17, 21, 25, 29], seed=33, n_families=14)
    per = ref.sample_stride
    win = sk.PeerGroup.window_bytes_knn(qry.n, 12, world)

    def body(rank, ctx, grp):
        a, b = sk.slice_rows(ref.n, rank, world, 64)
        shard = sk.SketchDb.upload(ctx, np.ascontiguousarray(ref.bins[a * per:b * per]), b - a, ref.kmers, ref.W)
        gq = sk.SketchDb.upload(ctx, qry.bins, qry.n, qry.kmers, qry.W, names=qry.names)
        out = {"ca": grp.cross_dists_knn(shard, gq, 12, sk.set_k(gq), a, ref.n, ref_names=ref.names, root=0),
               "jac": grp.cross_dists_knn(shard, gq, 5000, sk.set_k(gq, 21), a, ref.n, ref_names=ref.names, root=-1)}
        shard.free()
        gq.free()
        return out

    res = run_group(sk, world, max(win, sk.PeerGroup.window_bytes_knn(qry.n, ref.n, world)), body)
    ca = res[0]["ca"]
    oidx, od0, od1 = oracle.cross_dists_knn(ref, qry, 12, threads=8)
    assert_knn_matches(ca.idx, ca.d0, ca.d1, oidx, od0, od1, False)
    jac = res[1]["jac"]
    assert jac.knn == ref.n and jac.idx.shape == (qry.n, ref.n)      # knn clamped to the number of references (mod.rs:325)
    assert all(sorted(r.tolist()) == list(range(ref.n)) for r in jac.idx[:5])
    oidx, od0, _ = oracle.cross_dists_knn(ref, qry, ref.n, k_idx=1, threads=8)
    assert np.array_equal(jac.d0, od0)


def test_peer_group_load_skd_slices(sk, tmp_path):
    world = 2
    db = synth_db(333, 3, [17, 21], seed=35, n_families=4)
    prefix = write_skm_skd(str(tmp_path / "db"), db)
    order = np.random.default_rng(5).permutation(db.n)[:200].astype(np.uint64)      # a --subset order
    win = max(sk.PeerGroup.window_bytes_knn(200, 5, world), sk.PeerGroup.window_bytes_db(db.n, db.W, 2, world))

    def body(rank, ctx, grp):
        full = grp.load_skd(prefix + ".skd", db.n, db.kmers, db.W)
        sub = grp.load_skd(prefix + ".skd", len(order), db.kmers, db.W, sample_indices=order)
        pairs = np.array([(i, (7 * i + 1) % db.n) for i in range(0, db.n, 11)], dtype=np.uint64)
        assert np.array_equal(sk.match_counts(full, full, pairs=pairs), oracle.match_counts(db, db, pairs))
        r = grp.self_dists_knn(sub, knn=5, dist_type=sk.set_k(sub, 21), root=0)
        full.free()
        sub.free()
        return r

    res = run_group(sk, world, win, body)
    sub_db = oracle.Db(db.bins.reshape(db.n, -1)[order.astype(int)].reshape(-1), len(order), db.kmers, db.W)
    oidx, od0, _ = oracle.self_dists_knn(sub_db, 5, k_idx=1, threads=4)
    assert_knn_matches(res[0].idx, res[0].d0, None, oidx, od0, None, True)


def test_peer_group_argument_errors(sk):
    ctx = sk.Context(0)
    with pytest.raises(sk.SketchDistError):
        sk.PeerGroup(ctx, 2, 2, 1 << 20)                      # rank outside the group
    grp = sk.PeerGroup(ctx, 0, 2, 1 << 20)
    db = synth_db(100, 2, [17, 21], seed=1)
    g = sk.SketchDb.upload(ctx, db.bins, db.n, db.kmers, db.W)
    with pytest.raises(sk.SketchDistError, match="not connected"):
        grp.self_dists_knn(g, knn=5, dist_type=sk.set_k(g, 21))
    grp.close()
    solo = sk.PeerGroup(ctx, 0, 1, 1 << 12)                   # a group of one needs no peers
    with pytest.raises(sk.SketchDistError, match="too small"):
        solo.self_dists_knn(g, knn=5, dist_type=sk.set_k(g, 21))
    solo.close()
    solo = sk.PeerGroup(ctx, 0, 1, sk.PeerGroup.window_bytes_knn(100, 5, 1))
    a = solo.self_dists_knn(g, knn=5, dist_type=sk.set_k(g, 21))
    b = sk.self_dists_knn(g, knn=5, dist_type=sk.set_k(g, 21))
    assert np.array_equal(a.idx, b.idx) and np.array_equal(a.d0, b.d0)
    solo.close()
    g.free()
    ctx.close()


def test_peer_group_processes_ipc(sk, tmp_path):
    """One process per rank, windows mapped with cudaIpcOpenMemHandle (how bench.py and torchrun launches use it)."""
    world = 2
    ndev = n_devices()
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "peer_worker.py"), str(r), str(world), str(r % ndev),
                               str(tmp_path)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(world)]
    logs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            p.kill()
            out, _ = p.communicate()
        logs.append(out)
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    db = synth_db(900, 16, [15, 19, 23, 27], seed=77, n_families=9, unrelated_frac=0.05)
    got = np.load(tmp_path / "out0.npz")
    other = np.load(tmp_path / "out1.npz")
    oidx, od0, od1 = oracle.self_dists_knn(db, 20, threads=8)
    assert_knn_matches(got["idx"], got["d0"], got["d1"], oidx, od0, od1, False)
    oidx, od0, _ = oracle.self_dists_knn(db, 7, k_idx=1, threads=8)
    assert_knn_matches(got["jidx"], got["jd0"], None, oidx, od0, None, True)
    assert np.array_equal(got["jidx"], other["jidx"]) and np.array_equal(got["jd0"], other["jd0"])

"""One rank of a peer group in its own process (tests/test_gpu_peer.py::test_peer_group_processes_ipc):
python peer_worker.py RANK WORLD DEVICE DIR.  Handles are exchanged through files in DIR; rank 0 saves the rows."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import sketchlib.rust_b200 as sk  # noqa: E402
from helpers import synth_db  # noqa: E402


def main():
    rank, world, device, d = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    db = synth_db(900, 16, [15, 19, 23, 27], seed=77, n_families=9, unrelated_frac=0.05)
    ctx = sk.Context(device)
    win = max(sk.PeerGroup.window_bytes_knn(db.n, 20, world), sk.PeerGroup.window_bytes_db(db.n, db.W, len(db.kmers), world))
    grp = sk.PeerGroup(ctx, rank, world, win)
    with open(os.path.join(d, f"h{rank}.tmp"), "wb") as fh:
        fh.write(grp.export())
    os.rename(os.path.join(d, f"h{rank}.tmp"), os.path.join(d, f"h{rank}.bin"))
    handles = []
    for r in range(world):
        p = os.path.join(d, f"h{r}.bin")
        t0 = time.time()
        while not os.path.exists(p):
            if time.time() - t0 > 120:
                raise SystemExit(f"rank {rank}: no handle from rank {r}")
            time.sleep(0.01)
        handles.append(open(p, "rb").read())
    grp.connect(handles)
    s0, s1 = grp.my_sample_slice(db.n)
    per = db.sample_stride
    g = grp.upload(np.ascontiguousarray(db.bins[s0 * per:s1 * per]), db.n, db.kmers, db.W, names=db.names)
    for rep in range(3):   # several collective calls back to back: windows are reused safely
        res = grp.self_dists_knn(g, knn=20, dist_type=sk.set_k(g), root=0)
    resj = grp.self_dists_knn(g, knn=7, dist_type=sk.set_k(g, 19), root=-1)
    np.savez(os.path.join(d, f"out{rank}.npz"), jidx=resj.idx, jd0=resj.d0,
             **({"idx": res.idx, "d0": res.d0, "d1": res.d1} if rank == 0 else {}))
    grp.barrier()
    g.free()
    grp.close()
    ctx.close()


if __name__ == "__main__":
    main()

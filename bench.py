#!/usr/bin/env python
"""Benchmark of the pairwise sketch-distance path (BASELINE.json metric: sketch pairs/sec, core/accessory,
1024 bins, 8 k-mer lengths, at 1/2/4/8 B200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...        (one rank per GPU)

Workload at EVERY N (strong scaling): BASELINE.json configs[2], the config north_star's target is written on --
n = 100 000 synthetic genomes (mutated families), 1024-bin sketches, k = 15..29 step 2 (K = 8), all-vs-all
core/accessory with --knn 50 sparse output (the reference's self_dists_knn, src/distances/mod.rs:133-224).
A "step" is one whole job: 4 999 950 000 unique pairs.  The reference evaluates every ORDERED pair; here every
unordered pair is evaluated once and credited to both rows, so `value` counts unique pairs per second.

The N ranks form a peer group of the library (skd_comm_*, include/sketchdist.h): each rank evaluates an
equal-pair block of upper-triangle rows, the owner of a row merges its partial lists straight out of the peers'
HBM over NVLink (the library's own kernels; no NCCL on the data path), the final rows are gathered on rank 0.
torch.distributed (NCCL) is used only for the timing barrier, the max over ranks and to pass the 128-byte window
handles around.

Prints ONE JSON line (rank 0).
  value  unique pairs/s, database resident in every rank's HBM, final sparse rows left in rank 0's HBM; CUDA events on
         the library's compute stream from the entry barrier to the gathered rows, max over ranks.
  e2e    the same job through the public host API with HOST buffers: every rank uploads its 1/N slice of the .skd-layout
         bit planes from pinned host memory (skd_comm_db_upload; the other slices arrive from the peers over NVLink),
         the distances, exchange and merge run, and rank 0 copies the final rows (n x 50 x 16 B) to host memory.
  roofline       rank 0's tile_kernel launches against the live-measured LOP3 issue peak (tools/int_peak).
  parity_sample  after the timed region: random rows of the e2e result against the CPU oracle.
  dense_10k      secondary leg: BASELINE.json configs[1] (dense all-vs-all core/accessory, 10 000 genomes; weak-scaled
                 n = 10 000 sqrt(N) for N > 1), last round's headline.
  roofline_hbm   (N = 1) the HBM-bound side of the path: one query against 200 000 references.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

KMERS = list(range(15, 31, 2))  # 15..29 step 2 -> 8 k-mer lengths
W = 16                          # 1024 bins
N_KNN = 100_000                 # BASELINE.json configs[2]
KNN = 50
N_DENSE = 10_000                # BASELINE.json configs[1]
LOP3_PER_PAIR = len(KMERS) * W * 28   # K * W * (28 LOP3 + 2 POPC) per pair
POPC_PER_PAIR = len(KMERS) * W * 2
METRIC = "sketch pairs/sec (core/accessory, 1024 bins, 8 k)"
PER_SAMPLE = len(KMERS) * W * 14  # u64 words per sample in the .skd layout


def dense_n_for_gpus(n_gpus):
    return int(round(N_DENSE * n_gpus ** 0.5 / 64.0)) * 64 if n_gpus > 1 else N_DENSE


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, power, reasons = [], [], [], set()
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9 or not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_min_mhz": min(sm), "sm_max_mhz": max(mx),
                "power_w_max": max(power) if power else None, "reasons": sorted(reasons), "samples": len(sm)}


def int_peaks():
    """LOP3/POPC issue peaks measured live on this GPU by tools/int_peak (built by __graft_entry__.build)."""
    exe = os.path.join(ROOT, "tools", "int_peak")
    if not os.path.exists(exe):
        try:
            subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-o", exe, exe + ".cu", "-ldl"])
        except Exception:
            return None
    try:
        out = subprocess.run([exe, "0.5"], capture_output=True, text=True, timeout=120).stdout
        return json.loads(out)
    except Exception:
        return None


def measured_hbm():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        return 6650.0


def knn_workload_config(n_gpus):
    n = N_KNN
    return {
        "workload": ("BASELINE.json configs[2]: all-vs-all core/accessory with --knn 50 sparse output (self_dists_knn), "
                     f"n={n} synthetic mutated-family sketches, 1024 bins (W=16), k=15..29 step 2 (K=8); the same job at "
                     "every N (strong scaling)"),
        "n_samples": n, "knn": KNN, "pairs": n * (n - 1) // 2, "pairs_counted": "unique (unordered) pairs; the reference "
        "evaluates each twice", "sketch_bins": 64 * W, "kmers": KMERS,
        "sharding": "upper-triangle row blocks with equal pair counts per rank, database replicated in every HBM; "
                    "partial kNN lists merged by the rows' owners out of peer memory (NVLink), final rows gathered on rank 0",
        "l2": "the 1.43 GB database and the 2 GB key matrices of a step exceed the 126 MB L2; a 256 MB buffer is "
              "overwritten between timed steps as well",
    }


# ---------------------------------------------------------------------------------------------------------------
def make_host_db(n, seed, device):
    """The synthetic database as a numpy u64 array in the .skd layout; generated on the GPU when there is one."""
    import numpy as np
    import torch
    from sketchlib.rust_b200 import synth
    out = np.empty(n * PER_SAMPLE, dtype=np.uint64)
    for s0 in range(0, n, 50_000):
        s1 = min(n, s0 + 50_000)
        t = synth.make_bins(s1 - s0, W, KMERS, seed=seed + s0, device=device)
        out[s0 * PER_SAMPLE:s1 * PER_SAMPLE] = synth.as_u64_numpy(t)
        del t
    if device != "cpu":
        torch.cuda.empty_cache()
    return out


def oracle_knn_rate(db, rows_per_call, threads, seconds, calls_min=1):
    """Times oracle.self_dists_knn on consecutive row blocks for about `seconds`; returns (ordered pairs/s, rows, s)."""
    import oracle
    n = db.n
    r, rows_done, t0 = 0, 0, time.perf_counter()
    calls = 0
    while True:
        r1 = min(n, r + rows_per_call)
        oracle.self_dists_knn(db, KNN, threads=threads, rows=(r, r1))
        rows_done += r1 - r
        calls += 1
        r = r1 % n
        dt = time.perf_counter() - t0
        if dt >= seconds and calls >= calls_min:
            return rows_done * (n - 1) / dt, rows_done, dt


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm for the SAME job (self_dists_knn on the 100 000-genome
    database, knn 50, core/accessory) on this box's host cores.  The Rust crate cannot be built in this image (no
    cargo/rustc), so this is the C restatement (oracle/sketchdist_oracle.c: same loop nest, heap and early break),
    all host threads, each step a bounded block of rows.  Like the reference it evaluates every ordered pair, so its
    unique-pair rate is half its pair-evaluation rate."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    import oracle
    n = N_KNN
    threads = os.cpu_count() or 1
    dev = "cuda:0" if torch.cuda.is_available() else "cpu"
    bins = make_host_db(n, 3, dev)   # same seed as the GPU arm's database
    db = oracle.Db(bins, n, KMERS, W)
    t = time.perf_counter()
    oracle.self_dists_knn(db, KNN, threads=threads, rows=(0, 4 * threads))
    per_row = (time.perf_counter() - t) / (4 * threads)
    # bounded sample per step: the whole --steps/--warmup run should take about 90 s of wall time
    step_seconds = min(args.ref_seconds, max(0.5, 90.0 / max(1, args.steps + args.warmup)))
    rows = max(threads, int(step_seconds / per_row) // threads * threads)
    rows = min(rows, n)
    for w in range(args.warmup):
        oracle.self_dists_knn(db, KNN, threads=threads, rows=(0, rows))
    t0 = time.perf_counter()
    for s in range(args.steps):
        r0 = (s * rows) % (n - rows + 1)
        oracle.self_dists_knn(db, KNN, threads=threads, rows=(r0, r0 + rows))
    dt = time.perf_counter() - t0
    ordered = rows * (n - 1) * args.steps / dt
    value = ordered / 2.0
    sample = (f"{rows} consecutive rows of the n={n} job per step (each row scans all {n - 1} other samples and keeps a "
              f"50-entry heap), {threads} threads over rows; {ordered:.3e} pair evaluations/s = {value:.3e} unique pairs/s "
              f"because the reference evaluates every unordered pair twice")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64+f64", "data": "synthetic",
        "config": knn_workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def parity_sample(res_idx, res_d0, res_d1, host_bins, n_rows_sample, threads, seed=1):
    """Random rows of the GPU result against oracle.self_dists_knn on the same database (test infrastructure used as
    the checker, after the timed region).  A set mismatch counts only outside the tolerance band around the oracle's
    k-th key (neighbours there are exact or near ties and may legitimately swap)."""
    import numpy as np
    import oracle
    db = oracle.Db(host_bins, N_KNN, KMERS, W)
    rng = np.random.default_rng(seed)
    rows = sorted(rng.choice(N_KNN, size=n_rows_sample, replace=False).tolist())
    max_diff, mism, order_bad, checked = 0.0, 0, 0, 0
    tol = 1e-6
    for r in rows:
        oidx, od0, od1 = oracle.self_dists_knn(db, KNN, threads=1, rows=(r, r + 1))
        oidx, od0, od1 = oidx[0], od0[0], od1[0]
        idx, d0, d1 = res_idx[r], res_d0[r], res_d1[r]
        kth = float(od0[-1])
        ours, theirs = set(int(x) for x in idx), set(int(x) for x in oidx)
        mism += len(set(int(j) for j, d in zip(oidx, od0) if d < kth - 2 * tol) - ours)
        mism += len(set(int(j) for j, d in zip(idx, d0) if d < kth - 2 * tol) - theirs)
        pos = {int(j): t for t, j in enumerate(oidx)}
        for t, j in enumerate(idx):
            if int(j) in pos:
                u = pos[int(j)]
                max_diff = max(max_diff, abs(float(d0[t]) - float(od0[u])), abs(float(d1[t]) - float(od1[u])))
                checked += 1
        max_diff = max(max_diff, float(np.abs(d0.astype(np.float64) - od0).max()))   # sorted keys, slot by slot
        order_bad += int((np.diff(d0) < 0).sum())
    return {"rows": len(rows), "entries_compared": checked, "max_abs_diff": max_diff,
            "set_mismatches_outside_ties": mism, "rows_not_ascending": order_bad, "tolerance": tol,
            "checker": "oracle.self_dists_knn (CPU restatement of src/distances/mod.rs:133-224) on the same database"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-seconds", type=float, default=10.0, help="CPU seconds per reference step / baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stream-roofline", action="store_true", help="skip the HBM-bound few-queries leg (N=1 only)")
    ap.add_argument("--no-dense-leg", action="store_true", help="skip the dense_10k secondary leg")
    ap.add_argument("--parity-rows", type=int, default=32)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist
    all_cpus = os.sched_getaffinity(0)           # restored before any CPU-oracle work (it uses every host core)
    import sketchlib.rust_b200 as sk

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gather_floats(x):
        if world == 1:
            return [x]
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    n = N_KNN
    total_pairs = n * (n - 1) // 2
    ctx = sk.Context(local)
    gpu_cpus = os.sched_getaffinity(0)           # where skd_ctx_create left this thread (SKD_BIND_CPU)
    n_dense = dense_n_for_gpus(world)

    # ---- the peer group: windows sized for the kNN rows, the loader slices of both databases
    win = max(sk.PeerGroup.window_bytes_knn(n, KNN, world), sk.PeerGroup.window_bytes_db(n, W, len(KMERS), world),
              sk.PeerGroup.window_bytes_db(n_dense, W, len(KMERS), world))
    grp = sk.PeerGroup(ctx, rank, world, win)
    if world > 1:
        handles = [None] * world
        dist.all_gather_object(handles, grp.export())
        grp.connect(handles)

    # ---- synthetic database (same seed on every rank); rank 0 keeps the whole host copy for the oracle checks, every
    # rank keeps its own sample slice in pinned host memory for the e2e leg
    host_bins = make_host_db(n, 3, dev)
    s0, s1 = grp.my_sample_slice(n)
    pin_slice = ctx.pinned(((s1 - s0) * PER_SAMPLE,), np.uint64)
    pin_slice.array[:] = host_bins[s0 * PER_SAMPLE:s1 * PER_SAMPLE]
    if rank != 0:
        del host_bins
        host_bins = None
    g = grp.upload(pin_slice.array, n, KMERS, W)
    dt_ca = sk.set_k(g)
    r0, r1 = sk.triangle_row_partition(n, rank, world)
    my_pairs = int(ctx.lib.skd_self_pairs_in_rows(n, r0, r1))
    is_root = rank == 0
    d_idx = ctx.device_alloc(n * KNN * 8) if is_root else None
    d_d0 = ctx.device_alloc(n * KNN * 4) if is_root else None
    d_d1 = ctx.device_alloc(n * KNN * 4) if is_root else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    peaks = int_peaks() if rank == 0 else None

    def pairs_of_last_call():
        a, b = grp.last_row_block if world > 1 else (r0, r1)   # the split adapts to the ranks' measured times
        return int(ctx.lib.skd_self_pairs_in_rows(n, a, b))

    def step_resident():
        grp.self_dists_knn(g, KNN, dt_ca, root=0, out=(d_idx, d_d0, d_d1), out_loc=sk.SKD_DEVICE)
        return grp.last_op_ms, ctx.last_tile_kernel_ms, ctx.last_kernel_ms, grp.last_exchange_ms, pairs_of_last_call()

    # ---- value: database resident in HBM, final rows left in rank 0's HBM
    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = ctx.launch_count
    barrier()
    t_wall0 = time.time()
    recs = []
    for _ in range(args.steps):
        flush.fill_(1)                 # evict L2 between timed steps (untimed)
        barrier()
        recs.append(step_resident())
    barrier()
    t_wall1 = time.time()
    launches = ctx.launch_count - launches0
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    op_ms = sum(r[0] for r in recs) / len(recs)
    tile_ms_ee = sum(r[1] for r in recs) / len(recs)
    kern_ms_ee = sum(r[2] for r in recs) / len(recs)
    xchg_ms = sum(r[3] for r in recs) / len(recs)
    ms_per_step = max_over_ranks(op_ms)
    value = total_pairs / (ms_per_step * 1e-3)
    my_pairs = sum(r[4] for r in recs) / len(recs)   # pairs this rank evaluated per step (adaptive split)
    xchg_all = gather_floats(xchg_ms)
    # how much of the match counting the early exit left (this rank), measured on one more step
    ctx.work_counters(reset=True)
    pairs_counted_step = step_resident()[4]
    steps_done, steps_total, pairs_deferred = ctx.work_counters(reset=True)

    # ---- roofline leg: the same job with the kernel's early exit OFF (every k-mer length of every pair is counted --
    # the 3584 LOP3 per pair the roofline is quoted on); not part of `value`
    ctx.set_early_exit(False)
    step_resident()
    barrier()
    recs_full = []
    for _ in range(max(2, args.steps // 2)):
        flush.fill_(1)
        barrier()
        recs_full.append(step_resident())
    barrier()
    ctx.set_early_exit(True)
    op_ms_full = max_over_ranks(sum(r[0] for r in recs_full) / len(recs_full))
    tile_ms = sum(r[1] for r in recs_full) / len(recs_full)
    kern_ms = sum(r[2] for r in recs_full) / len(recs_full)
    tile_ms_all = gather_floats(tile_ms)
    pairs_all = gather_floats(sum(r[4] for r in recs_full) / len(recs_full))   # per rank, in the roofline leg

    # ---- e2e: pinned host bit planes (1/N per rank) -> HBM -> peers -> distances -> exchange -> host rows on rank 0
    h_idx = ctx.pinned((n, KNN), np.uint64) if is_root else None
    h_d0 = ctx.pinned((n, KNN), np.float32) if is_root else None
    h_d1 = ctx.pinned((n, KNN), np.float32) if is_root else None
    out_host = (h_idx.array, h_d0.array, h_d1.array) if is_root else None

    def step_e2e():
        gg = grp.upload(pin_slice.array, n, KMERS, W)
        grp.self_dists_knn(gg, KNN, dt_ca, root=0, out=out_host, out_loc=sk.SKD_HOST)
        gg.free()

    for _ in range(args.warmup):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / args.steps)
    e2e_value = total_pairs / e2e_s
    h2d_all = gather_floats(float((s1 - s0) * PER_SAMPLE * 8))

    parity = None
    if is_root and args.parity_rows > 0:
        os.sched_setaffinity(0, all_cpus)  # the oracle's threads inherit this thread's affinity
        parity = parity_sample(h_idx.array, h_d0.array, h_d1.array, host_bins, args.parity_rows, os.cpu_count() or 1)
        # the resident leg's rows (left in HBM) must be the same rows
        chk = np.empty((n, KNN), dtype=np.float32)
        ctx.memcpy(chk, sk.SKD_HOST, d_d0, sk.SKD_DEVICE, chk.nbytes)
        parity["resident_equals_e2e"] = bool(np.array_equal(chk, h_d0.array))
        parity["checksum_core"] = float(h_d0.array.astype(np.float64).sum())
        parity["checksum_accessory"] = float(h_d1.array.astype(np.float64).sum())
        os.sched_setaffinity(0, gpu_cpus)
    g.free()

    dense = None
    if not args.no_dense_leg:
        dense = dense_leg(args, sk, ctx, grp, rank, world, dev, barrier, max_over_ranks, flush, n_dense)

    if rank == 0:
        if peaks and peaks.get("lop3_per_s"):
            lop3_peak, peak_src = peaks["lop3_per_s"] / 1e12, "measured live by tools/int_peak (all-SM LOP3 chain microbenchmark)"
        else:
            smhz = (clocks or {}).get("sm_max_mhz", 1965.0)
            lop3_peak, peak_src = 148 * 64 * smhz * 1e6 / 1e12, "nominal 64 LOP3/clk/SM x 148 SMs x max SM clock (int_peak unavailable)"
        per_gpu = [p * LOP3_PER_PAIR / (t * 1e-3) / 1e12 for p, t in zip(pairs_all, tile_ms_all)]
        achieved = per_gpu[0]
        ncu = None
        p = os.path.join(ROOT, "profiles", "r2_ncu_dense_kernel_knn100k.json")
        if os.path.exists(p):
            try:
                ncu = json.load(open(p))
            except Exception:
                ncu = None
        roofline = {
            "bound": "int",
            "bound_note": "integer ALU (LOP3) issue: the kernel is bitwise, neither HBM- nor tensor-core-bound",
            "kernel": "skd::dense_kernel<uint16_t, false> (rect mode: kNN keys of upper-triangle row groups, persistent 64x128 tiles)",
            "achieved": achieved, "peak": lop3_peak, "unit": "TLOP3/s", "frac": achieved / lop3_peak,
            "per_gpu_frac": [a / lop3_peak for a in per_gpu],
            "tile_kernel_ms_per_step": tile_ms, "all_kernels_ms_per_step": kern_ms,
            "tile_kernel_share_of_step": tile_ms / kern_ms,
            "measured_on": "the same job with the kernel's early exit switched off (skd_ctx_set_early_exit(0)): every "
                           "k-mer length of every pair is counted, 3584 LOP3 per pair; ms_per_step of that run: "
                           f"{op_ms_full:.2f}",
            "early_exit": {
                "what": "default mode, the one `value` and `e2e` are measured in: like the reference's loop (break at the "
                        "first k-mer length without a match, jaccard.rs:89-91; n < 3 -> (1, 1)) a warp stops counting its "
                        "16 x 64 pair block after the third k-mer length when at most 32 of its pairs are still open and "
                        "hands those to defer_kernel; results are identical (parity_sample, resident_equals_e2e)",
                "match_count_steps_executed_frac": steps_done / max(steps_total, 1),
                "pairs_finished_by_defer_kernel_frac": pairs_deferred / max(pairs_counted_step, 1),
                "tile_and_defer_kernel_ms_per_step": tile_ms_ee, "all_kernels_ms_per_step": kern_ms_ee,
                "executed_tlop3_per_s": my_pairs * LOP3_PER_PAIR * (steps_done / max(steps_total, 1)) / (tile_ms_ee * 1e-3) / 1e12,
            },
            "peak_source": peak_src,
            "algorithmic_ops_per_pair": {"lop3": LOP3_PER_PAIR, "popc": POPC_PER_PAIR},
            "popc_frac": (pairs_all[0] * POPC_PER_PAIR / (tile_ms * 1e-3)) / peaks["popc_per_s"] if peaks and peaks.get("popc_per_s") else None,
            "mix_frac": achieved * 1e12 / peaks["mix_lop3_per_s"] if peaks and peaks.get("mix_lop3_per_s") else None,
            "peak_clock_mhz": (peaks or {}).get("sm_mhz"),
            "traffic": (ncu or {}).get("dram_bytes_per_launch"),
            "traffic_note": (ncu or {}).get("traffic_note"),
            "int_peaks": peaks,
        }
        config = knn_workload_config(world)
        if world > 1:
            config["sharding"] += ("; the first call splits the triangle into equal pair counts, later calls re-split it "
                                   "from the ranks' measured times (skd_comm_last_row_block)")
        line = {
            "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u32 bit planes (LOP3/POPC) + f64 epilogue -> f32", "data": "synthetic",
            "config": config,
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": int(sum(h2d_all)),
                    "d2h_bytes_per_step": n * KNN * 16, "ms_per_step": e2e_s * 1e3,
                    "h2d_bytes_per_rank": [int(x) for x in h2d_all],
                    "note": "every rank uploads its 1/N sample slice from pinned host memory, re-tiles it and pulls the other "
                            "slices from its peers' HBM (skd_comm_db_upload); rank 0 receives the final rows in host memory"},
            "exchange": {"ms_per_step_per_rank": xchg_all,
                         "what": "entry of the post-kernel barrier -> merged rows gathered on rank 0 (barriers + peer merge "
                                 "kernel + row gather), CUDA events per rank; includes waiting for the slowest rank",
                         "transport": "library kernels on peer memory (CUDA IPC windows over NVLink); no NCCL on the data path"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "parity_sample": parity,
            "wall_ms_per_step_incl_flush": (t_wall1 - t_wall0) / args.steps * 1e3,
        }
        if dense is not None:
            dense["roofline_frac_lop3_rank0"] = dense["achieved_tlop3_rank0"] / lop3_peak
            line["dense_10k"] = dense
        if world == 1 and not args.no_stream_roofline:
            from sketchlib.rust_b200 import synth
            line["roofline_hbm"] = stream_roofline(ctx, sk, synth, torch)
        if not args.no_cpu_baseline and world == 1:
            import oracle
            os.sched_setaffinity(0, all_cpus)
            threads = os.cpu_count() or 1
            db = oracle.Db(host_bins, n, KMERS, W)
            ordered, rows_done, secs = oracle_knn_rate(db, 8 * threads, threads, args.ref_seconds)
            line["cpu_baseline"] = {
                "value": ordered / 2.0, "unit": "pairs/s", "cores": threads, "kind": "port",
                "sample": f"{rows_done} rows of the same n={n} kNN-50 job ({secs:.1f} s of wall time), "
                          f"oracle/sketchdist_oracle.c (restated reference: the Rust crate cannot be built here), {threads} "
                          f"threads over rows; {ordered:.3e} ordered-pair evaluations/s, halved because the reference "
                          f"evaluates each unordered pair twice; like the reference it stops at the first k-mer length "
                          f"without matches, so unrelated pairs cost it 1/8 of the bit work",
            }
        print(json.dumps(line))
    barrier()
    for p in (d_idx, d_d0, d_d1):
        if p:
            ctx.device_free(p)
    grp.close()
    if world > 1:
        dist.destroy_process_group()


def dense_leg(args, sk, ctx, grp, rank, world, dev, barrier, max_over_ranks, flush, n):
    """BASELINE.json configs[1] (round 1's headline): dense all-vs-all core/accessory, n = 10 000 (x sqrt(N): every
    rank's triangle row block keeps ~5e7 pairs), results as f32 pairs: value = resident, e2e = host planes in (1/N per
    rank + peers), 400 MB of results per rank to pinned host memory."""
    import numpy as np
    import torch
    from sketchlib.rust_b200 import synth
    steps = max(3, min(args.steps, 10))
    total_pairs = n * (n - 1) // 2
    r0, r1 = sk.triangle_row_partition(n, rank, world)
    my_pairs = int(ctx.lib.skd_self_pairs_in_rows(n, r0, r1))
    dev_bins = synth.make_bins(n, W, KMERS, seed=world, device=dev)
    torch.cuda.synchronize()
    s0, s1 = grp.my_sample_slice(n)
    pin_in = ctx.pinned(((s1 - s0) * PER_SAMPLE,), np.uint64)
    ctx.memcpy(pin_in.array, sk.SKD_HOST, dev_bins.data_ptr() + s0 * PER_SAMPLE * 8, sk.SKD_DEVICE, (s1 - s0) * PER_SAMPLE * 8)
    g = sk.SketchDb.upload(ctx, dev_bins.data_ptr(), n, KMERS, W, bins_loc=sk.SKD_DEVICE)
    del dev_bins
    dt_ca = sk.set_k(g)
    out_bytes = my_pairs * 2 * 4
    d_out = ctx.device_alloc(max(out_bytes, 8))

    def step_resident():
        ctx._check(ctx.lib.skd_self_dists_all(ctx.handle, g.handle, dt_ca.c, None, 0.64, r0, r1, d_out, sk.SKD_DEVICE))
        return ctx.last_kernel_ms

    for _ in range(3):
        step_resident()
    ms = []
    barrier()
    for _ in range(steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        ms.append(step_resident())
    barrier()
    my_ms = sum(ms) / len(ms)
    ms_per_step = max_over_ranks(my_ms)
    # the kernel-quality figure: the same launch with the early exit off (3584 LOP3 for every pair)
    ctx.set_early_exit(False)
    step_resident()
    ms_full = []
    for _ in range(steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        ms_full.append(step_resident())
    ctx.set_early_exit(True)
    my_ms_full = sum(ms_full) / len(ms_full)
    barrier()
    pin_out = ctx.pinned((max(my_pairs * 2, 2),), np.float32)

    def step_e2e():
        if world == 1:
            gg = sk.SketchDb.upload(ctx, pin_in.array, n, KMERS, W, async_upload=True)  # overlaps PCIe with the tiles
        else:
            gg = grp.upload(pin_in.array, n, KMERS, W)
        ctx._check(ctx.lib.skd_self_dists_all(ctx.handle, gg.handle, dt_ca.c, None, 0.64, r0, r1,
                                              pin_out.array.ctypes.data, sk.SKD_HOST))
        gg.free()

    for _ in range(3):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step_e2e()
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / steps)
    checksum = float(pin_out.array[: my_pairs * 2].astype(np.float64).sum())
    g.free()
    ctx.device_free(d_out)
    pin_out.free()
    pin_in.free()
    return {
        "workload": f"BASELINE.json configs[1]: dense all-vs-all core/accessory (self_dists_all), n={n}"
                    + ("" if world == 1 else f" = 10000*sqrt({world}) (weak scaling: ~5e7 pairs per rank)"),
        "pairs": total_pairs, "steps": steps, "value": total_pairs / (ms_per_step * 1e-3), "unit": "pairs/s",
        "ms_per_step": ms_per_step,
        "roofline_frac_lop3_rank0": None,   # filled by the caller's peak if wanted: achieved below
        "achieved_tlop3_rank0": my_pairs * LOP3_PER_PAIR / (my_ms_full * 1e-3) / 1e12,
        "ms_per_step_early_exit_off_rank0": my_ms_full,
        "e2e": {"value": total_pairs / e2e_s, "unit": "pairs/s", "ms_per_step": e2e_s * 1e3,
                "h2d_bytes_per_step_per_rank": (s1 - s0) * PER_SAMPLE * 8, "d2h_bytes_per_step_per_rank": out_bytes},
        "result_checksum_rank0": checksum,
    }


def stream_roofline(ctx, sk, synth, torch, n_ref=200000, reps=5):
    """Second roofline of the path: ONE query against n_ref references is HBM-bound.  Dense
    cross core/accessory through the C ABI with the result left in HBM; algorithmic bytes = the reference planes read
    once (K*W*14*8 B per sample).  Timed by the library's CUDA events on its own stream; the 2.9 GB input exceeds L2."""
    per = PER_SAMPLE
    bins = torch.empty(n_ref * per, dtype=torch.int64, device="cuda")
    for s0 in range(0, n_ref, 50000):
        s1 = min(n_ref, s0 + 50000)
        bins[s0 * per:s1 * per] = synth.make_bins(s1 - s0, W, KMERS, seed=1000 + s0, device="cuda")
    refs = sk.SketchDb.upload(ctx, bins.data_ptr(), n_ref, KMERS, W, bins_loc=sk.SKD_DEVICE)
    del bins
    q = synth.make_bins(1, W, KMERS, seed=7, device="cuda")
    qry = sk.SketchDb.upload(ctx, q.data_ptr(), 1, KMERS, W, bins_loc=sk.SKD_DEVICE)
    out = ctx.device_alloc(n_ref * 8)
    dt = sk.set_k(refs)
    ms = []
    for i in range(reps + 2):
        ctx._check(ctx.lib.skd_cross_dists_all(ctx.handle, refs.handle, qry.handle, dt.c, None, None, 0.64, 0, n_ref, out,
                                               sk.SKD_DEVICE))
        if i >= 2:
            ms.append(ctx.last_kernel_ms)
    avg = sum(ms) / len(ms)
    algorithmic = n_ref * per * 8
    traffic = None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "r1_ncu_stream_kernel.json")))
        traffic = float(prof["launches"][0]["dram__bytes_read.sum"].split()[0]) * 1e9 * n_ref / 200000
    except Exception:
        pass
    ctx.device_free(out)
    refs.free()
    qry.free()
    peak = measured_hbm()
    return {"bound": "hbm", "kernel": "skd::stream_kernel<1>",
            "workload": f"dense cross core/accessory, {n_ref} references x 1 query, 1024 bins x 8 k, results in HBM",
            "achieved": algorithmic / (avg * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
            "frac": algorithmic / (avg * 1e-3) / 1e9 / peak, "ms_per_launch": avg,
            "algorithmic_bytes_per_launch": algorithmic, "traffic": traffic,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (copy: read+write bytes); this kernel only reads, so frac can exceed 1"}


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Benchmark of the pairwise sketch-distance path (BASELINE.json metric: sketch pairs/sec, core/accessory,
1024 bins, 8 k-mer lengths).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...        (one rank per GPU)

A "step" is one pass of the hot path over the whole synthetic database: dense all-vs-all core/accessory
distances (the reference's self_dists_all) -- at N=1 on BASELINE.json configs[1] (10 000 genomes, 1024-bin
sketches, k = 15..29 step 2 => 49 995 000 pairs, 143 MB of bit planes, 400 MB of f32 results).  At N > 1 the
database grows to n = 10 000 * sqrt(N) samples so that each rank's equal-pair-count block of upper-triangle
rows holds the same ~5e7 pairs ("weak" scaling); the database is replicated, every rank writes its own
slice of the condensed matrix, and there is no data-path collective.

Prints ONE JSON line (rank 0).  `value` = pairs/s with the database resident in HBM and results left in
HBM, timed with CUDA events on the library's compute stream (max over ranks); `e2e` = the same metric
through the public host API with pinned HOST buffers: upload of the .skd-layout bit planes, re-tiling,
distances, and the device->host copy of the result, all inside the timed region.
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
N_BASE = 10000
LOP3_PER_PAIR = len(KMERS) * W * 28   # K * W * (28 LOP3 + 2 POPC) per pair
POPC_PER_PAIR = len(KMERS) * W * 2
METRIC = "sketch pairs/sec (core/accessory, 1024 bins, 8 k)"


def n_for_gpus(n_gpus):
    n = int(round(N_BASE * n_gpus ** 0.5 / 64.0)) * 64 if n_gpus > 1 else N_BASE
    return n


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
        sm, mx, reasons = [], [], set()
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9 or not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def int_peaks():
    """LOP3/POPC issue peaks measured live on this GPU by tools/int_peak (built by __graft_entry__.build)."""
    exe = os.path.join(ROOT, "tools", "int_peak")
    if not os.path.exists(exe):
        src = exe + ".cu"
        try:
            subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-o", exe, src])
        except Exception:
            return None
    try:
        out = subprocess.run([exe, "0.5"], capture_output=True, text=True, timeout=120).stdout
        return json.loads(out)
    except Exception:
        return None


def cpu_reference_rate(db_bins_u64, n, seconds_target, threads):
    """Times the CPU oracle (restated reference, oracle/sketchdist_oracle.c) on a bounded sample of the
    workload: passes over the condensed pairs of the same database (chunked and threaded like the reference's
    rayon par_chunks_mut(1000)) repeated until about `seconds_target` seconds of wall time have been spent."""
    import oracle
    db = oracle.Db(db_bins_u64, n, KMERS, W)
    total = n * (n - 1) // 2
    probe = min(total, 400_000)
    t = time.perf_counter()
    oracle.self_dists_all(db, threads=threads, first=0, count=probe)
    rate = probe / (time.perf_counter() - t)
    count = int(min(total, max(probe, rate * seconds_target)) // 1000 * 1000) or probe
    passes = max(1, int(round(rate * seconds_target / count)))
    t = time.perf_counter()
    for _ in range(passes):
        oracle.self_dists_all(db, threads=threads, first=0, count=count)
    dt = time.perf_counter() - t
    return count * passes / dt, count, passes, dt


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm for the path on this box's host cores.  The Rust
    crate cannot be built in this image (no cargo/rustc), so this is the oracle port, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from sketchlib.rust_b200 import synth
    n = n_for_gpus(args.gpus)
    threads = os.cpu_count() or 1
    bins = synth.as_u64_numpy(synth.make_bins(n, W, KMERS, seed=args.gpus))
    import oracle
    db = oracle.Db(bins, n, KMERS, W)
    total = n * (n - 1) // 2
    probe = min(total, 400_000)
    t = time.perf_counter()
    oracle.self_dists_all(db, threads=threads, first=0, count=probe)
    rate = probe / (time.perf_counter() - t)
    # bounded sample per step: the whole --steps/--warmup run should take about 90 s of wall time
    step_seconds = min(args.ref_seconds, max(0.5, 90.0 / max(1, args.steps + args.warmup)))
    count = int(min(total, max(100_000, rate * step_seconds)) // 1000 * 1000)
    for _ in range(args.warmup):
        oracle.self_dists_all(db, threads=threads, first=0, count=count)
    t0 = time.perf_counter()
    for s in range(args.steps):
        first = (s * count) % max(1000, (total - count) // 1000 * 1000)
        oracle.self_dists_all(db, threads=threads, first=first, count=count)
    dt = time.perf_counter() - t0
    value = count * args.steps / dt
    sample = (f"{count} consecutive condensed pairs per step of the n={n} database (of {total}), rayon-style "
              f"chunks of 1000 pairs over {threads} threads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64+f64", "data": "synthetic",
        "config": workload_config(n, args.gpus),
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(n, n_gpus):
    return {
        "workload": ("BASELINE.json configs[1]: dense all-vs-all core/accessory (self_dists_all), "
                     f"n={n} synthetic mutated-family sketches, 1024 bins (W=16), k=15..29 step 2 (K=8)"
                     + ("" if n_gpus == 1 else f"; n = 10000*sqrt({n_gpus}) so each rank's triangle row block keeps ~5e7 pairs")),
        "n_samples": n, "pairs": n * (n - 1) // 2, "sketch_bins": 64 * W, "kmers": KMERS,
        "sharding": "upper-triangle row blocks with equal pair counts, database replicated, no collective",
        "l2": "inputs+outputs exceed L2 (143 MB bit planes + 400 MB results per step vs 126 MB) and a 256 MB "
              "buffer is overwritten between timed steps",
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-seconds", type=float, default=10.0, help="CPU seconds per reference step / baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stream-roofline", action="store_true", help="skip the HBM-bound few-queries leg (N=1 only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import sketchlib.rust_b200 as sk
    from sketchlib.rust_b200 import synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
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

    n = n_for_gpus(world)
    total_pairs = n * (n - 1) // 2
    r0, r1 = sk.triangle_row_partition(n, rank, world)
    ctx = sk.Context(local)
    my_pairs = int(ctx.lib.skd_self_pairs_in_rows(n, r0, r1))

    # ---- synthetic database, generated on the GPU (same seed on every rank => replicated database)
    dev_bins = synth.make_bins(n, W, KMERS, seed=world, device=f"cuda:{local}")
    torch.cuda.synchronize()
    db_bytes = dev_bins.numel() * 8
    pin_in = ctx.pinned((dev_bins.numel(),), np.uint64)     # host copy in pinned memory for the e2e leg
    ctx.memcpy(pin_in.array, sk.SKD_HOST, dev_bins.data_ptr(), sk.SKD_DEVICE, db_bytes)
    g = sk.SketchDb.upload(ctx, dev_bins.data_ptr(), n, KMERS, W, bins_loc=sk.SKD_DEVICE)
    del dev_bins
    dt_ca = sk.set_k(g)
    out_bytes = my_pairs * 2 * 4
    d_out = ctx.device_alloc(max(out_bytes, 8))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    peaks = int_peaks() if rank == 0 else None

    def step_resident():
        ctx._check(ctx.lib.skd_self_dists_all(ctx.handle, g.handle, dt_ca.c, None, 0.64, r0, r1, d_out, sk.SKD_DEVICE))
        return ctx.last_kernel_ms

    # ---- value: database resident in HBM, results left in HBM, CUDA-event time of the kernels
    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = ctx.launch_count
    barrier()
    t_wall0 = time.time()
    dev_ms = []
    for _ in range(args.steps):
        flush.fill_(1)                 # evict L2 between timed steps (untimed)
        torch.cuda.synchronize()
        dev_ms.append(step_resident())
    barrier()
    t_wall1 = time.time()
    launches = ctx.launch_count - launches0
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    ms_per_step = max_over_ranks(sum(dev_ms) / len(dev_ms))
    value = total_pairs / (ms_per_step * 1e-3)

    # ---- e2e: pinned host bit planes -> HBM -> re-tile -> distances -> pinned host results
    pin_out = ctx.pinned((max(my_pairs * 2, 2),), np.float32)

    # With N > 1 every rank needs the whole database but copies only its 1/N slice of the samples over PCIe;
    # the slices are exchanged on-device with one NCCL all-gather over NVLink before the re-tiling kernel.
    per_sample = len(KMERS) * W * 14
    s0, s1 = sk.row_partition(n, rank, world)
    max_slice = (n + world - 1) // world
    if world > 1:
        raw_slice = torch.zeros(max_slice * per_sample, dtype=torch.int64, device="cuda")
        raw_all = torch.empty(world * max_slice * per_sample, dtype=torch.int64, device="cuda")
        raw_db = torch.empty(n * per_sample, dtype=torch.int64, device="cuda")
        bounds = [sk.row_partition(n, r, world) for r in range(world)]
    h2d_bytes = (s1 - s0) * per_sample * 8

    def step_e2e():
        if world == 1:
            gg = sk.SketchDb.upload(ctx, pin_in.array, n, KMERS, W, async_upload=True)  # overlaps PCIe with the tiles
        else:
            ctx.memcpy(raw_slice.data_ptr(), sk.SKD_DEVICE, pin_in.array[s0 * per_sample:s1 * per_sample], sk.SKD_HOST,
                       h2d_bytes)
            dist.all_gather_into_tensor(raw_all, raw_slice)
            for r, (a, b) in enumerate(bounds):  # slices are padded to equal length for the collective
                raw_db[a * per_sample:b * per_sample] = raw_all[r * max_slice * per_sample:
                                                                r * max_slice * per_sample + (b - a) * per_sample]
            torch.cuda.synchronize()
            gg = sk.SketchDb.upload(ctx, raw_db.data_ptr(), n, KMERS, W, bins_loc=sk.SKD_DEVICE)
        ctx._check(ctx.lib.skd_self_dists_all(ctx.handle, gg.handle, dt_ca.c, None, 0.64, r0, r1,
                                              pin_out.array.ctypes.data, sk.SKD_HOST))
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
    checksum = float(pin_out.array[: my_pairs * 2].astype(np.float64).sum())

    if rank == 0:
        lop3_peak, peak_src = None, None
        if peaks and peaks.get("lop3_per_s"):
            lop3_peak, peak_src = peaks["lop3_per_s"] / 1e12, "measured live by tools/int_peak (all-SM LOP3 chain microbenchmark)"
        else:
            smhz = (clocks or {}).get("sm_max_mhz", 1965.0)
            lop3_peak, peak_src = 148 * 64 * smhz * 1e6 / 1e12, "nominal 64 LOP3/clk/SM x 148 SMs x max SM clock (int_peak unavailable)"
        # kernel-only rate of rank 0's slice (its own launch duration)
        my_ms = sum(dev_ms) / len(dev_ms)
        achieved = my_pairs * LOP3_PER_PAIR / (my_ms * 1e-3) / 1e12
        ncu = None
        p = os.path.join(ROOT, "profiles", "ncu_tile_kernel_summary.json")
        if os.path.exists(p):
            try:
                ncu = json.load(open(p))
            except Exception:
                ncu = None
        roofline = {
            "bound": "int",
            "bound_note": "integer ALU (LOP3) issue: the kernel is bitwise, neither HBM- nor tensor-core-bound",
            "kernel": "skd::tile_kernel<uint16_t>",
            "achieved": achieved, "peak": lop3_peak, "unit": "TLOP3/s", "frac": achieved / lop3_peak,
            "peak_source": peak_src,
            "algorithmic_ops_per_pair": {"lop3": LOP3_PER_PAIR, "popc": POPC_PER_PAIR},
            "popc_frac": (my_pairs * POPC_PER_PAIR / (my_ms * 1e-3)) / peaks["popc_per_s"] if peaks and peaks.get("popc_per_s") else None,
            "mix_frac": achieved * 1e12 / peaks["mix_lop3_per_s"] if peaks and peaks.get("mix_lop3_per_s") else None,
            "hbm": {"achieved_gbs": (g.hbm_bytes + out_bytes) / (my_ms * 1e-3) / 1e9, "peak_gbs": measured_hbm(),
                    "note": "algorithmic bytes = tiled database read once + results written once per launch"},
            "traffic": (ncu or {}).get("dram_bytes_per_launch"),
            "int_peaks": peaks,
        }
        line = {
            "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32 bit planes (LOP3/POPC) + f64 epilogue -> f32", "data": "synthetic",
            "config": workload_config(n, world),
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": out_bytes, "ms_per_step": e2e_s * 1e3,
                    "note": "per rank; with N > 1 each rank uploads 1/N of the samples and the database is completed "
                            "on-device by one NCCL all-gather over NVLink" if world > 1 else "single GPU"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
            "wall_ms_per_step_incl_flush": (t_wall1 - t_wall0) / args.steps * 1e3,
            "result_checksum_rank0": checksum,
        }
        if world == 1 and not args.no_stream_roofline:
            line["roofline_hbm"] = stream_roofline(ctx, sk, synth, torch)
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            rate, count, passes, secs = cpu_reference_rate(pin_in.array, n, args.ref_seconds, threads)
            line["cpu_baseline"] = {
                "value": rate, "unit": "pairs/s", "cores": threads, "kind": "port",
                "sample": f"{passes} pass(es) over the first {count} condensed pairs of the same database ({secs:.1f} s of "
                          f"wall time), oracle/sketchdist_oracle.c (restated reference: the Rust crate cannot be built "
                          f"here), {threads} threads over chunks of 1000 pairs; like the reference it stops at the first "
                          f"k-mer length without matches, so unrelated pairs cost it 1/8 of the bit work",
            }
        print(json.dumps(line))
    barrier()
    ctx.device_free(d_out)
    if world > 1:
        dist.destroy_process_group()


def stream_roofline(ctx, sk, synth, torch, n_ref=200000, reps=5):
    """Second roofline of the path: ONE query against n_ref references is HBM-bound.  Dense
    cross core/accessory through the C ABI with the result left in HBM; algorithmic bytes = the reference planes read
    once (K*W*14*8 B per sample).  Timed by the library's CUDA events on its own stream; the 2.9 GB input exceeds L2."""
    per = len(KMERS) * W * 14
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


def measured_hbm():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        return 6650.0


if __name__ == "__main__":
    main()

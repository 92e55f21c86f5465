"""Summarises one kernel of an .ncu-rep (ncu --set full --import-source on) as JSON for profiles/:
duration, pipe / issue utilisation, DRAM and L2 traffic, registers, opcode mix and warp-stall breakdown (from the
source page, weighted by executed instructions / samples).

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [--kernel dense_kernel] [--note "..."] > profiles/r2_ncu_x.json
"""
import argparse
import collections
import csv
import io
import json
import re
import subprocess

RAW_KEEP = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpc__cycles_elapsed.max", "sm__cycles_elapsed.avg.per_second",
]
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--kernel", default="")
    ap.add_argument("--note", default="")
    ap.add_argument("--command", default="")
    args = ap.parse_args()
    raw = ncu_csv(args.rep, "raw")
    hdr, units = raw[0], raw[1]
    rows = [r for r in raw[2:] if len(r) == len(hdr)]
    ki = hdr.index("Kernel Name")
    row = next((r for r in rows if args.kernel in r[ki]), rows[0])
    metrics = {}
    for h, u, v in zip(hdr, units, row):
        if h in RAW_KEEP:
            try:
                x = float(v.replace(",", ""))
            except ValueError:
                continue
            if u in UNIT_SCALE and "bytes" in h:
                x, u = x * UNIT_SCALE[u], "byte"
            metrics[h] = (x, u)
    flat = {f"{h} [{u}]" if u else h: x for h, (x, u) in metrics.items()}
    out = {"report": args.rep, "command": args.command, "note": args.note, "kernel": row[ki],
           "grid": row[hdr.index("Grid Size")] if "Grid Size" in hdr else None,
           "block": row[hdr.index("Block Size")] if "Block Size" in hdr else None, "metrics": flat}
    if "dram__bytes_read.sum" in metrics:
        out["dram_bytes_per_launch"] = metrics["dram__bytes_read.sum"][0] + metrics["dram__bytes_write.sum"][0]
    src = ncu_csv(args.rep, "source")
    # the source page holds one table per kernel: take the one whose title matches
    start = 0
    for i, r in enumerate(src):
        if r and r[0] == "Kernel Name" and args.kernel in (r[1] if len(r) > 1 else ""):
            start = i
            break
    sh = src[start + 1]
    ix = {h: i for i, h in enumerate(sh)}
    op, stall, tot = collections.Counter(), collections.Counter(), 0
    for r in src[start + 2:]:
        if r and r[0] == "Kernel Name":
            break
        if len(r) < len(sh):
            continue
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]].strip())
        o = m.group(2).split(".")[0] if m else "?"
        n = int(r[ix["Instructions Executed"]] or 0)
        op[o] += n
        tot += n
        for h in sh:
            if h.startswith("stall_") and "Not Issued" not in h:
                stall[h] += int(r[ix[h]] or 0)
    out["warp_instructions_executed"] = tot
    out["opcode_share_pct"] = {o: round(100.0 * n / tot, 2) for o, n in op.most_common(14)}
    st = sum(stall.values()) or 1
    out["warp_stall_samples_pct"] = {h: round(100.0 * n / st, 1) for h, n in stall.most_common(10)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

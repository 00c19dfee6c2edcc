"""Summarise an `ncu --set full` report for profiles/ (run HERE, no GPU needed):

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > gpurun_out/prof_raw.csv
    python tools/ncu_summarize.py gpurun_out/prof_raw.csv --title "r02 ..." --out profiles/r02_x_ncu_full_summary.md \
        [--traffic-json profiles/r02_gemm_traffic.json --traffic-kernel gemm_tc05]

One table row per captured launch with the counters the roofline discussion uses; the traffic JSON is what
bench.py reports as `roofline.traffic` (DRAM bytes per launch of the dominant kernel, with the git hash of
the build it was captured from).
"""
import argparse
import csv
import json
import re
import subprocess

COLS = [
    ("gpu__time_duration.sum", "us", 1e-3, "duration"),
    ("dram__bytes_read.sum", "MB", None, "DRAM read"),
    ("dram__bytes_write.sum", "MB", None, "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "%", 1, "DRAM %"),
    ("lts__t_sector_hit_rate.pct", "%", 1, "L2 hit"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "%", 1, "tensor pipe"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "%", 1, "XU pipe"),
    ("sm__issue_active.avg.pct_of_peak_sustained_active", "%", 1, "issue active"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "%", 1, "warps active"),
    ("launch__registers_per_thread", "", 1, "regs"),
    ("launch__grid_size", "", 1, "grid"),
    ("launch__block_size", "", 1, "block"),
    ("launch__occupancy_limit_shared_mem", "", 1, "occ lim smem"),
    ("launch__occupancy_limit_registers", "", 1, "occ lim regs"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "", 1, "smem bank conflicts"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "", 1, "stall long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "", 1, "stall short_sb"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "", 1, "stall mio"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "", 1, "stall barrier"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "", 1, "stall wait"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "", 1, "stall lg"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "", 1, "stall math"),
]


def num(s):
    try:
        return float(s.replace(",", ""))
    except (ValueError, AttributeError):
        return None


def to_bytes(v, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return v * mult.get(unit, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--title", default="ncu --set full summary")
    ap.add_argument("--out", required=True)
    ap.add_argument("--note", default="")
    ap.add_argument("--traffic-json")
    ap.add_argument("--traffic-kernel", default="gemm_tc05")
    args = ap.parse_args()
    rows = list(csv.reader(open(args.csv, newline="")))
    # ncu raw csv: header row, units row, then one row per launch
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units, data = rows[hdr_i], rows[hdr_i + 1], rows[hdr_i + 2:]
    idx = {name: i for i, name in enumerate(hdr)}
    kn = idx["Kernel Name"]
    lines = [f"# {args.title}", "", args.note, ""] if args.note else [f"# {args.title}", ""]
    present = [(m, u, s, label) for m, u, s, label in COLS if m in idx]
    lines.append("| # | kernel | " + " | ".join(label for *_, label in present) + " |")
    lines.append("|---|---|" + "---:|" * len(present))
    traffic = []
    for n, r in enumerate(data):
        if len(r) <= kn:
            continue
        name = re.sub(r"\(.*", "", r[kn])
        cells = []
        rd = wr = 0.0
        for m, u, s, label in present:
            v = num(r[idx[m]])
            unit = units[idx[m]]
            if v is None:
                cells.append("-")
                continue
            if m.startswith("dram__bytes"):
                b = to_bytes(v, unit)
                if "read" in m:
                    rd = b
                else:
                    wr = b
                cells.append(f"{b / 1e6:.2f}")
            elif m == "gpu__time_duration.sum":
                ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
                cells.append(f"{ns / 1e3:.1f}")
            else:
                cells.append(f"{v:.2f}" if abs(v) < 1000 else f"{v:.0f}")
        lines.append(f"| {n} | `{name}` | " + " | ".join(cells) + " |")
        if args.traffic_kernel in r[kn]:
            traffic.append(rd + wr)
    open(args.out, "w").write("\n".join(lines) + "\n")
    print("wrote", args.out, len(data), "launches")
    if args.traffic_json and traffic:
        git = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
        json.dump({"kernel": args.traffic_kernel, "source": f"{args.out} (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)",
                   "launches": len(traffic), "mean_dram_bytes_per_launch": int(sum(traffic) / len(traffic)),
                   "per_launch": [int(t) for t in traffic], "git": git,
                   "note": "captured from the build at this commit (working tree may carry later doc-only changes)"},
                  open(args.traffic_json, "w"), indent=2)
        print("wrote", args.traffic_json)


if __name__ == "__main__":
    main()

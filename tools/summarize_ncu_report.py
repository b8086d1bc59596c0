"""Key figures of one kernel launch out of an ``ncu --set full`` report (.ncu-rep), as the small JSON committed under profiles/.

    python tools/summarize_ncu_report.py gpurun_out/prof.ncu-rep [--launch 0] [--note "..."] > profiles/rNN_ncu_<kernel>_summary.json

Reads the report with ``ncu -i <rep> --page raw --csv`` (no GPU needed) and keeps: duration, clocks, instruction count, issue
utilisation, eligible warps, occupancy limits, pipe utilisation, DRAM bytes and the warp-stall breakdown per issued instruction.
"""
import argparse
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "launch__waves_per_multiprocessor", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
]
STALLS = ["barrier", "wait", "short_scoreboard", "long_scoreboard", "math_pipe_throttle", "mio_throttle", "not_selected", "dispatch_stall",
          "branch_resolving", "no_instruction", "lg_throttle", "membar", "sleeping", "selected"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--launch", type=int, default=0)
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    r = data[a.launch]
    col = {h: i for i, h in enumerate(hdr)}
    out = {"report": a.report.split("/")[-1], "kernel": r[col["Kernel Name"]], "note": a.note}
    for k in KEEP:
        if k in col:
            try:
                out[k] = float(r[col[k]].replace(",", ""))
            except ValueError:
                out[k] = r[col[k]]
            out[k + ".unit"] = units[col[k]]
    st = {}
    for s in STALLS:
        k = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s
        if k in col:
            st[s] = round(float(r[col[k]]), 3)
    out["warp_stalls_per_issued_instruction"] = st
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()

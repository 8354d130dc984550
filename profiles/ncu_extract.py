#!/usr/bin/env python
"""Turns an `ncu --set full` report into the per-kernel metric table and the DRAM-bytes-per-launch JSON kept under
profiles/.  Usage: python profiles/ncu_extract.py <report.ncu-rep> <out_metrics.txt> [<traffic.json> <config>]"""
import csv
import io
import json
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct",
    "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    # which execution pipe the issued instructions go to (round 2, capture P: the walk is short of the ALU pipe)
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
]
# display names of the trace kernel's launches, in launch order within a frame (restir_capi.cu)
TRACE_ORDER = ["trace_kernel<pixel>", "trace_kernel<own>", "trace_kernel<neighbours>"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units = rows[0], rows[1]
    traffic, insts, lines, n_trace = {}, {}, [], 0
    for row in rows[2:]:
        d = dict(zip(head, row))
        name = d["Kernel Name"]
        label = name.split("(")[0].replace("void ", "").replace("restir::", "")
        if label.startswith("trace_wide_kernel"):   # the kernel that walks the 4-wide image: same roles, same display names
            label = label.replace("trace_wide_kernel", "trace_kernel")
        if label.startswith("trace_kernel"):
            mode = label[label.index("<") + 1:].split(",")[0].replace("(int)", "").replace(">", "").strip()
            if mode == "2":
                label = "trace_kernel<segments>"
            elif mode == "1":
                label = "trace_kernel<neighbours>"
            else:  # pixel mode serves restirOmni's ray and the unbiased pass's own rays, alternating
                label = TRACE_ORDER[n_trace % 2]
                n_trace += 1
        else:
            fused = "<(bool)1>" in label or "<true>" in label     # the lighting pass runs inside this kernel (restir_frame_lit)
            label = label.split("<")[0]
            if fused and label in ("unbiased_finalize_kernel", "spatial_reuse_kernel", "spatial_reuse_staged_kernel"):
                label += "+lighting"
        lines.append("-----")
        lines.append(f"{'Kernel':90s} {label}   [{name[:100]}]")
        for m in METRICS:
            if m in d:
                lines.append(f"{m:90s} {d[m]} {units[head.index(m)]}")
        try:
            u_r, u_w = units[head.index("dram__bytes_read.sum")], units[head.index("dram__bytes_write.sum")]
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            traffic[label] = int(float(d["dram__bytes_read.sum"].replace(",", "")) * scale[u_r] + float(d["dram__bytes_write.sum"].replace(",", "")) * scale[u_w])
        except (KeyError, ValueError):
            pass
        try:
            insts[label] = int(float(d["smsp__inst_executed.sum"].replace(",", "")))
        except (KeyError, ValueError):
            pass
    open(out, "w").write("\n".join(lines) + "\n")
    if len(sys.argv) >= 5:
        path, config = sys.argv[3], sys.argv[4]
        try:
            doc = json.load(open(path))
        except (OSError, ValueError):
            doc = {}
        doc["_comment"] = "dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` captures (profiles/*_ncu_metrics.txt)"
        doc[config] = traffic
        doc[config + ":warp_instructions"] = insts   # smsp__inst_executed.sum per launch (bench.py: issue-slot utilisation)
        json.dump(doc, open(path, "w"), indent=1)
    print(f"{len(rows) - 2} launches -> {out}")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into the text files committed under profiles/.

    python profiles/summarize_ncu.py gpurun_out/x/prof.ncu-rep profiles/r01_name

writes <out>_metrics.txt (key raw metrics per profiled launch) and <out>_sass_mix.txt
(executed warp instructions by opcode + the hottest stall sites from the source page).
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg.per_second",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.per_cycle_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio"]


def ncu(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep, out):
    rows = ncu(rep, "raw")
    hdr, units = rows[0], rows[1]
    with open(out + "_metrics.txt", "w") as fh:
        fh.write(f"# ncu --set full --clock-control none, report {rep}\n")
        for r in rows[2:]:
            d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
            fh.write(f"\n== {d['Kernel Name']}  (ID {d['ID']})\n")
            for k in KEYS:
                if k in d:
                    fh.write(f"{k:95s} {d[k]:>18s} {u[k]}\n")
    src = ncu(rep, "source", ["--print-source", "sass"])
    hdr = None
    byop, samples, tot, sites = collections.Counter(), collections.Counter(), 0, []
    for r in src:
        if r and r[0] == "Address":
            hdr = {h: i for i, h in enumerate(r)}
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        s = r[hdr["Source"]].strip()
        ex, sm = int(r[hdr["Instructions Executed"]]), int(r[hdr["# Samples"]])
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", s)
        op = m.group(2) if m else "?"
        byop[op] += ex
        samples[op] += sm
        tot += ex
        sites.append((sm, ex, s))
    with open(out + "_sass_mix.txt", "w") as fh:
        fh.write(f"# executed warp instructions by opcode (first profiled launch of {rep}); total {tot}\n")
        ts = sum(samples.values()) or 1
        for op, c in byop.most_common(40):
            fh.write(f"{op:12s} {c:14d} {100 * c / tot:6.2f}%   stall samples {samples[op]:8d} {100 * samples[op] / ts:6.2f}%\n")
        fh.write("\n# hottest stall-sample sites: samples, executed, SASS\n")
        for sm, ex, s in sorted(sites, reverse=True)[:40]:
            fh.write(f"{sm:8d} {ex:12d}  {s}\n")
    print("wrote", out + "_metrics.txt", out + "_sass_mix.txt")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])

#!/usr/bin/env python
"""Summarise an `ncu --set full` report into a small committed text file.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/<name>.txt ["note"]

Reads the report with `ncu -i … --page raw --csv` / `--page source --csv` (no GPU needed) and keeps:
duration, DRAM bytes (traffic), tensor-pipe activity, issue activity, stall mix, registers, and the
stall samples split by kernel role (delimited by SASS landmarks of snsde_tc_kernel).
"""
import csv
import io
import re
import subprocess
import sys

KEEP = re.compile(r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|sm__cycles_elapsed\.max|"
                  r"sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)|"
                  r"sm__inst_executed_pipe_tensor_subpipe_hmma\.avg\.pct_of_peak_sustained_active|"
                  r"smsp__issue_active\.avg\.pct_of_peak_sustained_active|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
                  r"launch__(registers_per_thread|grid_size|block_size|shared_mem_per_block_dynamic)|smsp__inst_executed\.sum|"
                  r"smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|"
                  r"dram__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__throughput\.avg\.pct_of_peak_sustained_elapsed)$")


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    lines = [f"# ncu summary of {rep}", f"# {note}", ""]
    raw = ncu_csv(rep, "raw")
    hdr, units = raw[0], raw[1]
    for row in raw[2:]:
        name = row[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        lines.append(f"## kernel: {name}")
        for h, u, v in zip(hdr, units, row):
            if KEEP.match(h):
                lines.append(f"{h:95s} {v:>16s} {u}")
        lines.append("")
    src = ncu_csv(rep, "source")
    if len(src) > 2 and "# Samples" in src[1]:
        hdr = src[1]
        i_s, i_n, i_e = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
        data = src[2:]
        S = [int(r[i_n] or 0) for r in data]
        E = [int(r[i_e] or 0) for r in data]
        total = sum(S) or 1
        lines.append(f"## warp-stall samples by code region (total {total}, {len(data)} SASS instructions)")
        marks = [i for i, r in enumerate(data) if re.search(r"LDTM|TRYWAIT|UTCBAR|UBLKCP|BAR\.SYNC|FENCE\.VIEW|EXIT", r[i_s])]
        last = 0
        for i in marks + [len(data)]:
            seg = sum(S[last:i])
            if seg * 200 >= total:
                what = data[i][i_s].strip()[:60] if i < len(data) else "end"
                lines.append(f"  SASS[{last:5d}:{i:5d}) {100.0 * seg / total:5.1f}% samples, {sum(E[last:i]):>11d} warp-instr  -> up to: {what}")
            last = i
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main()

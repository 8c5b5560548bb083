#!/usr/bin/env python3
"""Condense an .ncu-rep (ncu --set full) into a small text file for profiles/: per captured launch
the duration, DRAM bytes, unit throughputs, occupancy, registers, instruction count and the top
warp-stall reasons.  usage: ncu_summary.py in.ncu-rep out.txt"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith("smsp__average_warp") and h.endswith(".ratio") and "stalled" in h]
    with open(out, "w") as f:
        f.write("# condensed from %s (ncu --set full --clock-control none), one block per captured launch\n" % rep.split("/")[-1])
        for r in rows[2:]:
            f.write("\n== %s  grid %s block %s\n" % (r[idx["Kernel Name"]], r[idx["Grid Size"]] if "Grid Size" in idx else "",
                                                   r[idx["Block Size"]] if "Block Size" in idx else ""))
            for w in WANT:
                if w in idx and r[idx[w]] != "":
                    f.write("%-72s %s %s\n" % (w, r[idx[w]], units[idx[w]]))
            st = []
            for h in stall:
                try:
                    st.append((float(r[idx[h]]), h))
                except ValueError:
                    pass
            st.sort(reverse=True)
            f.write("top warp-stall reasons (warp cycles per issued instruction):\n")
            for v, h in st[:6]:
                f.write("    %-76s %.2f\n" % (h.replace("smsp__average_", "").replace(".ratio", ""), v))
    print("wrote", out)


if __name__ == "__main__":
    main()

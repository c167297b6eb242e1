#!/usr/bin/env python
"""Key metrics of every launch in an `ncu --set full` report, as a small text table for profiles/.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/rNN_x_ncu.txt
"""
import csv
import io
import subprocess
import sys

WANT = [
    ("time_ms", "gpu__time_duration.sum"),
    ("dram_rd_MB", "dram__bytes_read.sum"),
    ("dram_wr_MB", "dram__bytes_write.sum"),
    ("dram_%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor_%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("issue_%", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("alu_%", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
    ("fma_%", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
    ("xu_%", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
    ("warps_%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("regs", "launch__registers_per_thread"),
    ("smem_dyn_KB", "launch__shared_mem_per_block_dynamic"),
    ("l2_hit_%", "lts__t_sector_hit_rate.pct"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {rep}  (ncu --set full --clock-control none; one line per captured launch)")
    print("# units: " + ", ".join(f"{k}={units[idx[m]]}" for k, m in WANT if m in idx))
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        print(f"{name}  grid={r[idx['Grid Size']]} block={r[idx['Block Size']]}")
        print("   " + "  ".join(f"{k}={r[idx[m]]}" for k, m in WANT if m in idx))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Key rows of an `ncu --set full` report as CSV (one row per captured launch): what profiles/*_ncu_full_*.csv hold.

    python tools/ncu_extract.py gpurun_out/prof_igemm_r2a.ncu-rep > profiles/r2_ncu_full_igemm.csv

Columns: duration, DRAM bytes read/written and achieved GB/s, DRAM / L2 / tensor / XU / issue utilisation, registers, dynamic
shared memory, cluster size, executed instructions. Needs the `ncu` CLI (reads the report, no GPU)."""
import csv
import io
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "smsp__inst_executed.sum", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "launch__cluster_size", "sm__warps_active.avg.pct_of_peak_sustained_active"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(head)}
    keep = ["Kernel Name", "Block Size", "Grid Size"] + [m for m in METRICS if m in col]
    w = csv.writer(sys.stdout)
    w.writerow(keep + ["dram_GBps"])
    w.writerow([units[col[k]] for k in keep] + ["GB/s"])
    for r in body:
        def val(name):
            v, u = float(r[col[name]].replace(",", "")), units[col[name]]
            scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}
            return v * scale.get(u, 1.0)
        gbps = (val("dram__bytes_read.sum") + val("dram__bytes_write.sum")) / val("gpu__time_duration.sum") * 1e-9
        w.writerow([r[col[k]] for k in keep] + [f"{gbps:.1f}"])


if __name__ == "__main__":
    main()

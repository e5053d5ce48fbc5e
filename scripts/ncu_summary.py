#!/usr/bin/env python3
"""Condense an `ncu --set full` report into the few counters DESIGN.md / bench.py quote.

    python scripts/ncu_summary.py gpurun_out/sor_rb_full.ncu-rep > profiles/<name>.txt

Runs `ncu -i <rep> --page raw --csv` (works without a GPU) and prints, per captured launch,
the kernel name, duration, DRAM traffic, pipe utilisation, occupancy limits and registers.
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True,
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# {rep}: {len(rows) - 2} captured launch(es); ncu --set full --clock-control none")
    for r in rows[2:]:
        print(f"kernel: {r[col['Kernel Name']]}  grid {r[col['Grid Size']]} block {r[col['Block Size']]}")
        for k in KEYS:
            if k in col:
                print(f"  {k:82s} {r[col[k]]:>16s} {units[col[k]]}")
        rd, wr = col.get("dram__bytes_read.sum"), col.get("dram__bytes_write.sum")
        if rd is not None and wr is not None:
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
            tot = float(r[rd]) * scale[units[rd]] + float(r[wr]) * scale[units[wr]]
            print(f"  {'traffic = dram read + write (bytes per launch)':82s} {tot:16.0f}")


if __name__ == "__main__":
    main()

"""Condense an .ncu-rep into the handful of numbers DESIGN.md / bench.py quote (run where ncu is installed).

    python scripts/ncu_summary.py profiles/<name>.ncu-rep > profiles/<name>_summary.txt
"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "derived__lts__lts2xbar_bytes.sum.per_second", "lts__t_bytes.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
    "launch__shared_mem_per_block_dynamic",
    "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"== {name[:100]}")
        for h, u, v in zip(hdr, units, vals):
            if any(h == k or h.endswith("." + k) or h.endswith(k) for k in KEYS):
                print(f"  {h} = {v} {u}")


if __name__ == "__main__":
    main(sys.argv[1])

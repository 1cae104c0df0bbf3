"""Compact per-kernel summary of an .ncu-rep (read here on the CPU box): duration, DRAM bytes, tensor/SM
activity, registers, plus the top stall sites.  Output goes under profiles/."""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "gpc__cycles_elapsed.max", "sm__cycles_active.avg", "dram__bytes_read.sum",
    "dram__bytes_write.sum", "lts__t_bytes.sum", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__inst_executed_pipe_uniform.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__cluster_size",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
]


def main(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    name_i = hdr.index("Kernel Name")
    for r in rows[2:]:
        print(f"### {r[name_i][:110]}")
        for i, h in enumerate(hdr):
            short = h.split(".TriageCompute.")[-1]
            if any(short == k or h == k for k in KEYS):
                print(f"  {short:70s} {r[i]:>16s} {units[i]}")
        print()


if __name__ == "__main__":
    main(sys.argv[1])

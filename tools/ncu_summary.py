"""Summarise an .ncu-rep (read here, no GPU needed) into a small text file for
profiles/: per kernel the duration, DRAM bytes, pipe utilisation, occupancy,
and the top warp-stall reasons.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_x.txt
"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    print("# source: {} (ncu --set full --clock-control none)".format(path))
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        print("\nkernel: {}".format(d["Kernel Name"]))
        for k in KEYS:
            if k in d:
                print("  {:<78} {:>16} {}".format(k, d[k], u[k]))
        stalls = []
        for h in hdr:
            if ("issue_stalled" in h and h.endswith("per_issue_active.ratio")
                    and "not_issued" not in h):
                try:
                    stalls.append((float(d[h].replace(",", "")), h))
                except ValueError:
                    pass
        print("  top warp stall reasons (warps stalled per issue-active cycle):")
        for v, h in sorted(stalls, reverse=True)[:6]:
            name = h.split("issue_stalled_")[1].split("_per_issue")[0]
            print("    {:<28} {:.3f}".format(name, v))


if __name__ == "__main__":
    main(sys.argv[1])

"""Print the judged metrics of an .ncu-rep capture (ncu -i ... --page raw --csv), one block per launch."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")][:100])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("  %-84s %18s %s" % (k, r[i], units[i]))
        print()


if __name__ == "__main__":
    main(sys.argv[1])

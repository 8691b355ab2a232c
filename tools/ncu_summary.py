"""Dump the headline metrics of an .ncu-rep (read here, no GPU) into a small text table for profiles/."""
import csv
import subprocess
import sys

WANT = ["launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.per_cycle_active", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    print(f"# {path}: {len(data)} launches")
    for r in data:
        print(f"\n## {r[name_i][:90]}")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"{w:80s} {r[i]:>16s} {units[i]}")


if __name__ == "__main__":
    main(sys.argv[1])

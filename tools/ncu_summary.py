"""Summarise .ncu-rep files into the text kept under profiles/ (reads them with `ncu -i ... --page raw --csv`)."""
import csv
import subprocess
import sys

WANT = """Kernel Name
gpu__time_duration.sum
dram__bytes_read.sum
dram__bytes_write.sum
lts__t_sectors.sum
lts__t_sector_hit_rate.pct
l1tex__t_sector_hit_rate.pct
l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum
l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
smsp__inst_executed.sum
sm__warps_active.avg.pct_of_peak_sustained_active
launch__registers_per_thread
launch__grid_size
launch__block_size
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
lts__throughput.avg.pct_of_peak_sustained_elapsed
l1tex__throughput.avg.pct_of_peak_sustained_elapsed
smsp__issue_active.avg.pct_of_peak_sustained_active
sm__throughput.avg.pct_of_peak_sustained_elapsed
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_membar_per_issue_active.ratio""".split("\n")

for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print(f"== {rep}")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"{w:82s} {vals[i]:>22s} {units[i]}")
        print()

"""Pick the headline metrics out of an `ncu --page raw --csv` export (one block per profiled launch)."""
import csv
import json
import sys

WANT = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__average_warp_latency_issue_stalled_barrier.ratio',
        'smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio', 'smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio',
        'smsp__average_warp_latency_issue_stalled_wait.ratio']


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        out.append({w: (r[idx[w]] + (" " + units[idx[w]] if units[idx[w]] else "")) for w in WANT if w in idx})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])

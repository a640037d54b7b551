"""Summarise an .ncu-rep (raw page) to the handful of metrics the profiles/ notes quote."""
import csv
import subprocess
import sys

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'launch__grid_size', 'launch__block_size', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'launch__waves_per_multiprocessor',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum',
        'l1tex__m_l1tex2xbar_write_bytes.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_warps', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active']


def main():
    out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[0]
    idx = [h.index(w) if w in h else -1 for w in WANT]
    for r in rows[2:]:
        print('---')
        for w, i in zip(WANT, idx):
            if i >= 0:
                print('   %-72s %s %s' % (w, r[i][:100], rows[1][i]))


if __name__ == "__main__":
    main()

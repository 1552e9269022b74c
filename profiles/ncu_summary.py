"""Summarise an .ncu-rep (raw page) into the handful of metrics the roofline discussion uses."""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum',
 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed',
 'launch__registers_per_thread','launch__occupancy_limit_registers','sm__warps_active.avg.pct_of_peak_sustained_active',
 'smsp__issue_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct',
 'lts__t_sectors_srcunit_tex_op_read.sum','lts__t_sectors_srcunit_tex_op_write.sum','smsp__inst_executed.sum',
 'smsp__thread_inst_executed_per_inst_executed.ratio',
 'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio']
def main(path):
    out = subprocess.run(['ncu','-i',path,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('## launch', r[hdr.index('ID')], r[hdr.index('Kernel Name')][:60], 'grid', r[hdr.index('Grid Size')], 'block', r[hdr.index('Block Size')])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w); print(f'{w:88s} {r[i]:>20s} {units[i]}')
if __name__ == '__main__':
    for p in sys.argv[1:]:
        print('#', p); main(p)

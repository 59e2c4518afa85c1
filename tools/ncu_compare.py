"""Side-by-side table of the metrics DESIGN.md quotes from several `ncu --page raw --csv` exports.
  python tools/ncu_compare.py a.raw.csv b.raw.csv ..."""
import csv
import sys


def load(p):
    rows = list(csv.reader(open(p)))
    h, u, b = rows[0], rows[1], rows[2]
    return {k: (v, uu) for k, uu, v in zip(h, u, b)}


tabs = [load(p) for p in sys.argv[1:]]
keys = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed']
for k in keys:
    print("%-72s" % k + "".join("%20s" % t.get(k, ('?', ''))[0][:19] for t in tabs) + "  " + tabs[0].get(k, ('', ''))[1])
print()
for k in sorted(tabs[0]):
    if 'issue_stalled' in k and 'per_issue_active' in k:
        print("%-72s" % k.replace('smsp__average_warps_issue_stalled_', 'stall ').replace('_per_issue_active.ratio', '') +
              "".join("%20s" % t.get(k, ('?', ''))[0][:19] for t in tabs))

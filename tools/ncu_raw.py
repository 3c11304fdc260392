"""Print the headline raw metrics and stall-reason breakdown of one kernel launch in an ncu report.
usage: python tools/ncu_raw.py REPORT.ncu-rep [row_index]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
row = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
d = {h[i]: rows[2 + row][i] for i in range(len(h))}
for k in ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
          'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_dynamic',
          'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
          'sm__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__cycles_active.avg',
          'sm__cycles_elapsed.max', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
          'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
          'smsp__warps_eligible.avg.per_cycle_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
          'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']:
    if k in d:
        print("%-62s %s %s" % (k, d[k], rows[1][h.index(k)]))
st = {k: float(v.replace(',', '')) for k, v in d.items()
      if k.startswith('smsp__pcsamp_warps_issue_stalled') and not k.endswith('not_issued') and v}
tot = sum(st.values()) or 1
for k, v in sorted(st.items(), key=lambda x: -x[1])[:10]:
    print('  %-40s %8.0f %5.1f%%' % (k.replace('smsp__pcsamp_warps_issue_stalled_', ''), v, 100 * v / tot))

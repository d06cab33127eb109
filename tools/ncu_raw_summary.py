"""Key numbers of an `ncu --page raw --csv` export (one kernel): stall mix, pipes, DRAM bytes, occupancy."""
import csv, sys, json
rows = list(csv.reader(open(sys.argv[1])))
h = rows[0]; u = rows[1]; v = rows[-1]
d = dict(zip(h, v)); un = dict(zip(h, u))
def f(k):
    try: return float(d[k].replace(',', ''))
    except Exception: return None
st = {n.replace('smsp__pcsamp_warps_issue_stalled_', ''): int(x) for n, x in d.items()
      if n.startswith('smsp__pcsamp_warps_issue_stalled_') and 'not_issued' not in n}
tot = sum(st.values()) or 1
print(d.get('Kernel Name'))
print('stalls: ' + ' '.join(f"{k}:{100 * x / tot:.0f}%" for k, x in sorted(st.items(), key=lambda kv: -kv[1]) if x / tot > 0.02))
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed.sum', 'sm__inst_executed.sum.per_cycle_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct']
for k in keys:
    print(f"  {k} = {d.get(k)} {un.get(k, '')}")

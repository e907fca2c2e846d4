"""Condense an .ncu-rep (read here with `ncu -i ... --page raw --csv`) into one line of key
metrics per captured launch.  Usage: python tools/ncu_extract.py file.ncu-rep [name-filter]"""
import csv, subprocess, sys, re, io
rep = sys.argv[1]; flt = sys.argv[2] if len(sys.argv) > 2 else ''
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
M = [('us', 'gpu__time_duration.sum'), ('regs', 'launch__registers_per_thread'),
     ('occ%', 'sm__warps_active.avg.pct_of_peak_sustained_active'),
     ('ipc', 'smsp__issue_active.avg.per_cycle_active'),
     ('Minst', 'smsp__inst_executed.sum'),
     ('rdMB', 'dram__bytes_read.sum'), ('wrMB', 'dram__bytes_write.sum'),
     ('dram%', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
     ('lsu%', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'),
     ('fma%', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'),
     ('bankc', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'),
     ('st:lsb', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio'),
     ('lg', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio'),
     ('mio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio'),
     ('math', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio'),
     ('bar', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio'),
     ('ssb', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio'),
     ('wait', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio'),
     ('nsel', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio'),
     ('brr', 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio'),
     ('noi', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio')]
def val(r, k):
    if k not in idx: return float('nan')
    v = r[idx[k]].replace(',', '')
    try: v = float(v)
    except ValueError: return float('nan')
    u = units[idx[k]]
    if k.startswith('dram__bytes'):
        v *= {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1, 'Gbyte': 1e3}.get(u, 1)
    if k == 'smsp__inst_executed.sum': v *= 1e-6
    if k == 'gpu__time_duration.sum': v *= {'ns': 1e-3, 'us': 1, 'ms': 1e3}.get(u, 1)
    return v
print('%-52s %-12s ' % ('kernel', 'grid') + ' '.join('%7s' % m[0] for m in M))
for r in rows[2:]:
    nm = re.sub(r'void |ssqb::|\(.*', '', r[idx['Kernel Name']])
    if flt and flt not in nm: continue
    print('%-52s %-12s ' % (nm[:52], r[idx['Grid Size']].replace(' ', '')) +
          ' '.join('%7.2f' % val(r, m[1]) for m in M))

# -*- coding: utf-8 -*-
"""Turn the scratch outputs of tools/profile_run.sh (gpurun_out/) into the tracked
summaries under profiles/ (CSV of selected ncu metrics, DRAM traffic, bench lines)."""
import csv, json, shutil, subprocess, sys, os
R = sys.argv[1] if len(sys.argv) > 1 else 'r1'
os.makedirs('profiles', exist_ok=True)
shutil.copy('gpurun_out/r1_launches.csv', f'profiles/{R}_launches.csv')
shutil.copy('gpurun_out/bench_r1.json', f'profiles/{R}_bench_n1.json')
shutil.copy('gpurun_out/bench_r1_b8.json', f'profiles/{R}_bench_n1_batch8.json')
raw = subprocess.run(['ncu', '-i', 'gpurun_out/r1_hot.ncu-rep', '--page', 'raw', '--csv'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
H, U = rows[0], rows[1]
keep = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.per_cycle_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'] + [
        'smsp__average_warps_issue_stalled_%s_per_issue_active.ratio' % k for k in
        ('barrier', 'long_scoreboard', 'short_scoreboard', 'mio_throttle', 'lg_throttle', 'wait',
         'math_pipe_throttle', 'no_instruction', 'not_selected')]
idx = [(k, H.index(k)) for k in keep if k in H]
tr = {}
with open(f'profiles/{R}_hot_kernels_ncu.csv', 'w', newline='') as f:
    w = csv.writer(f); w.writerow([k for k, _ in idx]); w.writerow([U[i] for _, i in idx])
    print("%-62s %8s %6s %7s %5s %6s %9s %9s" % ('kernel', 'us', 'grid', 'ins/pt', 'ipc', 'dram%', 'rdMB', 'wrMB'))
    for r in rows[2:]:
        w.writerow([r[i] for _, i in idx])
        o = {k: r[i] for k, i in idx}
        nm = o['Kernel Name'].replace('void ', '').replace('(FastArgs<T1>)', '')
        g = float(o['launch__grid_size'])
        elems = 4096 if '<float, 12' in nm else 8192
        ipp = float(o['smsp__inst_executed.sum']) * 32 / (g * elems)
        print("%-62s %8.1f %6d %7.1f %5s %6.1f %9.1f %9.1f" % (
            nm[:62], float(o['gpu__time_duration.sum']), g, ipp,
            o['smsp__issue_active.avg.per_cycle_active'],
            float(o['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']),
            float(o['dram__bytes_read.sum']), float(o['dram__bytes_write.sum'])))
        tr.setdefault(nm, []).append(dict(us=float(o['gpu__time_duration.sum']),
                                          dram_read_MB=float(o['dram__bytes_read.sum']),
                                          dram_write_MB=float(o['dram__bytes_write.sum'])))
flat = {k: {kk: sum(x[kk] for x in v) / len(v) for kk in v[0]} | {'launches_captured': len(v)}
        for k, v in tr.items()}
json.dump(flat, open(f'profiles/{R}_traffic.json', 'w'), indent=1)

"""Sum the per-kernel DRAM bytes of tools/traffic_step.py's ncu capture into profiles/r2_traffic.json.
Usage: python tools/traffic_sum.py gpurun_out/r2_traffic_raw.csv B [out.json]"""
import csv, json, re, sys
path, B = sys.argv[1], int(sys.argv[2])
out = sys.argv[3] if len(sys.argv) > 3 else 'profiles/r2_traffic.json'
unit = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 'usecond': 1e-6, 'nsecond': 1e-9, 'msecond': 1e-3}
per = {}
order = []
for r in csv.reader(open(path, errors='ignore')):
    if len(r) < 15 or not r[0].isdigit():
        continue
    kid, name, metric, u, val = int(r[0]), r[4], r[12], r[13], float(r[14].replace(',', ''))
    nm = re.sub(r'^void |ssqb::|\(.*', '', name)
    k = per.setdefault(kid, {'name': nm, 'grid': r[8]})
    if kid not in order:
        order.append(kid)
    k[metric] = val * unit.get(u, 1.0)
tot_r = sum(k.get('dram__bytes_read.sum', 0) for k in per.values())
tot_w = sum(k.get('dram__bytes_write.sum', 0) for k in per.values())
tot_t = sum(k.get('gpu__time_duration.sum', 0) for k in per.values())
alg = 4 * (1 + 4 * 300) * 160000 * B
by = {}
for k in per.values():
    key = re.sub(r'<.*', '', k['name'])
    d = by.setdefault(key, {'launches': 0, 'read_MB': 0.0, 'write_MB': 0.0, 'ms': 0.0})
    d['launches'] += 1
    d['read_MB'] += k.get('dram__bytes_read.sum', 0) / 1e6
    d['write_MB'] += k.get('dram__bytes_write.sum', 0) / 1e6
    d['ms'] += k.get('gpu__time_duration.sum', 0) * 1e3
interp = [k for k in per.values() if 'grid_interp' in k['name']]
rowk = [k for k in per.values() if 'sblk_rows' in k['name'] or 'cwt_rows_kernel' in k['name']]
def _per_launch(ks):
    return (sum(k.get('dram__bytes_read.sum', 0) + k.get('dram__bytes_write.sum', 0) for k in ks) / len(ks)) if ks else None
res = {'batch': B, 'algorithmic_MB': alg / 1e6, 'dram_read_MB': tot_r / 1e6, 'dram_write_MB': tot_w / 1e6,
       'traffic_over_algorithmic': (tot_r + tot_w) / alg, 'serialised_kernel_ms': tot_t * 1e3,
       'by_kernel': by,
       'dram_bytes_per_launch': {'grid_interp_with_epilogue': _per_launch(interp),
                                 'row_kernels_with_epilogue': _per_launch(rowk)},
       'launches': {'grid_interp_with_epilogue': len(interp), 'row_kernels_with_epilogue': len(rowk)},
       'note': 'ncu --cache-control none, one pass per kernel, every kernel of one step (B=%d in groups of 8 as in '
               'the timed run, %.1f GB of outputs >> 126 MB L2): the sum is the step\'s DRAM traffic; a kernel\'s '
               'own figure includes write-backs of lines dirtied by earlier kernels' % (B, alg / 1e9)}
json.dump(res, open(out, 'w'), indent=1)
print(json.dumps({k: v for k, v in res.items() if k != 'by_kernel'}, indent=1))
for k, v in sorted(by.items(), key=lambda kv: -kv[1]['ms']):
    print("%-36s x%-3d read %8.1f MB  write %8.1f MB  %.3f ms" % (k, v['launches'], v['read_MB'], v['write_MB'], v['ms']))

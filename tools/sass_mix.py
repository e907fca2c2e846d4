"""Dynamic SASS opcode mix of the kernels in an `ncu --page source --csv --print-source sass`
export: share of executed warp instructions and of stall samples per opcode."""
import csv, sys

def blocks(path):
    cur = None
    for row in csv.reader(open(path)):
        if not row:
            continue
        if row[0] == 'Kernel Name':
            cur = {'name': row[1], 'rows': []}
            yield cur
        elif row[0] == 'Address':
            cur['head'] = row
        elif cur is not None:
            cur['rows'].append(row)

for b in list(blocks(sys.argv[1])):
    h = b['head']; si = h.index('Source'); ii = h.index('Instructions Executed'); ss = h.index('# Samples')
    d, smp, tot = {}, {}, 0
    for row in b['rows']:
        op = row[si].split()
        if not op:
            continue
        o = (op[1] if op[0].startswith('@') else op[0]).split('.')[0]
        n = int(row[ii]); d[o] = d.get(o, 0) + n; tot += n
        smp[o] = smp.get(o, 0) + int(row[ss])
    ts = max(sum(smp.values()), 1)
    print(b['name'][:110], 'warp-instr', tot)
    for k, v in sorted(d.items(), key=lambda kv: -kv[1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 24]:
        print('  %-8s %6.2f%%   samples %5.1f%%' % (k, 100 * v / tot, 100 * smp[k] / ts))

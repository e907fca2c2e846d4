"""Static SASS opcode histogram of one kernel in libssq_b200.so (no GPU needed).
Usage: python tools/sass_static.py <substring of the mangled name> [top]"""
import subprocess, sys, re, collections
so = 'ssqueezepy_b200/libssq_b200.so'
out = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
cur, hist = None, {}
for ln in out.splitlines():
    m = re.match(r'\s+Function : (\S+)', ln)
    if m:
        cur = m.group(1); hist[cur] = collections.Counter(); continue
    m = re.match(r'\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)', ln)
    if m and cur:
        hist[cur][m.group(2).split('.')[0] + ('.' + '.'.join(m.group(2).split('.')[1:3]) if m.group(2).startswith(('LD', 'ST', 'RED', 'ATOM')) else '')] += 1
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for k, h in hist.items():
    if sys.argv[1] in k:
        print(k, 'total', sum(h.values()))
        print('  ' + '  '.join('%s %d' % kv for kv in h.most_common(top)))

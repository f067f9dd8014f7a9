"""Summarise an `ncu --page source --csv` dump: opcode mix, stall mix, hottest source lines."""
import collections
import csv
import re
import sys

path = sys.argv[1]
rows = list(csv.reader(open(path)))
out = []
kern = None
i = 0
blocks = []
cur = None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'hdr': None, 'rows': []}
        blocks.append(cur)
    elif cur is not None and cur['hdr'] is None and r and r[0] == 'Address':
        cur['hdr'] = r
    elif cur is not None and cur['hdr'] is not None:
        cur['rows'].append(r)
for b in blocks:
    hdr = b['hdr']
    isrc, iex, ismp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
    stalls = [k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
    byop = collections.Counter()
    st = collections.Counter()
    tot = samp = 0
    for r in b['rows']:
        if len(r) < len(hdr):
            continue
        try:
            ex = int(r[iex])
        except ValueError:
            continue
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[isrc])
        op = m.group(2).split('.')[0] if m else '?'
        byop[op] += ex
        tot += ex
        samp += int(r[ismp] or 0)
        for k in stalls:
            st[k] += int(r[hdr.index(k)] or 0)
    print(b['name'][:110])
    print('  warp instructions', tot, ' samples', samp)
    print('  ' + '  '.join('%s %.1f%%' % (op, 100.0 * c / tot) for op, c in byop.most_common(16)))
    print('  ' + '  '.join('%s %.1f%%' % (k[6:], 100.0 * v / max(samp, 1)) for k, v in st.most_common(9)))

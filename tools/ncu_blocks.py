"""Hot basic blocks of the first kernel in an `ncu --page source --csv` dump."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = None
data = []
for r in rows:
    if r and r[0] == 'Address':
        if hdr is None:
            hdr = r
            continue
        break
    if hdr is not None and len(r) >= len(hdr):
        data.append(r)
isrc, iex, ismp, ia = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Address')


def opcode(src):
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', src)
    return m.group(2) if m else '?'


blocks, cur = [], []
for r in data:
    cur.append(r)
    if opcode(r[isrc]) in ('BRA', 'BRX', 'EXIT'):
        blocks.append(cur)
        cur = []
tot = sum(int(r[iex] or 0) for r in data)
tsm = sum(int(r[ismp] or 0) for r in data)
res = []
for b in blocks:
    ex = sum(int(r[iex] or 0) for r in b)
    sm = sum(int(r[ismp] or 0) for r in b)
    ops = collections.Counter(opcode(r[isrc]) for r in b)
    res.append((ex, sm, b[0][ia], len(b), int(b[0][iex] or 0), dict(ops.most_common(4))))
res.sort(reverse=True)
for ex, sm, a, n, cnt, ops in res[:top]:
    print('%s len %4d x%-10d instr %5.1f%% samples %5.1f%%  %s' % (a, n, cnt, 100.0 * ex / tot, 100.0 * sm / max(tsm, 1), ops))

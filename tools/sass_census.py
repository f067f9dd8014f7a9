"""CPU only: census of the SASS of libqcs.so (cuobjdump -sass) -- the mnemonics that prove the Blackwell-native paths,
per kernel, and the back-to-back tcgen05.mma issue of the tensor-core kernel.  Writes profiles/r2_sass_census.txt."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'qcircuits_b200', 'libqcs.so')
KEYS = ['UTCHMMA', 'LDTM', 'STTM', 'UTCBAR', 'UTCATOMSWS', 'LDGSTS', 'UBLKCP', 'SYNCS', 'BRX', 'F2FP', 'FFMA2', 'FMUL2', 'FADD2',
        'DFMA', 'BAR.SYNC', 'ELECT', 'LD.E', 'ST.E']


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = kernels.setdefault(m.group(1), [])
            continue
        m = re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(.*?);', line)
        if m and cur is not None:
            cur.append(m.group(1).strip())
    lines = ['SASS census of qcircuits_b200/libqcs.so (cuobjdump -sass; tools/sass_census.py), round 2, final code.',
             'Mnemonics that prove the Blackwell-native paths: UTCHMMA = tcgen05.mma (kind::f16), LDTM / STTM = tcgen05.ld / st,',
             'UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc, LDGSTS = cp.async, UBLKCP = cp.async.bulk (TMA), SYNCS = mbarrier,',
             'BRX = the interpreter jump table, F2FP = the fp16 split of the tensor-core operands, LD.E / ST.E with a 64-bit',
             'descriptor in dist_pull_kernel = the NVLink peer loads / stores.', '']
    for name, ins in kernels.items():
        if not any(k in name for k in ('tile_kernel', 'dist_pull', 'dense_kernel', 'marginal', 'permute')):
            continue
        cnt = collections.Counter()
        for i in ins:
            op = re.sub(r'^@!?U?P\d+\s+', '', i).split()[0]
            for k in KEYS:
                if op == k or op.startswith(k + '.'):
                    cnt[k] += 1
        lines.append(name)
        lines.append('    %d instructions; ' % len(ins) + '  '.join('%s x%d' % (k, cnt[k]) for k in KEYS if cnt[k]))
    for name, ins in kernels.items():
        if 'tile_kernel_tc' in name and name.endswith('Li128EEEvT_'):
            idx = [j for j, i in enumerate(ins) if 'UTCHMMA' in i]
            lines += ['', 'The MMA issue of %s (instructions %d..%d of %d):' % (name, idx[0], idx[-1], len(ins))]
            lines += ['    ' + i for i in ins[idx[0] - 4:idx[-1] + 4]]
    path = os.path.join(ROOT, 'profiles', 'r2_sass_census.txt')
    open(path, 'w').write('\n'.join(lines) + '\n')
    print('\n'.join(l for l in lines if 'tile_kernel_tc' in l or 'UTCHMMA x' in l))


if __name__ == '__main__':
    main()

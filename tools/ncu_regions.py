#!/usr/bin/env python
"""Attribute the executed instructions of one profiled kernel to address ranges of its SASS.

    python tools/ncu_regions.py <report.ncu-rep> <kernel-regex> <peak_points_per_launch> name:hexstart ...

Reads `ncu -i report --page source --csv` (needs --import-source on at capture time).  Regions are
given as name:offset pairs (hex byte offset of the first instruction, from cuobjdump -sass); each
runs to the next.  Prints warp-instructions per warp-level peak-point, split FP64 / other."""
import collections
import csv
import subprocess
import sys


def main():
    rep, regex, pp = sys.argv[1], sys.argv[2], float(sys.argv[3])
    marks = [(a.split(':')[0], int(a.split(':')[1], 16)) for a in sys.argv[4:]] or [('all', 0)]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-id', '::regex:%s:1' % regex],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    hdr, body = rows[h], [r for r in rows[h + 1:] if len(r) > 6 and r[0].startswith('0x')]
    ia, isrc, ie, iss = (hdr.index(k) for k in ('Address', 'Source', 'Instructions Executed', '# Samples'))
    base = int(body[0][ia], 16)
    wpp = pp / 32
    marks.append(('end', 1 << 60))
    print('kernel region                inst/pp   fp64/pp  other/pp  samples   top opcodes (inst/pp)')
    for (name, a), (_, b) in zip(marks, marks[1:]):
        rr = [r for r in body if a <= int(r[ia], 16) - base < b]
        ops = collections.Counter()
        for r in rr:
            t = r[isrc].split()
            op = t[1] if t[0].startswith('@') else t[0]
            ops[op.split('.')[0]] += int(r[ie])
        n = sum(ops.values())
        f = sum(v for k, v in ops.items() if k in ('DFMA', 'DMUL', 'DADD', 'DSETP'))
        print('%-28s %7.3f  %7.3f  %7.3f  %7d   %s' % (name, n / wpp, f / wpp, (n - f) / wpp, sum(int(r[iss]) for r in rr),
                                                      ' '.join('%s %.2f' % (k, v / wpp) for k, v in ops.most_common(7))))


if __name__ == '__main__':
    main()

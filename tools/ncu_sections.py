# -*- coding: utf-8 -*-
""" Splits the per-instruction samples of `ncu --page source --csv` into code sections of equal execution count
(= loop nests / branches) and prints the instruction mix of the hot loop.
Usage: ncu -i X.ncu-rep --page source --csv | python tools/ncu_sections.py <kernel_index> [min_samples] """
import collections
import csv
import sys

kidx = int(sys.argv[1]) if len(sys.argv) > 1 else 0
min_s = int(sys.argv[2]) if len(sys.argv) > 2 else 300
rows = list(csv.reader(sys.stdin))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
k = starts[kidx]
hdr = rows[k + 1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[k + 2:starts[kidx + 1]] if len(r) > 10 and r[ix['# Samples']].isdigit()]
S = lambda r: int(r[ix['# Samples']] or 0)
N = lambda r: int(r[ix['Instructions Executed']] or 0)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(S(r) for r in data)
seg, cur = [], None
for i, r in enumerate(data):
    n = N(r)
    if cur is None or n != cur[0]:
        cur = [n, i, i, 0, 0, collections.Counter()]
        seg.append(cur)
    cur[2] = i
    cur[3] += S(r)
    cur[4] += 1
    for h in stalls:
        cur[5][h[6:]] += int(r[ix[h]] or 0)
print(rows[k][1][:70], 'total samples', tot, 'total inst', sum(N(r) for r in data))
for s in seg:
    if s[3] > min_s:
        top = ', '.join('%s %d' % (a, b) for a, b in s[5].most_common(4))
        print('  rows %4d-%4d exec %9d ninstr %4d samples %6d (%4.1f%%) per-instr %4.0f | %s | %s' %
              (s[1], s[2], s[0], s[4], s[3], 100. * s[3] / tot, s[3] / s[4], top, data[s[1]][ix['Source']][:36]))
hot = max(N(r) for r in data if 'DADD' in r[ix['Source']])
cls = collections.Counter()
for r in data:
    n = N(r)
    if n < hot // 20:
        continue
    op = [o for o in r[ix['Source']].split() if not o.startswith('@')][0].split('.')[0]
    cls[op] += n
chunks = max(N(r) for r in data if 'TRYWAIT' in r[ix['Source']])
print('per chunk (%d chunks): total %.0f' % (chunks, sum(cls.values()) / chunks),
      ' '.join('%s %.1f' % (a, b / chunks) for a, b in cls.most_common(24)))

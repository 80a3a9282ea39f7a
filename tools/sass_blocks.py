#!/usr/bin/env python
"""
sass_blocks.py -- opcode histogram per basic block of one kernel's SASS (cuobjdump -sass -fun <mangled> lib.so > file).
  python tools/sass_blocks.py file.sass [min_instructions]
Blocks are cut at labels and after branches; prints the blocks with at least `min_instructions` instructions.
"""
import re
import sys
from collections import Counter

pat = re.compile(r'^\s+/\*([0-9a-f]+)\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z0-9_]+)((?:\.[A-Za-z0-9_]+)*)')
blocks, cur, name = [], [], 'entry'
for line in open(sys.argv[1]):
    if re.match(r'^\s*\.L_x_\d+:', line):
        if cur:
            blocks.append((name, cur))
        cur, name = [], line.strip().rstrip(':')
        continue
    m = pat.match(line)
    if m:
        cur.append((m.group(1), m.group(2), m.group(3)))
        if m.group(2) in ('BRA', 'EXIT', 'RET', 'BRX'):
            blocks.append((name, cur))
            cur, name = [], name + '+'
if cur:
    blocks.append((name, cur))
lim = int(sys.argv[2]) if len(sys.argv) > 2 else 60
for name, ins in blocks:
    if len(ins) < lim:
        continue
    c = Counter(op for _, op, _ in ins)
    fp64 = sum(v for k, v in c.items() if k in ('DADD', 'DMUL', 'DFMA', 'DSETP'))
    print('%s @%s: %d instructions, %d fp64' % (name, ins[0][0], len(ins), fp64))
    print('   ' + ', '.join('%s %d' % kv for kv in c.most_common()))

# -*- coding: utf-8 -*-
""" Per-loop stall summary from `ncu -i X.ncu-rep --page source --csv` (development helper).
Usage: ncu -i X.ncu-rep --page source --csv | python tools/ncu_source_stalls.py [top_n] [kernel_index ...] """
import csv
import sys

top_n = int(sys.argv[1]) if len(sys.argv) > 1 else 25
only = [int(a) for a in sys.argv[2:]]
rows = list(csv.reader(sys.stdin))
kern, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'rows': []}
        kern.append(cur)
    elif cur is not None:
        cur['rows'].append(r)
seen = set()
for ki, k in enumerate(kern):
    if k['name'] in seen or (only and ki not in only):
        continue
    seen.add(k['name'])
    hdr = k['rows'][0]
    data = [r for r in k['rows'][1:] if len(r) > 10]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    val = lambda r, h: int(r[ix[h]] or 0)
    print('==', ki, k['name'])
    bars = [i for i, r in enumerate(data) if 'BAR.SYNC' in r[ix['Source']]]
    prev = 0
    for b in bars + [len(data) - 1]:
        d = data[prev:b + 1]
        ns = sum(val(r, '# Samples') for r in d)
        ni = sum(val(r, 'Instructions Executed') for r in d)
        agg = {h[6:]: sum(val(r, h) for r in d) for h in stalls}
        agg = {h: v for h, v in sorted(agg.items(), key=lambda x: -x[1]) if v > 0}
        if ns > 200:
            print('[%d..%d] samples %d inst %d' % (prev, b, ns, ni), agg)
        prev = b + 1
    for r in sorted(data, key=lambda r: -val(r, '# Samples'))[:top_n]:
        st = {h[6:]: val(r, h) for h in stalls}
        st = {h: v for h, v in sorted(st.items(), key=lambda x: -x[1]) if v > 0}
        print(data.index(r), r[ix['Address']][-5:], r[ix['# Samples']], r[ix['Instructions Executed']],
              r[ix['Source']][:50], dict(list(st.items())[:3]))

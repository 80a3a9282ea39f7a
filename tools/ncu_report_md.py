""" profiles/r2_ncu_sweepq.md and profiles/traffic.json from gpurun_out/r2_sweepq.ncu-rep (an `ncu --set full --import-source on`
capture of the x and the y q kernel on the bench batch; tools/final_profiles.sh has the command).  Needs ncu, no GPU. """
import collections, csv, io, json, os, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = os.path.join(ROOT, 'gpurun_out', 'r2_sweepq.ncu-rep')
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw))); h = r[0]; units = dict(zip(h, r[1]))
keys = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.per_cycle_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'sm__warps_active.avg.per_cycle_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']
pts = 64 * 2400 * 1200
fr = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}
out = ['# ncu capture of the q sweep kernels, round 2 (final kernels)', '',
       'Command (gpurun, one B200, clocks untouched): `ncu --set full --import-source on --clock-control none -k regex:fb_sweepq_kernel -s 6 -c 2 python tools/q_lab.py time`',
       '(bench workload: 64 fields x 2400 x 1200, N = 50000 per field, sigma 1 degree, 4 passes; the fourth x and y launch of the process).',
       'Report: `gpurun_out/r2_sweepq.ncu-rep` (scratch); this file (tools/ncu_report_md.py) holds what was read from it.  Times under ncu are',
       'cold-cache and serialised; the bench times the same kernels with CUDA events (profiles/r2_bench_line.json).', '']
traffic = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json')))
for row in r[2:]:
    d = dict(zip(h, row))
    out += ['## ' + d['Kernel Name'][:46] + ' ...', '', '| metric | value |', '|---|---|']
    out += ['| %s | %s %s |' % (k, d[k], units.get(k, '')) for k in keys if k in d]
    tot = float(d['dram__bytes_read.sum']) * fr[units['dram__bytes_read.sum']] + float(d['dram__bytes_write.sum']) * fr[units['dram__bytes_write.sum']]
    out += ['| DRAM bytes per grid point (read + write) | %.2f |' % (tot / pts), '']
    traffic['sweep_x_bytes_per_point' if ', 1>' in d['Kernel Name'] else 'sweep_y_bytes_per_point'] = tot / pts
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
starts = [i for i, x in enumerate(rows) if x and x[0] == 'Kernel Name'] + [len(rows)]
out += ['## Warp stall samples (source page, all samples) and instruction mix of the chunk loop', '']
seen = set()
for kidx in range(len(starts) - 1):
    k = starts[kidx]; hdr = rows[k + 1]; ix = {x: i for i, x in enumerate(hdr)}
    if 'Instructions Executed' not in ix or rows[k][1] in seen:
        continue
    seen.add(rows[k][1])
    data = [x for x in rows[k + 2:starts[kidx + 1]] if len(x) > 10 and x[ix['# Samples']].isdigit()]
    if sum(int(x[ix['Instructions Executed']]) for x in data) == 0:
        continue
    c = collections.Counter()
    for x in data:
        for hh in hdr:
            if hh.startswith('stall_') and 'Not' not in hh:
                c[hh[6:]] += int(x[ix[hh]] or 0)
    tot = sum(c.values())
    out.append('**' + rows[k][1][:46] + ' ...**: ' + ', '.join('%s %.1f %%' % (a, 100 * b / tot) for a, b in c.most_common(11)))
    chunks = max(int(x[ix['Instructions Executed']]) for x in data if 'TRYWAIT' in x[ix['Source']])
    mix = collections.Counter()
    for x in data:
        m = [t for t in x[ix['Source']].strip().split() if not t.startswith('@')]
        mix[m[0].split('.')[0]] += int(x[ix['Instructions Executed']])
    out += ['', 'instructions per 8-row chunk and warp (%d chunks): total %.0f — ' % (chunks, sum(mix.values()) / chunks) +
            ', '.join('%s %.1f' % (a, b / chunks) for a, b in mix.most_common(16)), '']
out += ['## Reading', '',
        '* x sweep (7 warps per CTA, 4 staging slots): its warps wait for TMA rows (`long_sb`) more than for anything else: memory bound.',
        '* y sweep: the algorithmic 20 B of DRAM traffic per point; issue slot busy about half of the cycles with 2 warps per scheduler; stalls are',
        '  fixed-latency waits, the fp64 pipe, `not_selected` (the other warp issued), shared-memory results, instruction fetch at the loop back edge.',
        '* DESIGN.md section 4.3 has the analysis (warps-per-CTA series, the variants that were measured and dropped).']
open(os.path.join(ROOT, 'profiles', 'r2_ncu_sweepq.md'), 'w').write('\n'.join(out) + '\n')
traffic['source'] = 'fp64: profiles/r2_ncu_sweepq.md (ncu --set full --clock-control none, fb_sweepq_kernel<4,1,1> / <4,1,2>, 64 fields per launch, dram__bytes_read.sum + dram__bytes_write.sum); fp32: profiles/r1_ncu_full_sweeps_fields64.csv'
json.dump(traffic, open(os.path.join(ROOT, 'profiles', 'traffic.json'), 'w'), indent=1)
print('\n'.join(out[:60]))

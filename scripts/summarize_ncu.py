"""Turn ncu outputs brought back in gpurun_out/ into small tracked summaries under profiles/.

    python scripts/summarize_ncu.py launches gpurun_out/launches_X.csv profiles/r1_launches_X.md
    python scripts/summarize_ncu.py full     gpurun_out/X.ncu-rep      profiles/r1_X.md
"""
import collections
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__waves_per_multiprocessor', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.per_cycle_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.max']


def launches(src, dst):
    with open(src) as f:
        lines = [l for l in f if l.startswith('"')]
    agg = collections.OrderedDict()
    total = 0.0
    for row in csv.DictReader(lines):
        name = row['Kernel Name'].split('(')[0].replace('oadg::<unnamed>::', '').replace('oadg::', '')
        us = float(row['Metric Value']) / 1e3
        agg.setdefault(name, []).append((us, row['Grid Size'], row['Block Size']))
        total += us
    with open(dst, 'w') as out:
        out.write('# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised)\n\n')
        out.write('source: `%s`; total %.1f us over %d launches\n\n' % (src, total, sum(len(v) for v in agg.values())))
        out.write('| kernel | launches | sum us | share | min us | max us | example grid x block |\n|---|---|---|---|---|---|---|\n')
        for k, v in sorted(agg.items(), key=lambda kv: -sum(x[0] for x in kv[1])):
            ts = [x[0] for x in v]
            out.write('| %s | %d | %.1f | %.1f%% | %.1f | %.1f | %s x %s |\n' %
                      (k[-60:], len(v), sum(ts), 100 * sum(ts) / total, min(ts), max(ts), v[0][1], v[0][2]))


def full(src, dst):
    raw = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, 'w') as out:
        out.write('# ncu --set full capture: `%s`\n\n' % src)
        for r in rows[2:]:
            name = r[hdr.index('Kernel Name')]
            out.write('## %s  grid %s block %s\n\n' % (name[:90], r[hdr.index('Grid Size')], r[hdr.index('Block Size')]))
            for k in KEYS:
                if k in hdr and r[hdr.index(k)] not in ('', 'n/a'):
                    out.write('- %s = %s %s\n' % (k, r[hdr.index(k)], units[hdr.index(k)]))
            st = []
            for i, h in enumerate(hdr):
                if 'issue_stalled' in h and h.endswith('.ratio'):
                    try:
                        st.append((float(r[i]), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
                    except ValueError:
                        pass
            out.write('- top stall reasons (warps per issue-active): ' +
                      ', '.join('%s %.2f' % (h, v) for v, h in sorted(st, reverse=True)[:5]) + '\n\n')
    if len(sys.argv) > 4:   # optional: DRAM traffic of the first captured launch as JSON (bench.py's roofline.traffic)
        import json
        r = rows[2]

        def val(k):
            v, u = float(r[hdr.index(k)]), units[hdr.index(k)].lower()
            return v * {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}.get(u, 1)
        json.dump({'kernel': r[hdr.index('Kernel Name')][:80], 'dram_read_bytes': val('dram__bytes_read.sum'),
                   'dram_write_bytes': val('dram__bytes_write.sum'), 'gpu_time_us': float(r[hdr.index('gpu__time_duration.sum')]),
                   'source': src, 'note': 'one ncu --set full capture of one launch (its plan differs from the bench average)'},
                  open(sys.argv[4], 'w'), indent=1)


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2], sys.argv[3])

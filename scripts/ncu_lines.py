"""Aggregate the warp-stall samples of an ncu source page (cuda,sass view, csv) by source line.
    ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > src.csv; python scripts/ncu_lines.py src.csv [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
agg = collections.Counter()
inst = collections.Counter()
stall = collections.defaultdict(collections.Counter)
lines = {}
names = ('stall_barrier', 'stall_long_sb', 'stall_short_sb', 'stall_wait', 'stall_mio', 'stall_math', 'stall_selected',
         'stall_not_selected', 'stall_sleep', 'stall_membar', 'stall_branch_resolving', 'stall_no_inst', 'stall_lg', 'stall_dispatch')
for r in rows:
    if len(r) >= 2 and r[0] == 'Line No':
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0] and r[0].isdigit():
        ln = int(r[0])
        try:
            s = int(r[4])
        except ValueError:
            continue
        agg[ln] += s
        lines[ln] = r[1]
        try:
            inst[ln] += int(r[hdr.index('Instructions Executed')])
        except ValueError:
            pass
        for name in names:
            if name in hdr:
                try:
                    stall[ln][name] += int(r[hdr.index(name)])
                except ValueError:
                    pass
tot = sum(agg.values())
itot = sum(inst.values())
print('total samples %d, warp instructions %d' % (tot, itot))
tall = collections.Counter()
for ln in stall:
    tall.update(stall[ln])
print('stalls overall: ' + ', '.join('%s=%.1f%%' % (n.replace('stall_', ''), 100.0 * c / tot) for n, c in tall.most_common(8)))
for ln, v in agg.most_common(top):
    st = ', '.join('%s=%d' % (n.replace('stall_', ''), c) for n, c in stall[ln].most_common(3))
    print('%5.1f%% smp %5.1f%% inst  line %4d  %s   [%s]' % (100.0 * v / tot, 100.0 * inst[ln] / max(itot, 1), ln, lines[ln].strip()[:80], st))

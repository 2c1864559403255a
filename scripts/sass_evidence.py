"""Count the TMA / mbarrier / tcgen05 SASS mnemonics per kernel of libOADG.so (dev container, no GPU needed):

    python scripts/sass_evidence.py > profiles/r2_sass_tma_tcgen05.txt
"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else 'oadg_b200/libOADG.so'
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
pat = re.compile(r'/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)')
want = re.compile(r'^(UTMALDG|UTMASTG|UTMACCTL|UBLKCP|UTCHMMA|UTCQMMA|UTCIMMA|UTCBAR|UTCCP|LDTM|STTM|SYNCS|UTMAPF|FENCE\.VIEW\.ASYNC)')
counts, first, fn = collections.OrderedDict(), {}, None
for line in sass.splitlines():
    if 'Function :' in line:
        fn = line.split('Function :')[1].strip()
        counts.setdefault(fn, collections.Counter())
        continue
    m = pat.search(line)
    if m and fn and want.match(m.group(1)):
        counts[fn][m.group(1)] += 1
        first.setdefault((fn, m.group(1)), line.strip())
names = subprocess.run(['c++filt'], input='\n'.join(counts), capture_output=True, text=True).stdout.splitlines()
print('# cuobjdump -sass %s: TMA (UTMALDG = cp.async.bulk.tensor), mbarrier (SYNCS) and tcgen05 (UTC*MMA, LDTM) '
      'instructions per kernel\n' % lib)
for (fn, c), name in zip(counts.items(), names):
    if not c:
        continue
    name = re.sub(r'\(anonymous namespace\)::', '', name)
    print(name[:150])
    for k, v in sorted(c.items()):
        print('    %-28s x %-4d e.g. %s' % (k, v, first[(fn, k)][:110]))
    print()

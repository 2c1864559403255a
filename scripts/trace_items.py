"""Timeline of the chain kernel's work items for one bench batch (GPU box; OADG_TRACE=1): when every item started and
finished, and which dependency released it last -- the critical path of the queue."""
import ctypes
import os
import sys

os.environ['OADG_TRACE'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from oadg_b200 import OAMix, _lib  # noqa: E402

KINDS = ('profile', 'mask', 'hist', 'lut', 'copy', 'bbo_blend', 'bbo_catchup', 'step')


class Trace(ctypes.Structure):
    _fields_ = [('kind', ctypes.c_int32), ('obj', ctypes.c_int32), ('ntiles', ctypes.c_int32), ('dep_count', ctypes.c_int32),
                ('t0', ctypes.c_double), ('t1', ctypes.c_double), ('deps', ctypes.c_int32 * 8)]


dev = torch.device('cuda:0')
G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
frames = [bench.make_image(s) for s in range(max(8, G))]
imgs = [torch.from_numpy(f).to(dev) for f, _ in frames]
gts = [g for _, g in frames]
mix = OAMix(**bench.OAMIX_CFG)
np.random.seed(1000)
for i in range(3):
    mix.oamix_batch(imgs[0:2], gts[0:2])
prof = {}
mix.oamix_batch(imgs[0:G], gts[0:G], profile=prof)
lib = _lib.load()
lib.oadg_oamix_last_trace.restype = ctypes.c_int
n = lib.oadg_oamix_last_trace(None, 0)
buf = (Trace * n)()
lib.oadg_oamix_last_trace(buf, n)
print('chain %.1f us, %d items' % (prof['chain_ms'] * 1e3, n))
end = max(t.t1 for t in buf)
# walk the critical path backwards from the item that finished last
k = max(range(n), key=lambda i: buf[i].t1)
path = []
while k >= 0:
    t = buf[k]
    deps = [d for d in t.deps[:min(t.dep_count, 8)] if d >= 0]
    crit = max(deps, key=lambda d: buf[d].t1) if deps else -1
    path.append((k, KINDS[t.kind], t.obj, t.ntiles, t.t0, t.t1, buf[crit].t1 if crit >= 0 else 0.0))
    k = crit
print('critical path (item, kind, obj, tiles, first claim us, last publish us, released at us):')
for row in reversed(path):
    print('  %4d %-12s obj %3d tiles %5d  start %7.1f  end %7.1f  (deps done %7.1f, ran %6.1f, waited %5.1f)' %
          (row + (row[5] - row[4], row[4] - row[6])))

# occupancy of the queue over time: items in flight and tiles of items that are ready (dependencies done) but not finished
import collections
bins = collections.OrderedDict()
step = max(end / 40.0, 1.0)
ready = []
for i in range(n):
    t = buf[i]
    deps = [d for d in t.deps[:min(t.dep_count, 8)] if d >= 0]
    ready.append(max([buf[d].t1 for d in deps]) if deps else 0.0)
print('time us : items in flight / tiles of ready-or-running items / items ready but not started')
for b in range(int(end / step) + 1):
    lo, hi = b * step, (b + 1) * step
    mid = 0.5 * (lo + hi)
    run = [i for i in range(n) if buf[i].t0 <= mid < buf[i].t1]
    avail = [i for i in range(n) if ready[i] <= mid < buf[i].t1]
    idle = [i for i in range(n) if ready[i] <= mid < buf[i].t0]
    print('  %7.1f : %4d / %6d / %4d   kinds in flight: %s' % (mid, len(run), sum(buf[i].ntiles for i in avail), len(idle),
          ' '.join('%s=%d' % (KINDS[k], sum(1 for i in run if buf[i].kind == k)) for k in range(8) if any(buf[i].kind == k for i in run))))

"""Discrete-event model of the chain kernel's work queue (dev container, no GPU): the exact item / dependency tables
the scheduler builds for seeded bench plans (through tests/hostsim), P CTAs that take tiles in ticket (= readiness)
order, per-kind tile durations measured on the B200.  Answers "what does the queue's shape cost": idle share, the
effect of views per launch, tile durations, priorities.

    python scripts/queue_sim.py [views_per_launch] [n_launches]"""
import ctypes
import heapq
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np  # noqa: E402

import bench  # noqa: E402
from conftest import build_hostsim  # noqa: E402
from oadg_b200 import OAMix  # noqa: E402

KINDS = ('profile', 'mask', 'hist', 'lut', 'copy', 'bbo_blend', 'bbo_catchup', 'step')


def load_queue(group, seed_base, launch):
    lib = ctypes.CDLL(build_hostsim())
    mix = OAMix(**bench.OAMIX_CFG)
    frames = [bench.make_image(s, 64, 128) for s in range(2)]   # tiny pixels: only the plan geometry matters
    gts = []
    for i in range(group):
        _, g = bench.make_image((launch * group + i) % 24)
        gts.append(g)
    np.random.seed(seed_base + launch)
    scores = [[20.0] * len(g) for g in gts]
    plan = mix.sample_plan([(bench.H, bench.W)] * group, gts, scores)
    blob = plan.blob
    # the scheduler only looks at pointers: hand it one dummy buffer per image / view
    dummy = np.zeros(16, np.uint8)
    src = (ctypes.c_void_p * group)(*[dummy.ctypes.data + 0] * group)
    dst = (ctypes.c_void_p * group)(*[dummy.ctypes.data + 0] * group)
    cap = 1 << 22
    out = np.zeros(cap, np.int32)
    lib.hostsim_oamix_dump.restype = ctypes.c_int
    n = lib.hostsim_oamix_dump(ctypes.c_void_p(blob.ctypes.data), ctypes.c_size_t(blob.nbytes), src, group, dst,
                               ctypes.c_void_p(out.ctypes.data), cap)
    assert n > 0, n
    items, i = [], 1
    for _ in range(out[0]):
        kind, obj, nt, aux, streaming, c0, c1, c2, c3, nd = (int(v) for v in out[i:i + 10])
        deps = [int(v) for v in out[i + 10:i + 10 + nd]]
        items.append(dict(kind=kind, obj=obj, ntiles=nt, aux=aux, streaming=streaming, cls=(c0, c1, c2, c3), deps=deps))
        i += 10 + nd
    return items, int(plan.depth_sums.sum())


def simulate(items, P, dur, overhead=0.0, priority=False, publish_lat=0.0, claim_ahead=False):
    """dur(item) -> us per tile.  publish_lat: delay between a tile's end and its successors' readiness;
    claim_ahead: a CTA takes its next ticket when it STARTS a tile (the scheduler-warp kernel).  Returns
    (makespan us, busy CTA-us)."""
    n = len(items)
    pend = [0] * n
    succ = [[] for _ in range(n)]
    for k, it in enumerate(items):
        for d in it['deps']:
            succ[d].append(k)
            pend[k] += items[d]['ntiles']
    bottom = [0.0] * n
    if priority:
        for k in range(n - 1, -1, -1):
            bottom[k] = dur(items[k]) + max([bottom[s] for s in succ[k]] + [0.0])
    ready = []      # heap of (order key, item)
    seq = [0]
    next_tile = [0] * n

    def push(k):
        if items[k]['ntiles'] > 0:
            heapq.heappush(ready, ((-bottom[k], seq[0]) if priority else (seq[0],), k))
            seq[0] += 1

    def take():     # next ready tile's item, or None
        if not ready:
            return None
        _, k = ready[0]
        next_tile[k] += 1
        if next_tile[k] >= items[k]['ntiles']:
            heapq.heappop(ready)
        return k

    def avail():    # ready, unclaimed tiles
        return sum(items[k]['ntiles'] - next_tile[k] for _, k in ready)

    for k in range(n):
        if pend[k] == 0:
            push(k)
    # events: (time, type, payload): type 0 = tile finished on cta, 1 = successors released
    events = []
    t, busy = 0.0, 0.0
    held = [None] * P       # claim-ahead: the tile a CTA holds for after its current one
    running = [False] * P
    waiting = []            # idle CTAs (no tile, nothing ready)

    def start(c, k):
        nonlocal busy
        d = dur(items[k]) + overhead
        busy += d
        running[c] = True
        heapq.heappush(events, (t + d, 0, c, k))
        if claim_ahead is True or (claim_ahead and avail() > claim_ahead):
            held[c] = take()

    for c in range(P):
        k = take()
        if k is None:
            waiting.append(c)
        else:
            start(c, k)
    while events:
        t, typ, c, k = heapq.heappop(events)
        if typ == 0:
            running[c] = False
            if publish_lat > 0:
                heapq.heappush(events, (t + publish_lat, 1, -1, k))
            else:
                for s2 in succ[k]:
                    pend[s2] -= 1
                    if pend[s2] == 0:
                        push(s2)
            nk = held[c] if claim_ahead else None
            held[c] = None
            if nk is None:
                nk = take()
            if nk is None:
                waiting.append(c)
            else:
                start(c, nk)
        else:
            for s2 in succ[k]:
                pend[s2] -= 1
                if pend[s2] == 0:
                    push(s2)
        while waiting and ready:
            c2 = waiting.pop()
            start(c2, take())
    return t, busy


# measured per-tile durations (us), scheduler-warp kernel, 3 CTAs per SM (profiles/r2_*):
DUR = dict(profile=6.8, mask=4.5, hist=16.0, lut=4.6, copy=10.2, bbo_blend=11.0, bbo_catchup=10.0,
           step_stream=5.3, step_other=21.0)


def dur_of(scale=1.0):
    def f(it):
        k = KINDS[it['kind']]
        if k == 'step':   # average over the item's tile classes (stream / per pixel / bg-only / box edge)
            c = it['cls']
            w = 512.0 / max(it['aux'], 1)   # durations were measured on 512-wide stream tiles and 256-wide others
            return scale * (c[0] * DUR['step_stream'] / (w if it['aux'] < 512 else 1.0) + c[1] * 19.0 + c[2] * 24.0 + c[3] * 19.0) / max(sum(c), 1)
        return scale * DUR[k]
    return f


if __name__ == '__main__':
    group = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    n_launch = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    P = 444
    tot = {}
    for launch in range(n_launch):
        items, dsum = load_queue(group, 1000, launch)
        tiles = sum(it['ntiles'] for it in items)
        for name, kw in (('fifo', {}), ('priority', dict(priority=True)), ('fifo, tiles x0.5', dict()),
                         ('fifo, P=inf', {}), ('fifo, publish 5us', dict(publish_lat=5.0)),
                         ('fifo, publish 20us', dict(publish_lat=20.0)), ('fifo, claim ahead', dict(claim_ahead=True)),
                         ('fifo, ahead+pub5', dict(claim_ahead=True, publish_lat=5.0)),
                         ('prio, ahead+pub5', dict(claim_ahead=True, publish_lat=5.0, priority=True)),
                         ('fifo, ahead>200+pub5', dict(claim_ahead=200, publish_lat=5.0)),
                         ('fifo, ahead>800+pub5', dict(claim_ahead=800, publish_lat=5.0)),
                         ('fifo, noahead+pub5, +1.5us/tile', dict(publish_lat=5.0, overhead=1.5))):
            d = dur_of(0.5 if 'x0.5' in name else 1.0)
            span, busy = simulate(items, 10 ** 6 if 'inf' in name else P, d, **kw)
            tot.setdefault(name, []).append((span, busy))
        print('launch %d: %d items, %d tiles, depth sum %d' % (launch, len(items), tiles, dsum))
    for name, v in tot.items():
        span = sum(x[0] for x in v)
        busy = sum(x[1] for x in v)
        print('%-18s makespan %8.1f us/view, busy share %5.1f %%' % (
            name, span / (n_launch * group), 100.0 * busy / (span * P) if 'inf' not in name else 0.0))

"""world_size-2 gloo tests (CPU) of the multi-GPU host logic:
  * OA-Loss with the RoI-embedding all-gather (oadg_b200/distributed.py): the collectives, the global pair map and
    the W x gradient convention, with a numpy stand-in for the CUDA kernels -- checked against the single-process
    oracle on the concatenated batch;
  * OA-Mix sharding by image: each rank draws its own plans from its own RNG stream, no collective, and the
    union of the per-rank results equals the single-process run image by image."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oamix_np, supcon_np, synth


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


class NumpyBackend(__import__('oadg_b200.distributed', fromlist=['UnpackedBackend']).UnpackedBackend):
    """CPU stand-in with the same decomposition as the kernels: local anchors x gathered contrasts, per-row
    statistics (lse, coef, n_pos, u), backward from the statistics of ALL rows."""

    def normalize(self, x, n_total, normalized_input):
        v = x.double().numpy()
        self.n1 = np.maximum(np.sqrt((v * v).sum(1, keepdims=True)), 1e-12) if normalized_input else np.ones((len(v), 1))
        f1 = v / self.n1
        self.n2 = np.maximum(np.sqrt((f1 * f1).sum(1, keepdims=True)), 1e-12)
        self.f1 = f1
        return torch.from_numpy(f1 / self.n2)

    @staticmethod
    def _P(y, pair, rows):
        bg = y.max()
        fg = y != bg
        P = (y[rows, None] == y[None, :]) & fg[rows, None] & fg[None, :]
        P[np.arange(len(rows)), rows] = False
        for k, i in enumerate(rows):
            if not fg[i] and pair[i] >= 0 and not fg[pair[i]]:
                P[k, pair[i]] = True
        return P.astype(np.float64), fg

    def forward(self, f_all, labels_all, pair_all, row0, n_rows, temperature, loss_weight, min_samples):
        f, y, pair = f_all.numpy(), labels_all.numpy(), pair_all.numpy()
        n = len(f)
        rows = np.arange(row0, row0 + n_rows)
        P, fg = self._P(y, pair, rows)
        if fg.sum() <= min_samples:
            return torch.zeros((), dtype=torch.float64), torch.zeros(n_rows, 4, dtype=torch.float64)
        z = f[rows] @ f.T / temperature
        e = np.exp(z)
        e[np.arange(n_rows), rows] = 0
        lse = np.log(e.sum(1))
        npos = P.sum(1)
        den = np.where(npos > 0, npos, 1.0)
        row = np.where(npos > 0, (P * z).sum(1) / den - lse, 0.0)
        coef = np.where(npos > 0, -(loss_weight / n) / den, 0.0)
        stats = np.stack([lse, coef, npos, coef * npos * np.exp(-lse)], 1)
        return torch.tensor(-loss_weight * row.sum() / n), torch.from_numpy(stats)

    def backward(self, x, f_all, labels_all, pair_all, stats_all, row0, temperature, normalized_input, grad):
        f, y, pair, st = f_all.numpy(), labels_all.numpy(), pair_all.numpy(), stats_all.numpy()
        n_rows = x.shape[0]
        rows = np.arange(row0, row0 + n_rows)
        if not st.any():
            return torch.zeros_like(x)
        Pij, _ = self._P(y, pair, rows)                    # P[i, j], i local
        Pall, _ = self._P(y, pair, np.arange(len(f)))      # full, to read P[j, i]
        Pji = Pall[:, rows].T
        z = f[rows] @ f.T / temperature
        A = st[rows, 1][:, None] * Pij + st[None, :, 1] * Pji - np.exp(z) * (st[rows, 3][:, None] + st[None, :, 3])
        A[np.arange(n_rows), rows] = 0
        df = A @ f / temperature
        u2 = self.f1 / self.n2
        g1 = (df - (u2 * df).sum(1, keepdims=True) * u2) / self.n2
        if normalized_input:
            u1 = x.double().numpy() / self.n1
            g1 = (g1 - (u1 * g1).sum(1, keepdims=True) * u1) / self.n1
        return torch.from_numpy(g1 * float(grad))


def _loss_worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oadg_b200.distributed import gathered_contrastive_loss
    torch.manual_seed(100 + rank)
    x = torch.randn(n, 256, dtype=torch.float64, requires_grad=True)
    g = torch.Generator().manual_seed(7 + rank)
    half = n // 2
    base = torch.full((half,), 5, dtype=torch.int64)
    idx = torch.randperm(half, generator=g)[:half // 3]
    base[idx] = torch.randint(0, 5, (len(idx),), generator=g)
    labels = torch.cat([base, base])
    pair_local = np.concatenate([np.arange(half) + half, np.arange(half)])
    loss = gathered_contrastive_loss(x, labels, temperature=0.06, loss_weight=0.01, min_samples=10,
                                     pair_local=pair_local, backend=NumpyBackend())
    loss.backward()
    q.put((rank, x.detach().numpy(), labels.numpy(), float(loss), x.grad.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_gathered_loss_equals_single_process_oracle_on_the_concatenated_batch():
    world, n = 2, 96
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_loss_worker, args=(r, world, port, n, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=120) for _ in range(world)])
    [p.join(60) for p in procs]
    x_all = np.concatenate([r[1] for r in res])
    y_all = np.concatenate([r[2] for r in res])
    half = n // 2
    pair_local = np.concatenate([np.arange(half) + half, np.arange(half)])
    from oadg_b200.distributed import gathered_pair_map
    pair_all = gathered_pair_map(pair_local, world)
    assert pair_all[n + 3] == n + 3 + half and pair_all[half] == 0
    ref, gref = supcon_np.supcon_loss(x_all, y_all, 0.06, 10, 0.01, want_grad=True, pair=pair_all)
    for rank, _, _, loss, grad in res:
        assert abs(loss - ref) <= 1e-12 * abs(ref)                       # every rank reports the global-mean loss
        want = world * gref[rank * n:(rank + 1) * n]                      # W x dL/dx_r (DDP averages by 1/W)
        assert np.linalg.norm(grad - want) <= 1e-9 * np.linalg.norm(want)


def _mix_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oadg_b200.oamix import OAMix
    t = OAMix(version='augmix')
    out = []
    for i in range(2):                                   # this rank's shard: images rank*2 + i
        idx = rank * 2 + i
        img, gt = synth.make_image(idx, 96, 160, 3)
        np.random.seed(500 + idx)                        # per-sample stream, as the reference reseeds its workers
        scores = oamix_np.fg_scores(img, gt)
        vp = t._sample_head(96, 160, gt)
        t._sample_tail(vp, gt, scores)
        out.append((idx, vp.ml_boxes.tolist(), [b.tolist() for b in vp.oa_boxes], float(vp.m)))
    gathered = [None] * world
    dist.all_gather_object(gathered, out)                # test-only: collect for comparison (no data-path collective)
    if rank == 0:
        q.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


def test_oamix_shards_by_image_without_a_collective():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_mix_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    gathered = q.get(timeout=120)
    [p.join(60) for p in procs]
    flat = sorted(x for part in gathered for x in part)
    assert [x[0] for x in flat] == [0, 1, 2, 3]
    for idx, ml, oa, m in flat:                          # identical to a single-process run of the same image
        img, gt = synth.make_image(idx, 96, 160, 3)
        np.random.seed(500 + idx)
        plan = oamix_np.sample_plan(img, gt, version='augmix')
        assert ml == plan['ml_boxes'].tolist() and oa == [b.tolist() for b in plan['oa_boxes']] and m == plan['m']


def test_peer_exchange_layout_sections_do_not_overlap():
    """The byte layout of a rank's exchange buffer (PeerExchange.layout): every section inside the buffer, aligned,
    disjoint, flags large enough for OADG_PEER_MAX sources."""
    from oadg_b200.distributed import PeerExchange, PEER_MAX
    for world in (2, 3, 8, PEER_MAX):
        for n in (1, 2048, 2088):
            lay = PeerExchange.layout(world, n, 260)
            spans = [(lay['off_rows'][h], world * n * 260 * 4) for h in (0, 1)]
            spans += [(lay['off_tail'][h], world * (n + 1) * 16) for h in (0, 1)]
            spans += [(lay['off_flag_rows'], 4 * PEER_MAX), (lay['off_flag_tail'], 4 * PEER_MAX), (lay['off_counter'], 4)]
            spans.sort()
            for (a, la), (b, _) in zip(spans, spans[1:]):
                assert a + la <= b, (world, n, spans)
            assert spans[-1][0] + spans[-1][1] <= lay['bytes']
            assert all(a % 16 == 0 for a, _ in spans)

"""OA-Loss CUDA path (plugin -> C ABI) against the oracle (float64 restatement pinned to the reference).
Tolerance from BASELINE.json north_star: 1e-5 relative on the loss and on the gradient."""
import numpy as np
import pytest

from conftest import GOLDEN
from oracle import supcon_np, synth

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _run(cuda, x, labels, **kw):
    import torch
    from oadg_b200 import ContrastiveLossPlus
    loss_fn = ContrastiveLossPlus(loss_weight=0.01, num_views=2, temperature=0.06, **kw)
    xd = x.to(cuda).requires_grad_(True)
    loss = loss_fn(xd, labels.to(cuda))
    if loss.requires_grad:
        loss.backward()
    return loss, xd.grad, loss_fn


@pytest.mark.parametrize('n', [2048, 2088, 2085])
def test_loss_and_grad_match_oracle_and_reference_golden(cuda, n):
    x, labels = synth.make_roi_set(n)
    loss, grad, fn = _run(cuda, x, labels)
    assert loss.dim() == 0 and loss.device.type == 'cuda' and fn.stats['launches'] >= 6
    ref, gref = supcon_np.supcon_loss(x.numpy(), labels.numpy(), 0.06, 10, 0.01, want_grad=True)
    assert abs(loss.item() - ref) <= RTOL * abs(ref)
    g = grad.cpu().numpy().astype(np.float64)
    assert np.linalg.norm(g - gref) <= RTOL * np.linalg.norm(gref)
    G = np.load(GOLDEN + '/supcon.npz')   # produced by the unmodified reference (f32 torch path)
    assert abs(loss.item() - float(G['n%d/f32/loss' % n])) <= RTOL * abs(ref)
    rows = np.concatenate([g[0:8], g[1024:1032], g[n - 8:n]])
    assert np.linalg.norm(rows - G['n%d/f64/grad_rows' % n]) <= RTOL * np.linalg.norm(G['n%d/f64/grad_rows' % n])


def test_upstream_gradient_and_dtype(cuda):
    import torch
    from oadg_b200 import ContrastiveLossPlus
    x, labels = synth.make_roi_set(2088, seed=3)
    fn = ContrastiveLossPlus(loss_weight=1.0, temperature=0.07)
    xd = x.to(cuda).requires_grad_(True)
    (fn(xd, labels.to(cuda)) * 3.0).backward()
    ref, gref = supcon_np.supcon_loss(x.numpy(), labels.numpy(), 0.07, 10, 1.0, want_grad=True)
    assert np.linalg.norm(xd.grad.cpu().numpy() - 3.0 * gref) <= RTOL * np.linalg.norm(3.0 * gref)


def test_quirks(cuda):
    import torch
    from oadg_b200 import ContrastiveLossPlus
    # #fg <= min_samples -> exactly 0 and zero gradient (contrastive_loss.py:211,229-230)
    x, labels = synth.make_roi_set(2048, n_fg=5)
    loss, grad, _ = _run(cuda, x, labels)
    assert loss.item() == 0.0 and (grad is None or float(grad.abs().max()) == 0.0)
    # no bg RoI in the batch: the highest fg class silently plays bg (contrastive_loss.py:197-200)
    x, labels = synth.make_roi_set(2048, n_fg=1024, n_cls=4)
    labels = labels.clamp(max=3)
    loss, grad, _ = _run(cuda, x, labels)
    ref, gref = supcon_np.supcon_loss(x.numpy(), labels.numpy(), 0.06, 10, 0.01, want_grad=True)
    assert abs(loss.item() - ref) <= RTOL * abs(ref)
    assert np.linalg.norm(grad.cpu().numpy() - gref) <= RTOL * np.linalg.norm(gref)
    # fewer than 2048 rows: the reference raises RuntimeError (contrastive_loss.py:205)
    with pytest.raises(RuntimeError):
        ContrastiveLossPlus()(torch.randn(1024, 256, device=cuda), torch.zeros(1024, 1, dtype=torch.int64, device=cuda))
    # normalized_input=False
    x, labels = synth.make_roi_set(2088, seed=5)
    loss, grad, _ = _run(cuda, x * 0.3, labels, normalized_input=False)
    ref, gref = supcon_np.supcon_loss((x * 0.3).numpy(), labels.numpy(), 0.06, 10, 0.01, normalized_input=False, want_grad=True)
    assert abs(loss.item() - ref) <= RTOL * abs(ref)
    assert np.linalg.norm(grad.cpu().numpy() - gref) <= RTOL * np.linalg.norm(gref)


def test_property_invariances(cuda):
    """Size-independent properties at full size: row-scale invariance (double normalisation) and
    invariance of the loss under a permutation that respects the two-view pairing."""
    import torch
    x, labels = synth.make_roi_set(2088, seed=11)
    l0, _, _ = _run(cuda, x, labels)
    scale = torch.rand(2088, 1) * 5 + 0.1
    l1, _, _ = _run(cuda, x * scale, labels)
    assert abs(l0.item() - l1.item()) <= 1e-5 * abs(l0.item())
    perm = torch.cat([torch.randperm(1023), torch.tensor([1023])])   # the last RoI's label seeds the rp rows
    idx = torch.cat([perm, perm + 1024, torch.arange(2048, 2088)])
    l2, _, _ = _run(cuda, x[idx], labels[torch.cat([perm, perm + 1024])])
    assert abs(l0.item() - l2.item()) <= 1e-5 * abs(l0.item())


def test_cuda_core_path_agrees_with_tensor_core_path(cuda):
    """A test build of the library (-DOADG_LOSS_FFMA: CUDA-core similarity kernels, an independent implementation of
    the same closed form; the product has no such switch) must meet the same tolerance as the tcgen05 path."""
    import os
    import subprocess
    import sys
    from oadg_b200 import build
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ffma = build.build_test_variant(os.path.join(root, 'tests', 'hostsim', 'libOADG_ffma.so'), ['-DOADG_LOSS_FFMA'])
    code = r'''
import numpy as np, torch
from oracle import supcon_np, synth
from oadg_b200 import ContrastiveLossPlus
x, labels = synth.make_roi_set(2088)
xd = x.cuda().requires_grad_(True)
loss = ContrastiveLossPlus(loss_weight=0.01, num_views=2, temperature=0.06)(xd, labels.cuda()); loss.backward()
ref, gref = supcon_np.supcon_loss(x.numpy(), labels.numpy(), 0.06, 10, 0.01, want_grad=True)
assert abs(loss.item() - ref) <= 1e-5 * abs(ref), (loss.item(), ref)
assert np.linalg.norm(xd.grad.cpu().numpy() - gref) <= 1e-5 * np.linalg.norm(gref)
print("FFMA-OK", loss.item())
'''
    out = subprocess.run([sys.executable, '-c', code], cwd=root, capture_output=True, text=True,
                         env=dict(os.environ, PYTHONPATH=root, OADG_LIB=ffma))
    assert 'FFMA-OK' in out.stdout, out.stdout[-1000:] + out.stderr[-3000:]


def test_gathered_entry_points_match_oracle_single_rank(cuda, tmp_path):
    """The cross-rank entry points (normalize / forward_gathered / backward_gathered) at world size 1 must equal
    the local loss: same tolerance against the oracle."""
    import torch
    import torch.distributed as dist
    from oadg_b200.distributed import gathered_contrastive_loss, CudaBackend
    if not dist.is_initialized():
        dist.init_process_group('gloo', init_method='file://%s' % (tmp_path / 'pg'), rank=0, world_size=1)
    try:
        x, labels = synth.make_roi_set(2088, seed=21)
        xd = x.to(cuda).requires_grad_(True)
        be = CudaBackend()
        loss = gathered_contrastive_loss(xd, labels.to(cuda), temperature=0.06, loss_weight=0.01, backend=be)
        loss.backward()
        ref, gref = supcon_np.supcon_loss(x.numpy(), labels.numpy(), 0.06, 10, 0.01, want_grad=True)
        assert abs(loss.item() - ref) <= RTOL * abs(ref)
        assert np.linalg.norm(xd.grad.cpu().numpy() - gref) <= RTOL * np.linalg.norm(gref)
        assert be.launches >= 6
    finally:
        dist.destroy_process_group()


def _two_rank_worker(rank, world, port, q, exchange):
    import os
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    from oadg_b200.distributed import CudaBackend, gathered_contrastive_loss
    be = CudaBackend()
    out = []
    for step in range(3):            # three steps: both halves of the exchange buffer and the first reuse of one
        x, labels = synth.make_roi_set(2048, seed=40 + rank + 10 * step)
        xd = x.cuda().requires_grad_(True)
        loss = gathered_contrastive_loss(xd, labels.cuda(), temperature=0.06, loss_weight=0.01, backend=be,
                                         exchange=exchange)
        loss.backward()
        out.append((x.numpy(), labels.numpy().reshape(-1), loss.item(), xd.grad.cpu().numpy()))
    # the other exchange gives the same bits (same kernels on the same gathered rows)
    other = CudaBackend()
    x2 = torch.from_numpy(out[-1][0]).cuda().requires_grad_(True)
    loss2 = gathered_contrastive_loss(x2, torch.from_numpy(out[-1][1]).cuda(), temperature=0.06, loss_weight=0.01,
                                      backend=other, exchange='nccl' if exchange == 'peer' else 'peer')
    loss2.backward()
    same = loss2.item() == out[-1][2] and bool((x2.grad.cpu().numpy() == out[-1][3]).all())
    # a backward that comes after the next forward must refuse (the statistics of its step are gone)
    stale = gathered_contrastive_loss(x2, torch.from_numpy(out[-1][1]).cuda(), backend=be, exchange=exchange)
    gathered_contrastive_loss(x2, torch.from_numpy(out[-1][1]).cuda(), backend=be, exchange=exchange)
    try:
        stale.backward()
        refused = False
    except RuntimeError:
        refused = True
    # RoI counts vary from step to step (random proposals): other row counts reuse the mapped buffers
    varied = []
    if exchange == 'peer':
        for n_rows in (2088, 2052):
            x, labels = synth.make_roi_set(n_rows, seed=70 + rank + n_rows)
            xd = x.cuda().requires_grad_(True)
            loss = gathered_contrastive_loss(xd, labels.cuda(), temperature=0.06, loss_weight=0.01, backend=be,
                                             exchange=exchange)
            loss.backward()
            varied.append((n_rows, x.numpy(), labels.numpy().reshape(-1), loss.item(), xd.grad.cpu().numpy()))
        assert be._px.n == 4096 and be._px.seq == 7
    # ranks that disagree on the row count: the step's result is meaningless, and the NEXT call must raise
    mismatch = None
    if exchange == 'peer':
        from oadg_b200._lib import OADGError
        n_bad = 2048 + 2 * rank
        x, labels = synth.make_roi_set(n_bad, seed=5)
        gathered_contrastive_loss(x.cuda(), labels.cuda(), backend=be, exchange=exchange)
        torch.cuda.synchronize()
        try:
            gathered_contrastive_loss(x.cuda(), labels.cuda(), backend=be, exchange=exchange)
            mismatch = False
        except OADGError as e:
            mismatch = 'disagree' in str(e)
    q.put((rank, out, same, refused, be.launches, varied, mismatch))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('exchange', ['peer', 'nccl'])
def test_gathered_loss_two_gpus_nccl(cuda, exchange):
    """2 ranks: all-gathered contrast set vs the single-process oracle on the concatenated batch, with the rows
    exchanged by peer stores over NVLink (no collective) and by NCCL all-gathers."""
    import socket
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    from oadg_b200.distributed import gathered_pair_map
    from oadg_b200 import reference_pair_map
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_two_rank_worker, args=(r, 2, port, q, exchange)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=300) for _ in range(2)], key=lambda r: r[0])
    [p.join(60) for p in procs]
    pair_all = gathered_pair_map(reference_pair_map(2048), 2)
    for step in range(3):
        x_all = np.concatenate([r[1][step][0] for r in res])
        y_all = np.concatenate([r[1][step][1] for r in res])
        ref, gref = supcon_np.supcon_loss(x_all, y_all, 0.06, 10, 0.01, want_grad=True, pair=pair_all)
        for rank, out, same, refused, launches, _, _ in res:
            _, _, loss, grad = out[step]
            assert abs(loss - ref) <= RTOL * abs(ref), (step, rank)
            want = 2 * gref[rank * 2048:(rank + 1) * 2048]
            assert np.linalg.norm(grad - want) <= RTOL * np.linalg.norm(want), (step, rank)
    assert res[0][1][2][2] == res[1][1][2][2]            # every rank reports the same bits
    assert all(r[2] for r in res), 'peer and nccl exchanges disagree'
    assert all(r[3] for r in res), 'a stale backward was not refused'
    if exchange == 'peer':
        for k, n_rows in enumerate((2088, 2052)):
            x_all = np.concatenate([r[5][k][1] for r in res])
            y_all = np.concatenate([supcon_np.pad_labels(r[5][k][2], n_rows) for r in res])
            ref, gref = supcon_np.supcon_loss(x_all, y_all, 0.06, 10, 0.01, want_grad=True,
                                              pair=gathered_pair_map(reference_pair_map(n_rows), 2))
            for rank, *_rest in res:
                _, _, _, loss, grad = res[rank][5][k]
                assert abs(loss - ref) <= RTOL * abs(ref), (n_rows, rank)
                want = 2 * gref[rank * n_rows:(rank + 1) * n_rows]
                assert np.linalg.norm(grad - want) <= RTOL * np.linalg.norm(want), (n_rows, rank)
        assert all(r[6] is True for r in res), 'a row-count mismatch between the ranks went unnoticed'


@pytest.mark.parametrize('world', [4, 8])
def test_gathered_entry_points_at_multi_rank_scale_on_one_gpu(cuda, world):
    """The contrast-set sizes the 4- and 8-GPU runs produce (n_total = W * 2088), without NCCL: one GPU plays every
    rank in turn through the same C-ABI calls `_GatheredSupCon` makes.  Sum of the W loss parts and each rank's
    gradient block against the oracle on the concatenated batch."""
    import torch
    from oadg_b200 import reference_pair_map
    from oadg_b200.distributed import CudaBackend, gathered_pair_map
    n, t, lw = 2088, 0.06, 0.01
    xs, ys = [], []
    for r in range(world):
        x, lab = synth.make_roi_set(n, seed=60 + r)
        lab = lab.view(-1)
        xs.append(x)
        ys.append(torch.cat([lab, lab[-1:].repeat(n - lab.shape[0])]))
    pair_np = gathered_pair_map(reference_pair_map(n), world)
    x_all, y_all = torch.cat(xs).numpy(), torch.cat(ys).numpy()
    ref, gref = supcon_np.supcon_loss(x_all, y_all, t, 10, lw, want_grad=True, pair=pair_np)
    be = CudaBackend()
    xd = [x.to(cuda) for x in xs]
    f_all = torch.cat([be.normalize(x, world * n, True) for x in xd]).contiguous()
    labels_all = torch.cat(ys).to(cuda)
    pair_all = torch.from_numpy(pair_np).to(cuda)
    parts, stats = zip(*[be.forward(f_all, labels_all, pair_all, r * n, n, t, lw, 10) for r in range(world)])
    loss = float(torch.stack(parts).double().sum())
    assert abs(loss - ref) <= RTOL * abs(ref)
    stats_all = torch.cat(stats).contiguous()
    one = torch.ones((), dtype=torch.float32, device=cuda)
    for r in (0, world // 2, world - 1):
        # the workspace carries one rank's normalize -> forward -> backward sequence (include/oadg.h): replay it
        be.normalize(xd[r], world * n, True)
        be.forward(f_all, labels_all, pair_all, r * n, n, t, lw, 10)
        gx = be.backward(xd[r], f_all, labels_all, pair_all, stats_all, r * n, t, True, one).cpu().numpy()
        want = gref[r * n:(r + 1) * n]
        assert np.linalg.norm(gx - want) <= RTOL * np.linalg.norm(want), r

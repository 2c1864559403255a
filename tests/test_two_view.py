"""SURVEY.md 8f rows f1 / f2 / f4 on CPU: the two-view plumbing that feeds OA-Loss its row order, the JSD /
first-view losses against goldens made by the reference (scripts/make_golden_f2.py) and, in the dev container,
against the reference modules themselves."""
import os

import numpy as np
import pytest
import torch

from oadg_b200 import consistency_losses as CL
from oadg_b200 import two_view as TV
from oadg_b200.contrastive_loss import reference_pair_map, yolo_pair_map
from oadg_b200.registry import LOSSES, build_loss

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, 'golden', 'consistency.npz'))


from conftest import CE_CFG, f2_inputs as _inputs  # noqa: E402


# ------------------------------------------------------------------------------------------------- f2: losses
@pytest.mark.parametrize('kind', ['roi', 'rpn'])
@pytest.mark.parametrize('prec', ['f64', 'f32'])
def test_cross_entropy_loss_plus_matches_reference_golden(kind, prec):
    pred, label, weight, avg = _inputs(kind)
    dt = torch.float64 if prec == 'f64' else torch.float32
    p = pred.detach().clone().to(dt).requires_grad_(True)
    loss = build_loss(dict(CE_CFG[kind]))(p, label, weight.to(dt), avg_factor=avg)
    loss.backward()
    tol = 1e-12 if prec == 'f64' else 2e-6
    assert abs(loss.item() - float(GOLD['%s/%s/loss' % (kind, prec)])) <= tol * max(1.0, abs(loss.item()))
    np.testing.assert_allclose(p.grad.numpy(), GOLD['%s/%s/grad' % (kind, prec)], rtol=tol * 10, atol=tol)


@pytest.mark.parametrize('kind', ['roi', 'rpn'])
def test_jsd_two_views_matches_reference_golden(kind):
    pred, label, _, _ = _inputs(kind)
    p = pred.detach().clone().requires_grad_(True)
    j = CL.jsdv1_3_2aug(p, label, None, reduction='mean', avg_factor=None)
    j.backward()
    assert abs(j.item() - float(GOLD[kind + '/jsd'])) <= 1e-12
    np.testing.assert_allclose(p.grad.numpy(), GOLD[kind + '/jsd_grad'], rtol=1e-10, atol=1e-14)
    # a symmetric divergence: identical views give zero, swapping the views changes nothing
    same = torch.cat([pred[:8], pred[:8]])
    assert CL.jsd_two_views_torch(same).item() == pytest.approx(0.0, abs=1e-12)
    half = pred.shape[0] // 2
    swapped = torch.cat([pred[half:], pred[:half]])
    assert CL.jsd_two_views_torch(swapped).item() == pytest.approx(CL.jsd_two_views_torch(pred).item(), rel=1e-12)


@pytest.mark.parametrize('name,typ,kw', [('smoothl1', 'SmoothL1LossPlus', dict(beta=1.0)), ('l1', 'L1LossPlus', {})])
def test_first_view_regression_losses_match_reference_golden(name, typ, kw):
    pred, tgt, w, avg = _inputs('reg')
    p = pred.detach().clone().requires_grad_(True)
    mod = build_loss(dict(type=typ, loss_weight=1.0, num_views=2, additional_loss='None', lambda_weight=0.0,
                          wandb_name='x', **kw))
    loss = mod(p, tgt, w, avg_factor=avg)
    loss.backward()
    assert abs(loss.item() - float(GOLD[name + '/loss'])) <= 1e-12
    np.testing.assert_allclose(p.grad.numpy(), GOLD[name + '/grad'], rtol=1e-12, atol=0)
    # only the first view's rows carry gradient
    assert float(p.grad[pred.shape[0] // 2:].abs().sum()) == 0.0


def test_losses_are_registered_under_the_reference_names():
    for name in ('CrossEntropyLossPlus', 'SmoothL1LossPlus', 'L1LossPlus', 'ContrastiveLossPlus'):
        assert LOSSES.get(name) is not None


def test_unsupported_variants_fail_loudly():
    with pytest.raises(NotImplementedError):
        build_loss(dict(type='CrossEntropyLossPlus', use_mask=True))
    with pytest.raises(NotImplementedError):
        build_loss(dict(type='SmoothL1LossPlus', additional_loss='jsd'))


@pytest.mark.skipif(not os.path.isdir('/root/reference'), reason='reference checkout only exists in the dev container')
def test_consistency_losses_match_live_reference_on_random_shapes():
    """In a subprocess: ref_loader installs stub mmcv/mmdet modules and must not meet the shim."""
    import subprocess
    import sys
    code = r'''
import torch
from oracle import ref_loader
R = ref_loader.load_reference()
from oadg_b200 import consistency_losses as CL
for seed in range(6):
    g = torch.Generator().manual_seed(100 + seed)
    n = int(torch.randint(2, 200, (1,), generator=g)) * 2
    c = int(torch.randint(2, 12, (1,), generator=g))
    pred = torch.randn(n, c, generator=g, dtype=torch.float64) * 5
    half = torch.randint(0, c, (n // 2,), generator=g)
    label, weight = torch.cat([half, half]), torch.rand(n, generator=g, dtype=torch.float64)
    for reduce_w in (False, True):
        kw = dict(use_sigmoid=False, num_views=2, additional_loss='jsdv1_3_2aug', lambda_weight=3.0, wandb_name='t',
                  additional_loss_weight_reduce=reduce_w, reduction='none' if reduce_w else 'mean')
        mine = CL.CrossEntropyLossPlus(**kw)(pred, label, weight, avg_factor=None if reduce_w else float(n))
        ref = R['CrossEntropyLossPlus'](**kw)(pred, label, weight, avg_factor=None if reduce_w else float(n))
        assert mine.shape == ref.shape and torch.allclose(mine, ref, rtol=1e-12, atol=0), (seed, reduce_w)
    logit = torch.randn(n, 1, generator=g, dtype=torch.float64) * 5
    lab = torch.cat([half.clamp(max=1)] * 2)
    kw = dict(use_sigmoid=True, num_views=2, additional_loss='jsdv1_3_2aug', lambda_weight=0.1, wandb_name='t')
    mine = CL.CrossEntropyLossPlus(**kw)(logit, lab, weight, avg_factor=37.0)
    ref = R['CrossEntropyLossPlus'](**kw)(logit, lab, weight, avg_factor=37.0)
    assert torch.allclose(mine, ref, rtol=1e-12, atol=0), seed
    tgt = torch.randn(n, c, generator=g, dtype=torch.float64)
    w4 = torch.rand(n, c, generator=g, dtype=torch.float64)
    for name, kw in (('SmoothL1LossPlus', dict(beta=0.5)), ('L1LossPlus', {})):
        kw = dict(kw, num_views=2, additional_loss='None', lambda_weight=0.0, wandb_name='t', loss_weight=2.0)
        mine = getattr(CL, name)(**kw)(pred, tgt, w4, avg_factor=11.0)
        ref = R[name](**kw)(pred, tgt, w4, avg_factor=11.0)
        assert torch.allclose(mine, ref, rtol=1e-12, atol=0), (seed, name)
from oadg_b200.two_view import TwoViewRPNLoss
g = torch.Generator().manual_seed(7)
n = 4 * 600                                              # 2 images x 2 views x 600 anchors
obj = torch.randn(n, 1, generator=g, dtype=torch.float64) * 3
dl = torch.randn(n, 4, generator=g, dtype=torch.float64)
tv = torch.randint(-1, 2, (n,), generator=g)             # torchvision labels: 1 fg, 0 bg, -1 neither
sampled = (torch.rand(n, generator=g) < 0.2) & (tv >= 0)
tgt = torch.randn(n, 4, generator=g, dtype=torch.float64)
mine = TwoViewRPNLoss()(obj, dl, tv, sampled, tgt)
ns = float(sampled.sum())
mm = torch.where(tv > 0, torch.zeros_like(tv), torch.ones_like(tv))
ref_cls = R['CrossEntropyLossPlus'](use_sigmoid=True, loss_weight=1.0, num_views=2, additional_loss='jsdv1_3_2aug',
                                    lambda_weight=0.1, wandb_name='rpn_cls')(obj, mm, sampled.double(), avg_factor=ns)
posw = (sampled & (tv > 0)).double().view(-1, 1).expand(-1, 4)
ref_box = R['L1LossPlus'](loss_weight=1.0, num_views=2, additional_loss='None', lambda_weight=0.0,
                          wandb_name='rpn_bbox')(dl, tgt, posw, avg_factor=ns)
assert torch.allclose(mine['loss_rpn_cls'], ref_cls, rtol=1e-12, atol=0)
assert torch.allclose(mine['loss_rpn_bbox'], ref_box, rtol=1e-12, atol=0)
print("LIVE-OK")
'''
    root = os.path.dirname(HERE)
    out = subprocess.run([sys.executable, '-c', code], cwd=root, capture_output=True, text=True,
                         env=dict(os.environ, PYTHONPATH=root))
    assert 'LIVE-OK' in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


# ------------------------------------------------------------------------------------------------- f4: yolo layout
def test_yolo_pair_map_is_the_half_split_layout():
    p = yolo_pair_map(1800)
    assert (p[:900] == np.arange(900) + 900).all() and (p[900:] == np.arange(900)).all()
    p = yolo_pair_map(1801)          # contrastive_loss.py:257-259: rp_total_size 1 -> rp_size 0, the odd row is unpaired
    assert p[1800] == -1 and (p[:900] == np.arange(900) + 900).all()
    # the RoI layout keeps its own rule (1024 rows per view, random proposals split in two)
    q = reference_pair_map(2088)
    assert q[0] == 1024 and q[2048] == 2068 and q[2068] == 2048


def test_yolo_oracle_matches_reference_golden():
    from oracle import supcon_np, synth
    for n in (1800, 1801):
        x, _ = synth.make_roi_set(max(n, 2048), seed=n)
        x = torch.nn.functional.normalize(x[:n].double(), dim=1).numpy()
        labels = GOLD['yolo%d/labels' % n].reshape(-1)
        loss, grad = supcon_np.supcon_loss(x, labels, temperature=0.06, min_samples=10, loss_weight=1.0,
                                           normalized_input=False, want_grad=True, pair=yolo_pair_map(n))
        assert abs(loss - float(GOLD['yolo%d/loss' % n])) <= 1e-10
        rows = np.concatenate([grad[:8], grad[n // 2:n // 2 + 8]])
        np.testing.assert_allclose(rows, GOLD['yolo%d/grad_rows' % n], rtol=1e-7, atol=1e-12)


# ------------------------------------------------------------------------------------------------- f1: plumbing
def _batch(b=2, h=64, w=96, views=2, per_view_lists=False):
    g = torch.Generator().manual_seed(3)
    data = {'img': torch.randn(b, 3, h, w, generator=g), 'img_metas': [dict(idx=i) for i in range(b)]}
    data['gt_bboxes'] = [torch.tensor([[4., 4., 40., 40.], [50., 10., 90., 60.]]) + i for i in range(b)]
    data['gt_labels'] = [torch.tensor([1, 3]) for _ in range(b)]
    for v in range(2, views + 1):
        data['img%d' % v] = torch.randn(b, 3, h, w, generator=g)
        if per_view_lists:
            data['gt_bboxes%d' % v] = [x + 100 for x in data['gt_bboxes']]
            data['gt_labels%d' % v] = [x + 1 for x in data['gt_labels']]
    return data


def test_integrate_data_stacks_views_and_replicates_lists():
    data = _batch()
    img1, img2 = data['img'].clone(), data['img2'].clone()
    out = TV.integrate_data(data, {})
    assert out['num_views'] == 2 and out['batch_size'] == 2 and 'img2' not in out
    assert torch.equal(out['img'], torch.cat([img1, img2]))
    assert len(out['gt_bboxes']) == 4 and out['gt_bboxes'][2] is out['gt_bboxes'][0]
    assert [m['idx'] for m in out['img_metas']] == [0, 1, 0, 1]
    # per-view lists are appended instead of replicated
    data = _batch(per_view_lists=True)
    out = TV.integrate_data(data, {})
    assert float(out['gt_bboxes'][2][0, 0]) == 104.0 and 'gt_bboxes2' not in out and 'gt_labels2' not in out
    assert out['gt_labels'][3].tolist() == [2, 4]
    # 'inv' puts the augmented view first (base.py:25-27)
    data = _batch()
    out = TV.integrate_data(data, {'inv': True})
    assert torch.equal(out['img'][:2], img2)


def test_replicated_sampling_gives_every_view_the_same_rows():
    g = torch.Generator().manual_seed(0)
    gt = [torch.tensor([[10., 10., 60., 60.]]), torch.tensor([[20., 20., 90., 80.], [5., 5., 30., 30.]])]
    lab = [torch.tensor([2]), torch.tensor([0, 5])]
    props = [torch.rand(300, 4, generator=g) * 50 for _ in range(4)]
    for p in props:
        p[:, 2:] += p[:, :2] + 4
    res = TV.replicate_sampling(props, gt * 2, lab * 2, batch_size=2, num_views=2, num=64, generator=g)
    assert len(res) == 4 and res[2] is res[0] and res[3] is res[1]
    for r in res[:2]:
        assert r.bboxes.shape[0] == 64 and r.pos_bboxes.shape[0] <= 16
        assert torch.equal(r.bboxes[:r.pos_bboxes.shape[0]], r.pos_bboxes)       # positives first
        assert (TV.box_iou(r.pos_bboxes, r.pos_gt_bboxes).diagonal() >= 0.5).all()
    rois = TV.bbox2roi([r.bboxes for r in res])
    assert rois.shape == (256, 5) and rois[:, 0].tolist() == sum([[float(i)] * 64 for i in range(4)], [])
    assert torch.equal(rois[:64, 1:], rois[128:192, 1:])                         # row k of view 1 == row k of view 2


def test_random_proposals_respect_the_iou_band_and_the_count():
    g = torch.Generator().manual_seed(5)
    gt = [torch.tensor([[100., 100., 300., 300.]]), torch.tensor([[50., 60., 400., 500.]])] * 2
    oamix = [torch.tensor([[100., 100., 300., 300.], [600., 600., 700., 700.]])] * 4
    out = TV.random_proposals((512, 1024), gt, 2, multilevel_boxes=oamix, oamix_boxes=None, num_bboxes=10,
                              generator=g)
    assert len(out) == 4
    for i, b in enumerate(out):
        assert b.shape[1] == 4 and 1 <= b.shape[0] <= 11
        assert torch.equal(b[0], oamix[0][1])           # the box on top of gt[0] was filtered, the far one kept
        fresh = b[1:]
        assert (fresh[:, 2] <= 1024).all() and (fresh[:, 3] <= 512).all() and (fresh[:, 2] > fresh[:, 0]).all()
        assert (TV.box_iou(fresh, gt[i % 2]).max(dim=1)[0] <= 0.7).all()


class _CountingLoss(torch.nn.Module):
    min_samples = 10

    def __init__(self):
        super().__init__()
        self.calls = []

    def forward(self, feats, labels):
        self.calls.append((feats.shape[0], labels.shape[0]))
        return feats.sum() * 0 + 1.0


def test_head_gates_the_contrastive_term_and_always_emits_it():
    torch.manual_seed(0)
    head = TV.Shared2FCContrastiveHead(in_channels=4, roi_feat_size=2, fc_out_channels=32, num_classes=8, out_dim_cont=16)
    head.loss_cont = _CountingLoss()
    n = 64
    cls, reg, cont = head(torch.randn(n + 6, 4, 2, 2))
    assert cls.shape == (n + 6, 9) and reg.shape == (n + 6, 32) and cont.shape == (n + 6, 16)
    labels = torch.full((n,), 8)
    labels[:12] = torch.arange(12) % 8
    w, t, tw = torch.ones(n), torch.zeros(n, 4), torch.zeros(n, 4)
    out = head.loss(cls[:n], reg[:n], cont, labels, w, t, tw)
    assert head.loss_cont.calls == [(n + 6, n)] and float(out['loss_cont']) == 1.0
    labels[10:] = 8                                   # 10 foreground rows: not MORE than min_samples -> gated off
    out = head.loss(cls[:n], reg[:n], cont, labels, w, t, tw)
    assert len(head.loss_cont.calls) == 1
    assert float(out['loss_cont']) == 0.0 and out['loss_cont'].requires_grad      # still on the graph, for DDP
    out['loss_cont'].backward()
    assert head.fc_cont[0].weight.grad is not None


def test_two_view_roi_head_row_order_on_cpu():
    """[v1 img0, v1 img1, v2 img0, v2 img1, rp...]: 2 * B * num rows, then the random-proposal rows."""
    torch.manual_seed(1)
    g = torch.Generator().manual_seed(1)
    head = TV.TwoViewRoIHead(num_classes=8, featmap_names=('0',), num=32)
    seen = {}

    class Spy(torch.nn.Module):
        min_samples = 0

        def forward(self, feats, labels):
            seen['rows'], seen['labels'] = feats.shape[0], labels.view(-1).clone()
            return feats.sum() * 0

    head.bbox_head.loss_cont = Spy()
    feats = {'0': torch.randn(4, 256, 16, 24)}
    shapes = [(64, 96)] * 4
    gt = [torch.tensor([[4., 4., 40., 40.]]), torch.tensor([[50., 10., 90., 60.]])] * 2
    lab = [torch.tensor([1]), torch.tensor([3])] * 2
    props = [torch.tensor([[0., 0., 30., 30.], [5., 5., 41., 41.], [48., 9., 88., 58.], [60., 30., 80., 50.]])] * 4
    rp = [torch.tensor([[1., 1., 9., 9.], [2., 2., 12., 12.]])] * 4
    out = head.forward_train(feats, shapes, props, gt, lab, 2, 2, random_proposal_list=rp, generator=g)
    per = head.last_rois.shape[0] // 4
    assert seen['rows'] == 4 * per + 8 and seen['labels'].shape[0] == 4 * per
    assert torch.equal(head.last_rois[:2 * per, 1:], head.last_rois[2 * per:, 1:])
    assert torch.equal(seen['labels'][:2 * per], seen['labels'][2 * per:])
    assert set(out) == {'loss_cls', 'loss_bbox', 'loss_cont'}


def test_rpn_losses_take_the_first_view_and_the_jsd_of_all_anchors():
    """TwoViewRPNLoss against the formula written out: BCE on the sampled anchors of view 1 and L1 on its sampled
    positives, both over the number of anchors sampled in ALL views, plus 0.1 x the JSD (sigmoid, 1 - sigmoid) between
    the views over every anchor, over the same count."""
    g = torch.Generator().manual_seed(3)
    n = 2 * 500
    obj = torch.randn(n, 1, generator=g, dtype=torch.float64) * 2
    dl, tgt = torch.randn(n, 4, generator=g, dtype=torch.float64), torch.randn(n, 4, generator=g, dtype=torch.float64)
    tv = torch.randint(-1, 2, (n,), generator=g)
    sampled = (torch.rand(n, generator=g) < 0.3) & (tv >= 0)
    out = TV.TwoViewRPNLoss()(obj, dl, tv, sampled, tgt)
    ns = float(sampled.sum())
    h = n // 2
    bce = torch.nn.functional.binary_cross_entropy_with_logits(obj[:h, 0], (tv[:h] > 0).double(), reduction='none')
    want_cls = (bce * sampled[:h]).sum() / ns + 0.1 * CL.jsd_two_views_torch(obj) / ns
    pos = (sampled & (tv > 0))[:h]
    want_box = (torch.abs(dl[:h] - tgt[:h]) * pos.view(-1, 1)).sum() / ns
    # (the reference feeds float32 labels to binary_cross_entropy_with_logits, so its float64 result carries ~1e-8)
    assert out['loss_rpn_cls'].item() == pytest.approx(want_cls.item(), rel=1e-7)
    assert out['loss_rpn_bbox'].item() == pytest.approx(want_box.item(), rel=1e-12)


def test_rpn_forward_oadg_on_a_small_torchvision_rpn():
    """torchvision's head / anchors / assignment / sampler / proposal filter with the OA-DG losses on top: proposals
    for every image of the integrated batch, finite losses, gradients reach the RPN head from view 1's sampled anchors
    and (through the JSD term) from both views' logits."""
    from torchvision.models.detection.anchor_utils import AnchorGenerator
    from torchvision.models.detection.image_list import ImageList
    from torchvision.models.detection.rpn import RegionProposalNetwork, RPNHead
    torch.manual_seed(0)
    anchors = AnchorGenerator(sizes=((32,), (64,)), aspect_ratios=((0.5, 1.0, 2.0),) * 2)
    rpn = RegionProposalNetwork(anchors, RPNHead(16, 3), 0.7, 0.3, 64, 0.5, dict(training=200, testing=100),
                                dict(training=50, testing=50), 0.7).train()
    feats = {'0': torch.randn(4, 16, 16, 32), '1': torch.randn(4, 16, 8, 16)}
    images = ImageList(torch.randn(4, 3, 128, 256), [(128, 256)] * 4)
    gt = [torch.tensor([[10., 10., 100., 100.]]), torch.tensor([[30., 20., 200., 120.], [5., 60., 60., 120.]])] * 2
    targets = [dict(boxes=b, labels=torch.ones(len(b), dtype=torch.long)) for b in gt]
    boxes, losses = TV.rpn_forward_oadg(rpn, images, feats, targets, TV.TwoViewRPNLoss())
    assert len(boxes) == 4 and all(b.shape[1] == 4 and 0 < b.shape[0] <= 50 for b in boxes)
    total = losses['loss_rpn_cls'] + losses['loss_rpn_bbox']
    assert torch.isfinite(total)
    total.backward()
    assert rpn.head.cls_logits.weight.grad.abs().sum() > 0 and rpn.head.bbox_pred.weight.grad.abs().sum() > 0


def test_dc5_architecture_of_config_5():
    """R-DC5 (configs/_base_/models/faster_rcnn_r50_caffe_dc5.py): one stride-16 map of 2048 channels, 15 anchors per
    location, a 2048-channel RoI head with the 7 DWD classes."""
    torch.manual_seed(0)
    m = TV.TwoViewFasterRCNN(num_classes=7, backbone='resnet50', arch='dc5')
    feats = m.backbone(torch.randn(1, 3, 96, 160))
    assert list(feats) == ['0'] and feats['0'].shape == (1, 2048, 6, 10)
    assert m.rpn.head.cls_logits.out_channels == 15 and m.rpn.head.bbox_pred.out_channels == 60
    head = m.roi_head.bbox_head
    assert head.shared_fcs[0].in_features == 2048 * 7 * 7 and head.fc_cls.out_features == 8 and head.fc_reg.out_features == 28
    with pytest.raises(ValueError):
        TV.TwoViewFasterRCNN(arch='c4')


def test_fixed_count_pads_and_truncates_random_proposals():
    """Under the gathered loss every rank must bring the same number of rows: each image's random proposals are cut or
    padded (by repeating the last box) to a fixed count."""
    b = torch.arange(12.).view(3, 4)
    assert torch.equal(TV.fixed_count(b, 2), b[:2])
    out = TV.fixed_count(b, 5)
    assert out.shape == (5, 4) and torch.equal(out[:3], b) and torch.equal(out[3:], b[-1:].expand(2, -1))
    assert TV.fixed_count(b[:0], 3).shape == (3, 4)
    head = TV.TwoViewRoIHead(num_classes=8, featmap_names=('0',), gather=True)
    assert head.rp_per_image == 10 and head.bbox_head.gather
    assert TV.TwoViewRoIHead(num_classes=8, featmap_names=('0',)).rp_per_image is None

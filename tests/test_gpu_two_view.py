"""GPU parity for SURVEY.md 8f rows f1 / f2 / f4: the fused JSD kernel, supcontrast_yolo and the two-view RoI step."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, 'golden', 'consistency.npz'))


@pytest.mark.parametrize('n,c,scale', [(128, 9, 3.0), (2048, 1, 4.0), (5, 2, 1.0), (70000, 1, 6.0), (1024, 32, 20.0)])
def test_jsd_kernel_matches_torch_form(n, c, scale):
    """value and gradient of the fused kernel vs the reference's op sequence in float64 (tolerance: float32 round-off
    of ~c exp/log per row, 2e-6 relative on the sum, 2e-6 absolute on the per-logit gradient)."""
    from oadg_b200 import consistency_losses as CL
    g = torch.Generator().manual_seed(n + c)
    pred = (torch.randn(2 * n, c, generator=g) * scale)
    pred[0, :] = 40.0 if c == 1 else pred[0, :]            # saturated sigmoid: 1 - s == 0 exactly in float32
    x = pred.cuda().requires_grad_(True)
    loss = CL.jsd_two_views(x)
    (loss * 0.5).backward()
    ref_in = pred.double().requires_grad_(True)
    # the float64 form of the same float32 probabilities is what the kernel approximates; saturation differs in f64,
    # so the saturated row is compared in float32 only
    ref = CL.jsd_two_views_torch(ref_in)
    (ref * 0.5).backward()
    assert loss.item() == pytest.approx(ref.item(), rel=2e-6, abs=1e-6)
    gk, gr = x.grad.cpu().double(), ref_in.grad
    if c == 1:
        gk, gr = torch.cat([gk[1:n], gk[n + 1:]]), torch.cat([gr[1:n], gr[n + 1:]])
    assert float((gk - gr).abs().max()) <= 2e-6
    assert torch.isfinite(x.grad).all()


def test_jsd_kernel_matches_reference_golden_and_is_deterministic():
    from oadg_b200 import consistency_losses as CL
    import conftest as T
    from oadg_b200.registry import build_loss
    for kind in ('roi', 'rpn'):
        pred, label, weight, avg = T.f2_inputs(kind)
        x = pred.float().cuda().requires_grad_(True)
        j = CL.jsdv1_3_2aug(x, label.cuda(), None)
        j.backward()
        assert j.item() == pytest.approx(float(GOLD[kind + '/jsd']), rel=2e-6)
        np.testing.assert_allclose(x.grad.cpu().numpy(), GOLD[kind + '/jsd_grad'], rtol=0, atol=3e-6)
        again = CL.jsd_two_views(x.detach())
        assert again.item() == j.item()
        # the whole module, as the head calls it
        mod = build_loss(dict(T.CE_CFG[kind])).cuda()
        y = pred.float().cuda().requires_grad_(True)
        loss = mod(y, label.cuda(), weight.float().cuda(), avg_factor=avg)
        loss.backward()
        assert loss.item() == pytest.approx(float(GOLD['%s/f32/loss' % kind]), rel=3e-6)
        np.testing.assert_allclose(y.grad.cpu().numpy(), GOLD['%s/f32/grad' % kind], rtol=0, atol=3e-6)


@pytest.mark.parametrize('n', [1800, 1801])
def test_supcontrast_yolo_matches_reference_golden(n):
    from oadg_b200.contrastive_loss import supcontrast_yolo
    from oracle import synth
    x, _ = synth.make_roi_set(max(n, 2048), seed=n)
    x = torch.nn.functional.normalize(x[:n].double(), dim=1).float().cuda().requires_grad_(True)
    labels = torch.from_numpy(GOLD['yolo%d/labels' % n]).cuda()
    loss = supcontrast_yolo(x, labels, temper=0.06, min_samples=10)
    loss.backward()
    # 3xTF32 similarity + float32 row sums: the tolerance the OA-Loss parity tests use
    assert loss.item() == pytest.approx(float(GOLD['yolo%d/loss' % n]), rel=2e-5)
    rows = torch.cat([x.grad[:8], x.grad[n // 2:n // 2 + 8]]).cpu().numpy()
    np.testing.assert_allclose(rows, GOLD['yolo%d/grad_rows' % n], rtol=2e-3, atol=2e-7)


def test_two_view_step_runs_on_the_gpu_and_feeds_the_loss_its_row_order():
    """Config-3 shape: 2 images x 2 views, 512 RoIs per image + random proposals -> loss_cont from the CUDA path."""
    from oadg_b200 import two_view as TV
    torch.manual_seed(0)
    dev = torch.device('cuda:0')
    g = torch.Generator(device=dev).manual_seed(0)
    head = TV.TwoViewRoIHead(num_classes=8, num=512).to(dev)
    feats = {str(i): torch.randn(4, 256, 128 >> i, 256 >> i, device=dev) for i in range(4)}
    shapes = [(512, 1024)] * 4
    gt1 = [torch.tensor([[40. + 30 * k, 60. + 20 * k, 140. + 40 * k, 200. + 22 * k] for k in range(12)], device=dev),
           torch.tensor([[500. - 30 * k, 90. + 25 * k, 640. - 28 * k, 180. + 30 * k] for k in range(9)], device=dev)]
    lab1 = [torch.arange(12, device=dev) % 8, torch.arange(9, device=dev) % 8]
    props = []
    for i in range(4):
        base = gt1[i % 2].repeat(40, 1) + torch.randn(gt1[i % 2].shape[0] * 40, 4, device=dev, generator=g) * 6
        far = torch.rand(600, 4, device=dev, generator=g) * 300
        far[:, 2:] += far[:, :2] + 8
        props.append(torch.cat([base, far]))
    rp = TV.random_proposals((512, 1024), gt1 * 2, 2, multilevel_boxes=[gt1[0][:3] + 300.0] * 4, generator=g)
    out = head.forward_train(feats, shapes, props, gt1 * 2, lab1 * 2, 2, 2, random_proposal_list=rp, generator=g)
    assert head.last_rois.shape[0] == 2048
    assert torch.equal(head.last_rois[:1024, 1:], head.last_rois[1024:, 1:])
    assert head.bbox_head.loss_cont.stats.get('launches', 0) > 0          # the CUDA OA-Loss ran
    total = sum(out.values())
    total.backward()
    assert torch.isfinite(total) and float(out['loss_cont']) > 0
    assert head.bbox_head.fc_cont[2].weight.grad.abs().sum() > 0

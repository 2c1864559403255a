"""Two-view training-step plumbing around the OA-Loss head: what hands
``ContrastiveLossPlus`` its rows in the order it assumes.

The detector itself (backbone, FPN, RPN, RoIAlign) stays stock torch / torchvision (BASELINE north_star); this module
restates the OA-DG-specific glue of the reference:

* ``integrate_data``            detectors/base.py:22-48       views concatenated along the batch, per-view lists replicated
* ``replicate_sampling``        roi_heads/contrastive_roi_head.py:84-97   RoIs are sampled on view 1 only and reused for every
                                view, so row k of view 1 and row k of view 2 are the same RoI
* ``random_proposals``          detectors/two_stage.py:162-204, bbox_augmentation.py (generate_random_bboxes_xy)
                                OA-Mix boxes (IoU-filtered against gt) + fresh random boxes per image; drawn on the DEVICE
                                here (no numpy round trip), same acceptance rule
* ``Shared2FCContrastiveHead``  roi_heads/bbox_heads/contrastive_head.py:141-366 two shared FCs, cls / reg branches and
                                ``fc_cont`` = Linear(1024, 256) + ReLU + Linear(256, 256); ``loss`` gates the contrastive
                                term on the number of foreground RoIs (:125-129) but -- unlike the reference -- always emits
                                ``loss_cont`` (a zero attached to the graph) so that DDP sees the same parameters every step
* ``TwoViewRPNLoss`` / ``rpn_forward_oadg``  dense_heads/anchor_head.py:402-547 with the configs' RPN losses (first-view
                                BCE + 0.1 x JSD over all anchors, first-view L1) on torchvision's RPN
* ``TwoViewRoIHead`` / ``TwoViewFasterRCNN``  the step: RoI row order [v1 img0, v1 img1, v2 img0, v2 img1, rp ...], 512
                                rows per image, positives first (core/bbox/samplers/sampling_result.py:53-55)
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .consistency_losses import CrossEntropyLossPlus, L1LossPlus, SmoothL1LossPlus
from .contrastive_loss import ContrastiveLossPlus


# ---------------------------------------------------------------------------------------------------------------
def integrate_data(data, train_cfg=None):
    """detectors/base.py:22-48.  ``data``: the collated batch of a two-view pipeline: ``img`` [B,3,H,W], ``img2`` ...,
    and per-image lists (``gt_bboxes``, ``gt_labels``, ``img_metas``, ``multilevel_boxes``, ``oamix_boxes``), optionally
    with per-view versions ``key2``.  Returns the same dict with the views stacked along the batch (view-major:
    [v1 img0 .. v1 imgB-1, v2 img0 ..]) and every list extended to B * num_views entries."""
    train_cfg = train_cfg or {}
    batch_size = len(data['img'])
    if 'inv' in train_cfg:                                # (:25-27: a present-but-false 'inv' leaves a single view)
        if train_cfg['inv']:
            data['img'] = torch.cat([data['img2'], data['img']], dim=0)
    else:
        data['img'] = torch.cat([v for k, v in data.items() if ('img' in k) and ('img_metas' not in k)], dim=0)
    num_views = int(len(data['img']) / batch_size)
    for i in range(2, num_views + 1):
        for key in ('img', 'gt_bboxes', 'gt_labels', 'gt_instance_inds', 'img_metas', 'multilevel_boxes', 'oamix_boxes'):
            if f'{key}{i}' in data:
                if key != 'img':
                    data[key] = list(data[key]) + list(data[f'{key}{i}'])
                del data[f'{key}{i}']
            elif key in data and key != 'img':
                data[key] = list(data[key]) + [data[key][b] for b in range(batch_size)]
    data['num_views'] = num_views
    data['batch_size'] = batch_size
    return data


# ---------------------------------------------------------------------------------------------------------------
class SamplingResult:
    """The fields of mmdet's SamplingResult the head consumes; ``bboxes`` = positives then negatives
    (core/bbox/samplers/sampling_result.py:53-55)."""

    def __init__(self, pos_bboxes, neg_bboxes, pos_gt_bboxes, pos_gt_labels):
        self.pos_bboxes, self.neg_bboxes = pos_bboxes, neg_bboxes
        self.pos_gt_bboxes, self.pos_gt_labels = pos_gt_bboxes, pos_gt_labels

    @property
    def bboxes(self):
        return torch.cat([self.pos_bboxes, self.neg_bboxes], dim=0)


def box_iou(a, b):
    """IoU of [n,4] x [m,4] xyxy boxes (mmdet bbox_overlaps, mode='iou')."""
    if a.numel() == 0 or b.numel() == 0:
        return a.new_zeros((a.shape[0], b.shape[0]))
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    lt = torch.max(a[:, None, :2], b[None, :, :2])
    rb = torch.min(a[:, None, 2:], b[None, :, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    return inter / (area_a[:, None] + area_b[None, :] - inter).clamp(min=1e-6)


def assign_and_sample(proposals, gt_bboxes, gt_labels, num=512, pos_fraction=0.25, pos_iou=0.5, generator=None):
    """MaxIoUAssigner(0.5, 0.5, 0.5) + RandomSampler(num, pos_fraction, add_gt_as_proposals=True) of the OA-DG configs
    (configs/_base_/models/faster_rcnn_r50_fpn.py rcnn train_cfg): gt boxes join the proposals, a proposal is positive
    when its best IoU >= 0.5, up to num * pos_fraction positives and the rest negatives, each drawn at random."""
    boxes = torch.cat([gt_bboxes, proposals[:, :4]], dim=0)
    iou = box_iou(boxes, gt_bboxes)
    if gt_bboxes.shape[0]:
        best, arg = iou.max(dim=1)
    else:
        best, arg = boxes.new_zeros(boxes.shape[0]), boxes.new_zeros(boxes.shape[0], dtype=torch.long)
    pos = torch.nonzero(best >= pos_iou, as_tuple=False).view(-1)
    neg = torch.nonzero(best < pos_iou, as_tuple=False).view(-1)
    n_pos = min(int(num * pos_fraction), pos.numel())
    pos = pos[torch.randperm(pos.numel(), generator=generator, device=pos.device)[:n_pos]]
    n_neg = min(num - n_pos, neg.numel())
    neg = neg[torch.randperm(neg.numel(), generator=generator, device=neg.device)[:n_neg]]
    return SamplingResult(boxes[pos], boxes[neg], gt_bboxes[arg[pos]], gt_labels[arg[pos]])


def replicate_sampling(proposal_list, gt_bboxes, gt_labels, batch_size, num_views, **kw):
    """contrastive_roi_head.py:84-97: RoIs are assigned and sampled on the first ``batch_size`` images (view 1) and the
    very same sampling results serve every view."""
    per_img = [assign_and_sample(proposal_list[i], gt_bboxes[i], gt_labels[i], **kw) for i in range(batch_size)]
    out = []
    for _ in range(num_views):
        out.extend(per_img)
    return out


def bbox2roi(bbox_list):
    """[n_i, 4] boxes of image i -> [sum n_i, 5] rows (image index, x1, y1, x2, y2)."""
    rois = []
    for i, b in enumerate(bbox_list):
        idx = b.new_full((b.shape[0], 1), float(i))
        rois.append(torch.cat([idx, b[:, :4]], dim=1))
    return torch.cat(rois, dim=0)


# ---------------------------------------------------------------------------------------------------------------
def random_proposals(img_shape, gt_bboxes, num_views, multilevel_boxes=None, oamix_boxes=None, num_bboxes=10,
                     scales=(0.01, 0.3), ratios=(0.3, 1 / 0.3), iou_max=0.7, iou_min=0.0, max_iters=500, generator=None,
                     bbox_from='oagrb'):
    """Random-proposal boxes per image of the integrated batch (two_stage.py:162-204): the OA-Mix multi-level / object-aware
    boxes that overlap the FIRST image's gt less than ``iou_max`` (the reference filters every list entry against
    ``gt_bboxes[0]``, :178,186), plus up to ``num_bboxes`` fresh random boxes per image accepted when their best IoU with
    that image's gt lies in [iou_min, iou_max] (generate_random_bboxes_xy).  Everything stays on the device: the
    ``max_iters`` candidates are drawn at once and the first ``num_bboxes`` accepted ones kept, which is the sequential
    rule of the reference applied to a pre-drawn candidate stream.  (The reference passes (H, W) where (W, H) is
    expected, :164,193 -- here width and height mean what they say.)"""
    h, w = int(img_shape[0]), int(img_shape[1])
    n_img = len(gt_bboxes)
    dev = gt_bboxes[0].device
    out = []
    for i in range(n_img):
        parts = []
        for extra in (multilevel_boxes, oamix_boxes):
            if extra is not None and i < len(extra):
                b = torch.as_tensor(extra[i], device=dev).to(torch.float32).view(-1, 4)
                if b.numel() and gt_bboxes[0].numel():
                    b = b[box_iou(b, gt_bboxes[0]).max(dim=1)[0] < iou_max]
                parts.append(b)
        gt = gt_bboxes[i % num_views]                     # :193, the reference's own indexing
        u = torch.rand(max_iters, 4, device=dev, generator=generator)
        x1 = (u[:, 0] * w).floor()
        y1 = (u[:, 1] * h).floor()
        area = (scales[0] + (scales[1] - scales[0]) * u[:, 2]) * h * w
        ratio = ratios[0] + (ratios[1] - ratios[0]) * u[:, 3]
        bw, bh = (area / ratio).sqrt().floor(), (area * ratio).sqrt().floor()
        cand = torch.stack([x1, y1, torch.minimum(x1 + bw, x1.new_tensor(float(w))),
                            torch.minimum(y1 + bh, y1.new_tensor(float(h)))], dim=1)
        if gt.numel():
            best = box_iou(cand, gt).max(dim=1)[0]
            cand = cand[(best <= iou_max) & (best >= iou_min)]
        parts.append(cand[:num_bboxes])
        out.append(torch.cat(parts, dim=0) if parts else cand.new_zeros((0, 4)))
    return out


def _world_size():
    import torch.distributed as dist
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def fixed_count(boxes, k):
    """Exactly k boxes: the first k, or the list padded by repeating its last box (an all-zero box when it is empty).
    The gathered loss needs the same number of rows on every rank, and the number of random proposals that survive
    the IoU filters varies."""
    n = boxes.shape[0]
    if n >= k:
        return boxes[:k]
    pad = boxes[-1:].expand(k - n, -1) if n else boxes.new_zeros((k, 4))[:k - n]
    return torch.cat([boxes, pad], dim=0)


# ---------------------------------------------------------------------------------------------------------------
class TwoViewRPNLoss(nn.Module):
    """The RPN losses of the OA-DG configs (..._oadg.py:19-26) in the layout of ``AnchorHead.loss`` /
    ``loss_single`` (dense_heads/anchor_head.py:402-453,455-547): the objectness logits of every anchor of every image of
    the integrated batch (view-major), ``CrossEntropyLossPlus(use_sigmoid=True)`` = binary cross entropy on the SAMPLED
    anchors of the first view + ``lambda_weight`` (0.1) x the JSD between the two views' objectness over ALL anchors, and
    ``L1LossPlus`` on the sampled positives of the first view; both averaged by the number of anchors sampled over all
    images.  The reference evaluates the two losses level by level and adds the levels up; every term is a sum over
    anchors divided by the same ``avg_factor``, so one call over all levels gives the same number."""

    def __init__(self, loss_cls=None, loss_bbox=None):
        super().__init__()
        cfg = dict(use_sigmoid=True, loss_weight=1.0, num_views=2, additional_loss='jsdv1_3_2aug', lambda_weight=0.1,
                   wandb_name='rpn_cls')
        cfg.update({k: v for k, v in (loss_cls or {}).items() if k != 'type'})
        self.loss_cls = CrossEntropyLossPlus(**cfg)
        cfg = dict(loss_weight=1.0, num_views=2, additional_loss='None', lambda_weight=0.0, wandb_name='rpn_bbox')
        cfg.update({k: v for k, v in (loss_bbox or {}).items() if k != 'type'})
        self.loss_bbox = L1LossPlus(**cfg)

    def forward(self, objectness, pred_deltas, anchor_labels, sampled, reg_targets):
        """objectness [B * A, 1], pred_deltas [B * A, 4] (image-major, B = batch_size * num_views, view-major images);
        anchor_labels [B * A] in torchvision's convention (1 foreground, 0 background, -1 neither); sampled [B * A] bool
        (the anchors the sampler kept); reg_targets [B * A, 4]."""
        n_sampled = max(float(sampled.sum()), 1.0)
        # mmdet's RPN labels: class 0 = object, 1 (= num_classes) = background
        labels = torch.where(anchor_labels > 0, torch.zeros_like(anchor_labels), torch.ones_like(anchor_labels))
        weights = sampled.to(objectness.dtype)
        loss_cls = self.loss_cls(objectness, labels, weights, avg_factor=n_sampled)
        pos = (sampled & (anchor_labels > 0)).to(pred_deltas.dtype).view(-1, 1).expand(-1, 4)
        loss_bbox = self.loss_bbox(pred_deltas, reg_targets, pos, avg_factor=n_sampled)
        return dict(loss_rpn_cls=loss_cls, loss_rpn_bbox=loss_bbox)


def rpn_forward_oadg(rpn, images, features, targets, rpn_loss):
    """torchvision's ``RegionProposalNetwork.forward`` with the OA-DG losses in place of its own: same head, anchors,
    target assignment (IoU 0.7 / 0.3), sampler (256 per image, half positive) and proposal filtering."""
    from torchvision.models.detection.rpn import concat_box_prediction_layers
    feats = list(features.values())
    objectness, deltas = rpn.head(feats)
    anchors = rpn.anchor_generator(images, feats)
    num_images = len(anchors)
    per_level = [o[0].numel() for o in objectness]
    obj_flat, deltas_flat = concat_box_prediction_layers(objectness, deltas)
    proposals = rpn.box_coder.decode(deltas_flat.detach(), anchors).view(num_images, -1, 4)
    boxes, _ = rpn.filter_proposals(proposals, obj_flat.detach(), images.image_sizes, per_level)
    labels, matched = rpn.assign_targets_to_anchors(anchors, targets)
    reg_targets = rpn.box_coder.encode(matched, anchors)
    pos_masks, neg_masks = rpn.fg_bg_sampler(labels)
    sampled = torch.cat([(p | n).bool() for p, n in zip(pos_masks, neg_masks)])
    losses = rpn_loss(obj_flat, deltas_flat, torch.cat(labels).long(), sampled, torch.cat(reg_targets))
    return boxes, losses


# ---------------------------------------------------------------------------------------------------------------
class Shared2FCContrastiveHead(nn.Module):
    """contrastive_head.py:141-366 with the OA-DG config values (..._oadg.py:17-44): roi features [n, 256, 7, 7] ->
    two shared FCs (1024) -> ``fc_cls`` (num_classes + 1), ``fc_reg`` (4 * num_classes), ``fc_cont``."""

    def __init__(self, in_channels=256, roi_feat_size=7, fc_out_channels=1024, num_classes=8, out_dim_cont=256,
                 loss_cont=None, loss_cls=None, loss_bbox=None, target_stds=(0.1, 0.1, 0.2, 0.2), gather=False):
        super().__init__()
        self.num_classes = num_classes
        # gather=True (BASELINE config 4): under an initialised process group of several ranks the contrast set is the
        # embeddings of ALL ranks (oadg_b200.distributed); every rank must then bring the same number of rows
        self.gather = gather
        d = in_channels * roi_feat_size * roi_feat_size
        self.shared_fcs = nn.ModuleList([nn.Linear(d, fc_out_channels), nn.Linear(fc_out_channels, fc_out_channels)])
        self.fc_cls = nn.Linear(fc_out_channels, num_classes + 1)
        self.fc_reg = nn.Linear(fc_out_channels, 4 * num_classes)
        # _add_linear_relu(num_linear=2, feat_channels=out_dim_cont, return_relu=True): Linear, ReLU, Linear
        self.fc_cont = nn.Sequential(nn.Linear(fc_out_channels, out_dim_cont), nn.ReLU(inplace=True),
                                     nn.Linear(out_dim_cont, out_dim_cont))
        cfg = dict(loss_weight=0.01, num_views=2, temperature=0.06)
        cfg.update({k: v for k, v in (loss_cont or {}).items() if k != 'type'})
        self.loss_cont = ContrastiveLossPlus(**cfg)
        self.loss_cont.num_classes = num_classes          # contrastive_head.py:58
        # ..._oadg.py:33-38: first-view cross entropy + 10 x JSD between the views, first-view smooth L1
        cfg = dict(use_sigmoid=False, loss_weight=1.0, num_views=2, additional_loss='jsdv1_3_2aug', lambda_weight=10,
                   wandb_name='roi_cls')
        cfg.update({k: v for k, v in (loss_cls or {}).items() if k != 'type'})
        self.loss_cls = CrossEntropyLossPlus(**cfg)
        cfg = dict(beta=1.0, loss_weight=1.0, num_views=2, additional_loss='None', lambda_weight=0.0, wandb_name='roi_bbox')
        cfg.update({k: v for k, v in (loss_bbox or {}).items() if k != 'type'})
        self.loss_bbox = SmoothL1LossPlus(**cfg)
        self.register_buffer('target_stds', torch.tensor(target_stds, dtype=torch.float32), persistent=False)

    def forward(self, x):
        x = x.flatten(1)
        for fc in self.shared_fcs:
            x = F.relu(fc(x))
        return self.fc_cls(x), self.fc_reg(x), self.fc_cont(x)

    def get_targets(self, sampling_results):
        """bbox_head.py:328-394 (labels / weights / DeltaXYWH targets per sampled RoI; background = num_classes)."""
        labels, weights, targets, tweights = [], [], [], []
        for res in sampling_results:
            n_pos, n_neg = res.pos_bboxes.shape[0], res.neg_bboxes.shape[0]
            lab = res.pos_bboxes.new_full((n_pos + n_neg,), self.num_classes, dtype=torch.long)
            lab[:n_pos] = res.pos_gt_labels
            t = res.pos_bboxes.new_zeros((n_pos + n_neg, 4))
            tw = res.pos_bboxes.new_zeros((n_pos + n_neg, 4))
            if n_pos:
                p, g = res.pos_bboxes, res.pos_gt_bboxes
                pw, ph = p[:, 2] - p[:, 0], p[:, 3] - p[:, 1]
                gw, gh = g[:, 2] - g[:, 0], g[:, 3] - g[:, 1]
                dx = ((g[:, 0] + g[:, 2]) - (p[:, 0] + p[:, 2])) * 0.5 / pw
                dy = ((g[:, 1] + g[:, 3]) - (p[:, 1] + p[:, 3])) * 0.5 / ph
                t[:n_pos] = torch.stack([dx, dy, torch.log(gw / pw), torch.log(gh / ph)], dim=1) / self.target_stds
                tw[:n_pos] = 1.0
            labels.append(lab)
            weights.append(lab.new_ones(lab.shape, dtype=torch.float32))
            targets.append(t)
            tweights.append(tw)
        return torch.cat(labels), torch.cat(weights), torch.cat(targets), torch.cat(tweights)

    def loss(self, cls_score, bbox_pred, cont_feats, labels, label_weights, bbox_targets, bbox_weights):
        """contrastive_head.py:60-138: CrossEntropyLossPlus over the sampled RoIs of all views, SmoothL1LossPlus on
        the positives, and the gated contrastive term on the embeddings (which may carry extra random-proposal rows)."""
        losses = {}
        avg = max(float((label_weights > 0).sum()), 1.0)           # :76 (a host sync in the reference as well)
        losses['loss_cls'] = self.loss_cls(cls_score, labels, label_weights, avg_factor=avg)
        pos = (labels >= 0) & (labels < self.num_classes)
        if pos.any():
            pred = bbox_pred.view(bbox_pred.shape[0], -1, 4)[pos, labels[pos]]
            losses['loss_bbox'] = self.loss_bbox(pred, bbox_targets[pos], bbox_weights[pos],
                                                 avg_factor=bbox_targets.shape[0])
        else:
            losses['loss_bbox'] = bbox_pred[pos].sum()
        lab = labels.contiguous().view(-1, 1)
        if self.gather and cont_feats is not None and _world_size() > 1:
            # the foreground gate is the kernel's, on the labels of all ranks: a per-rank gate could leave a rank out
            # of the exchange the others wait for
            from .distributed import gathered_contrastive_loss
            lc = self.loss_cont
            losses['loss_cont'] = gathered_contrastive_loss(cont_feats, lab, temperature=lc.temperature,
                                                            loss_weight=lc.loss_weight, min_samples=lc.min_samples,
                                                            normalized_input=lc.normalized_input)
            return losses
        n_fg = int((lab != lab.max()).sum())              # the reference's host sync (:125-126)
        if cont_feats is not None and cont_feats.numel() > 0 and n_fg > self.loss_cont.min_samples:
            losses['loss_cont'] = self.loss_cont(cont_feats, lab)
        elif cont_feats is not None:
            # DDP: the contrastive branch must take part in every backward; the reference omits the key
            losses['loss_cont'] = cont_feats.sum() * 0.0
        return losses


class TwoViewRoIHead(nn.Module):
    """ContrastiveRoIHead (roi_heads/contrastive_roi_head.py) on torchvision's MultiScaleRoIAlign."""

    def __init__(self, num_classes=8, featmap_names=('0', '1', '2', '3'), roi_size=7, num=512, pos_fraction=0.25,
                 loss_cont=None, in_channels=256, gather=False, rp_per_image=None):
        super().__init__()
        from torchvision.ops import MultiScaleRoIAlign
        self.roi_align = MultiScaleRoIAlign(list(featmap_names), roi_size, 0)
        self.bbox_head = Shared2FCContrastiveHead(in_channels=in_channels, num_classes=num_classes, roi_feat_size=roi_size,
                                                  loss_cont=loss_cont, gather=gather)
        self.rp_per_image = rp_per_image if rp_per_image is not None else (10 if gather else None)
        self.num, self.pos_fraction = num, pos_fraction
        self.last_rois = None

    def _extract(self, feats, box_lists, image_shapes):
        return self.roi_align(feats, [b[:, :4] for b in box_lists], image_shapes)

    def forward_train(self, feats, image_shapes, proposal_list, gt_bboxes, gt_labels, num_views, batch_size,
                      random_proposal_list=None, generator=None):
        """feats: OrderedDict of FPN maps of the INTEGRATED batch (B * num_views images, view-major)."""
        sampling = replicate_sampling(proposal_list, gt_bboxes, gt_labels, batch_size, num_views, num=self.num,
                                      pos_fraction=self.pos_fraction, generator=generator)
        boxes = [res.bboxes for res in sampling]
        self.last_rois = bbox2roi(boxes)
        cls_score, bbox_pred, cont_feats = self.bbox_head(self._extract(feats, boxes, image_shapes))
        if random_proposal_list is not None:              # contrastive_roi_head.py:146-149: embeddings only
            if self.rp_per_image is not None:
                random_proposal_list = [fixed_count(b, self.rp_per_image) for b in random_proposal_list]
            _, _, cont_rp = self.bbox_head(self._extract(feats, list(random_proposal_list), image_shapes))
            cont_feats = torch.cat([cont_feats, cont_rp], dim=0)
        targets = self.bbox_head.get_targets(sampling)
        return self.bbox_head.loss(cls_score, bbox_pred, cont_feats, *targets)


class _DC5Backbone(nn.Module):
    """ResNet with a dilated last stage and no neck: one stride-16 map of 2048 channels (the reference's
    ``configs/_base_/models/faster_rcnn_r50_caffe_dc5.py``: strides (1, 2, 2, 1), dilations (1, 1, 1, 2), out_indices (3,))."""

    def __init__(self, name):
        super().__init__()
        import torchvision
        net = getattr(torchvision.models, name)(weights=None, replace_stride_with_dilation=[False, False, True])
        self.body = nn.Sequential(net.conv1, net.bn1, net.relu, net.maxpool, net.layer1, net.layer2, net.layer3, net.layer4)
        self.out_channels = 2048

    def forward(self, x):
        from collections import OrderedDict
        return OrderedDict([('0', self.body(x))])


class TwoViewFasterRCNN(nn.Module):
    """Faster R-CNN (torchvision backbone + FPN or dilated-C5, RPN: stock torch) with the OA-DG two-view step around it
    (detectors/two_stage.py forward_train): integrate the views, RPN on all of them, RoIs sampled on view 1 and
    replicated, random proposals, contrastive head.  ``arch='fpn'``: BASELINE configs 3 / 4 (R50-FPN);
    ``arch='dc5'``: config 5 (R101-DC5, ``configs/OA-DG/dwd/faster_rcnn_r101_dc5_1x_dwd_oadg.py``)."""

    def __init__(self, num_classes=8, backbone='resnet50', trainable_layers=5, rpn_pre_nms=2000, rpn_post_nms=1000,
                 random_proposal_cfg=None, loss_cont=None, arch='fpn', gather=False):
        super().__init__()
        from torchvision.models.detection.anchor_utils import AnchorGenerator
        from torchvision.models.detection.rpn import RegionProposalNetwork, RPNHead
        if arch == 'fpn':
            from torchvision.models.detection.backbone_utils import resnet_fpn_backbone
            self.backbone = resnet_fpn_backbone(backbone_name=backbone, weights=None, trainable_layers=trainable_layers)
            anchors = AnchorGenerator(sizes=((32,), (64,), (128,), (256,), (512,)), aspect_ratios=((0.5, 1.0, 2.0),) * 5)
            names, per_loc = ('0', '1', '2', '3'), 3
        elif arch == 'dc5':
            self.backbone = _DC5Backbone(backbone)
            anchors = AnchorGenerator(sizes=((32, 64, 128, 256, 512),), aspect_ratios=((0.5, 1.0, 2.0),))
            names, per_loc = ('0',), 15
        else:
            raise ValueError("arch: 'fpn' or 'dc5'")
        self.rpn = RegionProposalNetwork(anchors, RPNHead(self.backbone.out_channels, per_loc), 0.7, 0.3, 256, 0.5,
                                         dict(training=rpn_pre_nms, testing=1000), dict(training=rpn_post_nms, testing=1000), 0.7)
        self.roi_head = TwoViewRoIHead(num_classes=num_classes, featmap_names=names, loss_cont=loss_cont,
                                       in_channels=self.backbone.out_channels, gather=gather)
        self.rpn_loss = TwoViewRPNLoss()
        self.random_proposal_cfg = random_proposal_cfg

    def forward(self, data, generator=None):
        """``forward_train`` under the name DistributedDataParallel wraps."""
        return self.forward_train(data, generator=generator)

    def forward_train(self, data, generator=None):
        from torchvision.models.detection.image_list import ImageList
        data = integrate_data(data)
        img = data['img']
        nv, bs = data['num_views'], data['batch_size']
        shapes = [tuple(img.shape[-2:])] * img.shape[0]
        feats = self.backbone(img)
        targets = [dict(boxes=b, labels=l) for b, l in zip(data['gt_bboxes'], data['gt_labels'])]
        proposals, rpn_losses = rpn_forward_oadg(self.rpn, ImageList(img, shapes), feats, targets, self.rpn_loss)
        rp = None
        if self.random_proposal_cfg is not None:
            rp = random_proposals(img.shape[-2:], data['gt_bboxes'], nv, data.get('multilevel_boxes'),
                                  data.get('oamix_boxes'), generator=generator, **self.random_proposal_cfg)
        roi_feats = {k: v for k, v in feats.items() if k != 'pool'}
        losses = self.roi_head.forward_train(roi_feats, shapes, proposals, data['gt_bboxes'], data['gt_labels'], nv, bs,
                                             random_proposal_list=rp, generator=generator)
        losses.update(rpn_losses)
        return losses

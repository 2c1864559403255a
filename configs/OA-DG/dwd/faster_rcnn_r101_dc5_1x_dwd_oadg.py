# OA-DG (OA-Mix + OA-Loss) on Diverse Weather: BASELINE.json config 5.  The reference ships the DWD baseline
# (configs/OA-DG/dwd/faster_rcnn_r101_dc5_1x_dwd.py) and an OA-Mix-only variant but no OA-DG one; this file composes it
# from that baseline + the model overrides of the Cityscapes OA-DG config (rpn / roi losses with the two-view JSD term,
# contrastive RoI head, random proposals), with the 7 DWD classes.  Drop it next to the reference's dwd configs; the
# '/ws/external/' prefix is the reference's own convention (oadg_b200.Config.fromfile(..., base_remap=...) maps it).
_base_ = ['/ws/external/configs/OA-DG/dwd/faster_rcnn_r101_dc5_1x_dwd.py']

num_views = 2
num_classes = 7
oa_loss = dict(jsd_rpn=0.1, jsd_roi=10, cont=0.01, temperature=0.06)


def __plus(kind, name, **kw):
    """A first-view ("Plus") loss of the reference's registry: all of them share these keys."""
    cfg = dict(type=kind, loss_weight=1.0, num_views=num_views, wandb_name=name)
    cfg.update(kw)
    return cfg


def __jsd(weight):
    return dict(additional_loss='jsdv1_3_2aug', lambda_weight=weight)


def __no_extra():
    return dict(additional_loss='None', lambda_weight=0.0)


model = dict(
    rpn_head=dict(
        loss_cls=__plus('CrossEntropyLossPlus', 'rpn_cls', use_sigmoid=True, **__jsd(oa_loss['jsd_rpn'])),
        loss_bbox=__plus('L1LossPlus', 'rpn_bbox', **__no_extra())),
    roi_head=dict(
        type='ContrastiveRoIHead',
        bbox_head=dict(
            type='Shared2FCContrastiveHead',
            num_classes=num_classes,
            with_cont=True,
            out_dim_cont=256,
            cont_predictor_cfg=dict(num_linear=2, feat_channels=256, return_relu=True),
            loss_cls=__plus('CrossEntropyLossPlus', 'roi_cls', use_sigmoid=False, log_pos_ratio=True,
                           **__jsd(oa_loss['jsd_roi'])),
            loss_bbox=__plus('SmoothL1LossPlus', 'roi_bbox', beta=1.0, **__no_extra()),
            loss_cont=dict(type='ContrastiveLossPlus', loss_weight=oa_loss['cont'], num_views=num_views,
                           temperature=oa_loss['temperature']))),
    train_cfg=dict(random_proposal_cfg=dict(bbox_from='oagrb', num_bboxes=10, scales=(0.01, 0.3),
                                            ratios=(0.3, 1 / 0.3), iou_max=0.7, iou_min=0.0)))

# the two-view transform (num_views=2, keep_orig=True: `img` stays the source frame, `img2` is the OA-Mix view)
oamix_config = dict(
    type='OAMix', version='augmix', num_views=num_views, keep_orig=True, severity=10,
    random_box_ratio=(3, 1 / 3), random_box_scale=(0.01, 0.1),              # multi-level boxes
    oa_random_box_scale=(0.005, 0.1), oa_random_box_ratio=(3, 1 / 3),       # object-aware boxes
    spatial_ratio=4, sigma_ratio=0.3)                                       # low-resolution mask blur

custom_imports = dict(imports=['mmdet.datasets.pipelines.oa_mix'], allow_failed_imports=False)
img_norm_cfg = dict(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375], to_rgb=True)
train_pipeline = [
    dict(type='LoadImageFromFile'),
    dict(type='LoadAnnotations', with_bbox=True),
    dict(type='Resize', img_scale=(1280, 600), keep_ratio=True),            # configs/_base_/datasets/s-dgod.py:9
    dict(type='RandomFlip', flip_ratio=0.5),
    oamix_config,
    dict(type='Normalize', **img_norm_cfg),
    dict(type='Pad', size_divisor=32),
    dict(type='DefaultFormatBundle'),
    dict(type='Collect', keys=['img', 'img2', 'gt_bboxes', 'gt_bboxes2', 'gt_labels', 'multilevel_boxes',
                               'oamix_boxes']),
]
data = dict(samples_per_gpu=2, train=dict(dataset=dict(pipeline=train_pipeline)))

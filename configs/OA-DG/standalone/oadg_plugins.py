# The OA-DG plugin settings on their own (no `_base_`): what a detector config of the reference adds on top of its
# baseline to switch OA-Mix and OA-Loss on -- transform, the four first-view / consistency losses, the contrastive loss,
# random proposals -- in the reference's own keys, so that this file builds through `oadg_b200.Config` / `build_from_cfg`
# (or mmcv's) on a machine that has neither the reference tree nor mmcv.  Values: configs/OA-DG/cityscapes/
# faster_rcnn_r50_fpn_1x_cityscapes_oadg.py:5-56 of the reference.
num_views = 2

oamix_config = dict(
    type='OAMix', version='augmix', num_views=num_views, keep_orig=True, severity=10,
    random_box_ratio=(3, 1 / 3), random_box_scale=(0.01, 0.1),
    oa_random_box_scale=(0.005, 0.1), oa_random_box_ratio=(3, 1 / 3),
    spatial_ratio=4, sigma_ratio=0.3)

img_norm_cfg = dict(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375], to_rgb=True)

# this repository's extension of the transform: Normalize + Pad + DefaultFormatBundle written by the mix kernel
oamix_fused_config = dict(oamix_config, fused_output=dict(img_norm_cfg, size_divisor=32))

train_pipeline = [oamix_config, dict(type='Normalize', **img_norm_cfg), dict(type='Pad', size_divisor=32)]

losses = dict(
    rpn_cls=dict(type='CrossEntropyLossPlus', use_sigmoid=True, loss_weight=1.0, num_views=num_views,
                 additional_loss='jsdv1_3_2aug', lambda_weight=0.1, wandb_name='rpn_cls'),
    rpn_bbox=dict(type='L1LossPlus', loss_weight=1.0, num_views=num_views, additional_loss='None', lambda_weight=0.0,
                  wandb_name='rpn_bbox'),
    roi_cls=dict(type='CrossEntropyLossPlus', use_sigmoid=False, loss_weight=1.0, num_views=num_views,
                 additional_loss='jsdv1_3_2aug', lambda_weight=10, wandb_name='roi_cls', log_pos_ratio=True),
    roi_bbox=dict(type='SmoothL1LossPlus', beta=1.0, loss_weight=1.0, num_views=num_views, additional_loss='None',
                  lambda_weight=0.0, wandb_name='roi_bbox'),
    cont=dict(type='ContrastiveLossPlus', loss_weight=0.01, num_views=num_views, temperature=0.06))

random_proposal_cfg = dict(bbox_from='oagrb', num_bboxes=10, scales=(0.01, 0.3), ratios=(0.3, 1 / 0.3), iou_max=0.7,
                           iou_min=0.0)
custom_imports = dict(imports=['mmdet.datasets.pipelines.oa_mix'], allow_failed_imports=False)

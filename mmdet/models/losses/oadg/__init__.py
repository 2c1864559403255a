"""Dotted-path alias of reference mmdet/models/losses/oadg/{contrastive_loss,contrastive_loss_plus}.py."""
from oadg_b200.contrastive_loss import ContrastiveLossPlus, supcontrast, supcontrast_yolo  # noqa: F401
from oadg_b200.consistency_losses import CrossEntropyLossPlus, SmoothL1LossPlus, L1LossPlus  # noqa: F401

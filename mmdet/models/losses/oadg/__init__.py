"""Dotted-path alias of reference mmdet/models/losses/oadg/{contrastive_loss,contrastive_loss_plus}.py."""
from oadg_b200.contrastive_loss import ContrastiveLossPlus, supcontrast  # noqa: F401

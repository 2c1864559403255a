"""``from mmdet.models.builder import LOSSES, build_loss`` (reference mmdet/models/builder.py:13,43)."""
from oadg_b200.registry import MODELS, LOSSES, build_loss  # noqa: F401

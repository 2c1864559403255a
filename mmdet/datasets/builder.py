"""``from mmdet.datasets.builder import PIPELINES`` (reference mmdet/datasets/builder.py:28)."""
from oadg_b200.registry import PIPELINES, build_from_cfg  # noqa: F401

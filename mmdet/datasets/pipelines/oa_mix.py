"""Dotted-path alias of the reference module mmdet/datasets/pipelines/oa_mix.py."""
from oadg_b200.oamix import OAMix, get_aug_list  # noqa: F401

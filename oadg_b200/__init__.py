"""oadg_b200: B200-native (sm_100a) OA-Mix + OA-Loss hot path of OA-DG behind the
MMDetection-2.x plugin surface (@PIPELINES OAMix, @LOSSES ContrastiveLossPlus).

Importing the package registers both plugins in ``oadg_b200.registry`` (and in the real
mmdet registries when mmdet/mmcv are installed).  Compute entry points
need libOADG.so (``python -m oadg_b200.build``) and a CUDA device; there is no CPU fallback.
"""
from .registry import PIPELINES, LOSSES, MODELS, Registry, build_from_cfg, build_loss, Config, Compose  # noqa: F401
from .oamix import OAMix, get_aug_list  # noqa: F401
from .contrastive_loss import ContrastiveLossPlus, supcontrast, reference_pair_map  # noqa: F401

__version__ = '0.1.0'

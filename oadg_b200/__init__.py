"""oadg_b200: B200-native (sm_100a) OA-Mix + OA-Loss hot path of OA-DG behind the
MMDetection-2.x plugin surface (@PIPELINES OAMix, @LOSSES ContrastiveLossPlus).

Importing the package registers the plugins in ``oadg_b200.registry``; ``oadg_b200.plugins.register_into_mmdet()``
registers them into a real MMDetection's registries.  Compute entry points
need libOADG.so (``python -m oadg_b200.build``) and a CUDA device; there is no CPU fallback.
"""
from .registry import PIPELINES, LOSSES, MODELS, Registry, build_from_cfg, build_loss, Config, Compose  # noqa: F401
from .oamix import OAMix, get_aug_list  # noqa: F401
from .contrastive_loss import ContrastiveLossPlus, supcontrast, supcontrast_yolo, reference_pair_map  # noqa: F401
from .consistency_losses import CrossEntropyLossPlus, SmoothL1LossPlus, L1LossPlus, jsdv1_3_2aug  # noqa: F401

__version__ = '0.1.0'

"""Import the UNMODIFIED reference hot-path modules from /root/reference.

ORACLE / TEST INFRASTRUCTURE ONLY.  Dev-container only: /root/reference does
not exist on the GPU box, so nothing in ``-m gpu`` tests / smoke / bench calls
this.  It is used (a) by ``scripts/make_golden.py`` to generate the committed
fixtures under ``tests/golden/`` and (b) by ``-m "not gpu"`` tests that pin the
numpy restatements in this package to the reference itself.

mmcv / mmdet are not installed here, so the four reference files are loaded by
path under a tiny stub that provides exactly the names they import:
``mmcv``, ``mmcv.utils.build_from_cfg``, ``mmdet.datasets.builder.PIPELINES``,
``mmdet.models.builder.LOSSES``, ``mmdet.core.{PolygonMasks,find_inside_bboxes}``.
``cv2.saliency`` (opencv-contrib, absent) is provided by ``oracle.saliency_np``.
"""
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get('OADG_REFERENCE_ROOT', '/root/reference')


def available():
    return os.path.isfile(os.path.join(REF_ROOT, 'mmdet/datasets/pipelines/oa_mix.py'))


class _Registry:
    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            self.module_dict[name or cls.__name__] = cls
            return cls
        if module is not None:
            return deco(module)
        return deco

    def get(self, key):
        return self.module_dict.get(key)

    def build(self, cfg):
        cfg = dict(cfg)
        return self.module_dict[cfg.pop('type')](**cfg)


def _load(name, relpath):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


_CACHE = {}


def load_reference():
    """Return dict(OAMix=..., ContrastiveLossPlus=..., supcontrast=..., modules...)."""
    if _CACHE:
        return _CACHE
    if not available():
        raise RuntimeError('reference tree not present at %s' % REF_ROOT)
    for k in list(sys.modules):
        if k == 'mmdet' or k.startswith('mmdet.') or k == 'mmcv' or k.startswith('mmcv.'):
            raise RuntimeError('ref_loader must run in a process that has not imported '
                               'the product mmdet/mmcv shim (%s is loaded)' % k)
    from . import saliency_np
    saliency_np.install_cv2_shim()

    def pkg(name):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
        return m

    mmcv = pkg('mmcv')
    mmcv_utils = pkg('mmcv.utils')

    def build_from_cfg(cfg, registry, default_args=None):
        cfg = dict(cfg)
        if default_args:
            for k, v in default_args.items():
                cfg.setdefault(k, v)
        return registry.get(cfg.pop('type'))(**cfg)

    mmcv_utils.build_from_cfg = build_from_cfg
    mmcv_utils.Registry = _Registry
    mmcv.utils = mmcv_utils

    pkg('mmdet')
    core = pkg('mmdet.core')
    core.PolygonMasks = type('PolygonMasks', (), {})
    core.find_inside_bboxes = lambda *a, **k: None
    pkg('mmdet.core.evaluation')
    _load('mmdet.core.evaluation.bbox_overlaps', 'mmdet/core/evaluation/bbox_overlaps.py')
    pkg('mmdet.datasets')
    builder = pkg('mmdet.datasets.builder')
    builder.PIPELINES = _Registry('pipeline')
    pkg('mmdet.datasets.pipelines')
    _load('mmdet.datasets.pipelines.compose', 'mmdet/datasets/pipelines/compose.py')
    augmix = _load('mmdet.datasets.pipelines.augmix', 'mmdet/datasets/pipelines/augmix.py')
    bbox_aug = _load('mmdet.datasets.pipelines.bbox_augmentation',
                     'mmdet/datasets/pipelines/bbox_augmentation.py')
    oa_mix = _load('mmdet.datasets.pipelines.oa_mix', 'mmdet/datasets/pipelines/oa_mix.py')

    pkg('mmdet.models')
    mbuilder = pkg('mmdet.models.builder')
    mbuilder.LOSSES = _Registry('models')
    pkg('mmdet.models.losses')
    pkg('mmdet.models.losses.oadg')
    closs = _load('mmdet.models.losses.oadg.contrastive_loss',
                  'mmdet/models/losses/oadg/contrastive_loss.py')
    clossp = _load('mmdet.models.losses.oadg.contrastive_loss_plus',
                   'mmdet/models/losses/oadg/contrastive_loss_plus.py')

    # the consistency losses: losses/utils.py + oadg/{cross_entropy,smooth_l1}_loss_plus.py
    mmcv.jit = lambda *a, **k: (lambda f: f)
    _load('mmdet.models.losses.utils', 'mmdet/models/losses/utils.py')
    cel = _load('mmdet.models.losses.oadg.cross_entropy_loss_plus', 'mmdet/models/losses/oadg/cross_entropy_loss_plus.py')
    sl1 = _load('mmdet.models.losses.oadg.smooth_l1_loss_plus', 'mmdet/models/losses/oadg/smooth_l1_loss_plus.py')
    _CACHE.update(dict(CrossEntropyLossPlus=cel.CrossEntropyLossPlus, jsdv1_3_2aug=cel.jsdv1_3_2aug,
                       SmoothL1LossPlus=sl1.SmoothL1LossPlus, L1LossPlus=sl1.L1LossPlus,
                       supcontrast_yolo=closs.supcontrast_yolo))
    _CACHE.update(dict(
        OAMix=oa_mix.OAMix, oa_mix=oa_mix, augmix=augmix, bbox_augmentation=bbox_aug,
        supcontrast=closs.supcontrast, supcontrast_mask=closs.supcontrast_mask,
        ContrastiveLossPlus=clossp.ContrastiveLossPlus,
        PIPELINES=builder.PIPELINES, LOSSES=mbuilder.LOSSES))
    return _CACHE

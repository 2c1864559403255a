"""Register the B200 implementations into a REAL MMDetection's registries, over the stock classes of the same names.

    # configs/.../my_oadg.py
    custom_imports = dict(imports=['mmdet.datasets.pipelines.oa_mix', 'oadg_b200.plugins'], allow_failed_imports=False)

Importing this module calls ``register_into_mmdet()``.  The reference keeps its registries in
``mmdet/datasets/builder.py:28`` (``PIPELINES``) and ``mmdet/models/builder.py:13`` (``LOSSES``); both are
``mmcv.utils.Registry`` objects whose ``register_module(name=, force=, module=)`` replaces an existing entry.  When
the ``mmdet`` that resolves is this repository's import-path shim, its registries ARE ``oadg_b200.registry`` and
there is nothing to do."""
import importlib

from . import consistency_losses, contrastive_loss, oamix
from . import registry as _own

PIPELINE_CLASSES = {'OAMix': oamix.OAMix}
LOSS_CLASSES = {'ContrastiveLossPlus': contrastive_loss.ContrastiveLossPlus,
                'CrossEntropyLossPlus': consistency_losses.CrossEntropyLossPlus,
                'SmoothL1LossPlus': consistency_losses.SmoothL1LossPlus,
                'L1LossPlus': consistency_losses.L1LossPlus}


def register_into_mmdet(pipelines=None, losses=None, which=None):
    """``pipelines`` / ``losses``: registry objects (default: ``mmdet.datasets.builder.PIPELINES`` and
    ``mmdet.models.builder.LOSSES`` of whatever ``mmdet`` is importable).  ``which``: iterable of class names to
    register (default: all).  Returns the list of names registered into a foreign registry."""
    if pipelines is None:
        pipelines = importlib.import_module('mmdet.datasets.builder').PIPELINES
    if losses is None:
        losses = importlib.import_module('mmdet.models.builder').LOSSES
    done = []
    for reg, own, classes in ((pipelines, _own.PIPELINES, PIPELINE_CLASSES), (losses, _own.LOSSES, LOSS_CLASSES)):
        if reg is own:
            continue        # the shim: already registered by importing oadg_b200
        for name, cls in classes.items():
            if which is not None and name not in which:
                continue
            reg.register_module(name=name, force=True, module=cls)
            done.append(name)
    return done


try:
    REGISTERED = register_into_mmdet()
except ImportError:      # no mmdet at all: the package's own registry serves (oadg_b200.build_from_cfg)
    REGISTERED = []

"""Import-path shim (NOT MMDetection): lets the reference configs' dotted paths resolve to the
B200 plugins when the real mmdet is not installed, e.g.
``custom_imports = dict(imports=['mmdet.datasets.pipelines.oa_mix'])``
(reference configs/OA-DG/cityscapes/faster_rcnn_r50_fpn_1x_cityscapes_oadg.py:61).

The shim never shadows a real MMDetection: when another ``mmdet`` package is importable from the rest of
``sys.path`` (this repository's root merely comes first, as it does under pytest or ``python bench.py``), that
package is loaded in this one's place, and the B200 classes go into ITS registries through ``oadg_b200.plugins``
(INTEGRATION.md)."""
import importlib.machinery
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_root = os.path.dirname(_here)
_elsewhere = [p for p in sys.path if os.path.abspath(p or os.getcwd()) != _root]
_spec = importlib.machinery.PathFinder.find_spec('mmdet', _elsewhere)
if _spec is not None and _spec.origin and os.path.dirname(os.path.abspath(_spec.origin)) != _here:
    _real = importlib.util.module_from_spec(_spec)
    sys.modules['mmdet'] = _real          # `import mmdet` hands out what sys.modules holds once this file returns
    _spec.loader.exec_module(_real)
else:
    __version__ = '2.20.0+oadg_b200'

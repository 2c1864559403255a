"""Import-path shim (NOT MMDetection): lets the reference configs' dotted paths resolve to the
B200 plugins when the real mmdet is not installed, e.g.
``custom_imports = dict(imports=['mmdet.datasets.pipelines.oa_mix'])``
(reference configs/OA-DG/cityscapes/faster_rcnn_r50_fpn_1x_cityscapes_oadg.py:61).
With a real mmdet on the path, register through ``oadg_b200.plugins`` instead."""
__version__ = '2.20.0+oadg_b200'

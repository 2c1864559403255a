"""Host logic without a GPU: C ABI surface, plan records, registry / config surface, plan sampler vs oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import OAMIX_CFG, ROOT, sampler_cfg
from oracle import oamix_np, synth

REF = '/root/reference'


def test_c_abi_exports_every_declared_symbol(libpath):
    from oadg_b200 import _lib, plan
    header = open(os.path.join(ROOT, 'include', 'oadg.h')).read()
    declared = set(re.findall(r'\b(oadg_[a-z_0-9]+)\s*\(', header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = ctypes.CDLL(libpath)
    for name in declared:
        assert hasattr(lib, name), name
    lib = _lib.load()
    sizes = (ctypes.c_int32 * 6)()
    lib.oadg_struct_sizes(sizes)
    assert list(sizes) == plan.STRUCT_SIZES
    assert lib.oadg_error_string(-2).startswith(b'OADG_E_PLAN')


def test_plan_blob_validation(libpath):
    from oadg_b200 import _lib
    lib = _lib.load()
    need = ctypes.c_size_t(0)
    junk = np.zeros(64, np.uint8)
    assert lib.oadg_oamix_workspace_bytes(junk.ctypes.data, junk.nbytes, ctypes.byref(need)) == -2
    assert lib.oadg_oamix_workspace_bytes(None, 0, ctypes.byref(need)) == -1
    from oadg_b200.oamix import OAMix
    img, gt = synth.make_image(0, 96, 160, 3)
    np.random.seed(1)
    t = OAMix(version='augmix')
    vp = t._sample_head(96, 160, gt)
    t._sample_tail(vp, gt, [20.0, 5.0, -1])
    blob = t._pack([(vp, gt, 0)])
    assert lib.oadg_oamix_workspace_bytes(blob.ctypes.data, blob.nbytes, ctypes.byref(need)) == 0
    assert need.value > 96 * 160 * 3 * 6
    bad = blob.copy()
    bad[8:12] = np.frombuffer(np.int32(-5).tobytes(), np.uint8)   # n_views < 0
    assert lib.oadg_oamix_workspace_bytes(bad.ctypes.data, bad.nbytes, ctypes.byref(need)) == -2


def test_product_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    from oadg_b200 import OAMix, ContrastiveLossPlus, _lib
    img, gt = synth.make_image(0, 64, 64, 1)
    with pytest.raises(_lib.OADGError):
        OAMix()(dict(img=img, gt_bboxes=gt))
    with pytest.raises(_lib.OADGError):
        ContrastiveLossPlus()(torch.randn(2048, 256), torch.zeros(2048, 1, dtype=torch.int64))


def test_registry_surface_and_ctor_contract():
    import oadg_b200
    from oadg_b200 import PIPELINES, LOSSES, build_from_cfg, build_loss
    import mmdet.datasets.pipelines.oa_mix as shim
    from mmdet.datasets.builder import PIPELINES as P2
    from mmdet.models.builder import LOSSES as L2
    assert P2 is PIPELINES and L2 is LOSSES and shim.OAMix is oadg_b200.OAMix
    t = build_from_cfg(dict(type='OAMix', version='augmix.all', use_mix=True, mixture_width=1, mixture_depth=-1,
                            use_oa=True, oa_version='saliency_sparse', use_mrange=False, use_multilevel=True), PIPELINES)
    assert t.mixture_width == 1 and t.kwargs['oa_version'] == 'saliency_sparse' and len(t.aug_list) == 15
    assert t.score_thresh == 10 and t.aug_prob_coeff == 1.0 and repr(t) == 'OAMix'
    with pytest.raises(NotImplementedError):
        build_from_cfg(dict(type='OAMix', version='nope'), PIPELINES)
    with pytest.warns(UserWarning):
        oadg_b200.OAMix(num_views=1, keep_orig=True)
    loss = build_loss(dict(type='ContrastiveLossPlus', loss_weight=0.01, num_views=2, temperature=0.06, version='r-cnn'))
    assert (loss.loss_weight, loss.temperature, loss.num_views, loss.min_samples, loss.normalized_input) == (0.01, 0.06, 2, 10, True)
    assert loss.kwargs == dict(version='r-cnn')
    import torch
    out = loss(torch.zeros(0, 256), torch.zeros(0, 1, dtype=torch.int64))   # contrastive_loss_plus.py:38-39
    assert out.shape == (1,) and out.device.type == 'cpu' and float(out) == 0.0
    with pytest.raises(KeyError):
        build_from_cfg(dict(type='Missing'), PIPELINES)


@pytest.mark.skipif(not os.path.isdir(REF + '/configs/OA-DG'), reason='reference configs only exist in the dev container')
def test_reference_configs_load_unchanged_and_build_our_plugins():
    from oadg_b200 import Config, PIPELINES, build_from_cfg, build_loss, OAMix, ContrastiveLossPlus
    remap = {'/ws/external/': REF + '/'}
    n = 0
    for sub, dirs, files in os.walk(REF + '/configs/OA-DG'):
        for f in sorted(files):
            if not f.endswith('.py'):
                continue
            cfg = Config.fromfile(os.path.join(sub, f), base_remap=remap)
            n += 1
            if 'oamix_config' in cfg:
                t = build_from_cfg(cfg.oamix_config, PIPELINES)
                assert isinstance(t, OAMix)
                if 'train_pipeline' in cfg:
                    assert any(isinstance(s, dict) and s.get('type') == 'OAMix' for s in cfg.train_pipeline)
            lc = cfg.get('model', {}).get('roi_head', {}).get('bbox_head', {}).get('loss_cont')
            if lc is not None:
                loss = build_loss(lc)
                assert isinstance(loss, ContrastiveLossPlus) and loss.temperature == 0.06 and loss.loss_weight == 0.01
    assert n >= 8
    cfg = Config.fromfile(REF + '/configs/OA-DG/cityscapes/faster_rcnn_r50_fpn_1x_cityscapes_oadg.py', base_remap=remap)
    assert cfg.model.roi_head.type == 'ContrastiveRoIHead' and cfg.model.backbone.depth == 50   # merged through _base_
    assert cfg.data.samples_per_gpu == 2 and cfg.custom_imports['imports'] == ['mmdet.datasets.pipelines.oa_mix']


def test_group_sizes_of_a_loop_of_known_length():
    from oadg_b200.oamix import OAMix
    for total in list(range(0, 40)) + [100, 101, 1000]:
        for gmax in (1, 2, 3, 4, 8):
            for first in (1, 2, 4):
                sizes = OAMix._group_sizes(total, gmax, first)
                assert sum(sizes) == total and all(1 <= g <= gmax for g in sizes), (total, gmax, first, sizes)
                # at most one group below full size after the ramp-up, and only when the ramp-up is saturated
                tail = sizes[3:]
                assert sum(1 for g in tail if g < gmax) <= 1
    assert OAMix._group_sizes(20, 4) == [2, 2, 4, 4, 4, 4]        # not 1, 2, 4, 4, 4, 4, 1
    assert OAMix._group_sizes(19, 4) == [1, 2, 4, 4, 4, 4]
    assert OAMix._group_sizes(3, 4) == [1, 2]


def test_mmdet_shim_steps_aside_for_a_real_mmdet(tmp_path):
    """With another `mmdet` package further down sys.path, `import mmdet` must give THAT package even though the
    repository root (with the shim) comes first."""
    import subprocess
    import sys
    pkg = tmp_path / 'site' / 'mmdet'
    pkg.mkdir(parents=True)
    (pkg / '__init__.py').write_text("REAL = True\n__version__ = '2.20.0'\n")
    (pkg / 'sub.py').write_text('X = 1\n')
    code = ("import sys; sys.path[:0] = [%r]; sys.path.append(%r); import mmdet, mmdet.sub; "
            "print(getattr(mmdet, 'REAL', False), mmdet.__version__, mmdet.sub.X)" % (ROOT, str(tmp_path / 'site')))
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, cwd=str(tmp_path))
    assert out.stdout.split() == ['True', '2.20.0', '1'], out.stdout + out.stderr
    # and without one, the shim serves the reference configs' dotted paths
    code = "import sys; sys.path[:0] = [%r]; import mmdet, mmdet.datasets.pipelines.oa_mix as m; print(mmdet.__version__, m.OAMix.__name__)" % ROOT
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, cwd=str(tmp_path))
    assert out.stdout.split() == ['2.20.0+oadg_b200', 'OAMix'], out.stdout + out.stderr


def test_standalone_plugin_config_builds_everywhere():
    """configs/OA-DG/standalone/oadg_plugins.py has no `_base_`: the plugin surface is exercised on boxes that have
    neither the reference tree nor mmcv (the reference's own configs are loaded by the dev-container tests above)."""
    from oadg_b200 import (Config, PIPELINES, build_from_cfg, build_loss, OAMix, ContrastiveLossPlus,
                           CrossEntropyLossPlus, SmoothL1LossPlus, L1LossPlus)
    cfg = Config.fromfile(os.path.join(ROOT, 'configs/OA-DG/standalone/oadg_plugins.py'))
    t = build_from_cfg(cfg.oamix_config, PIPELINES)
    assert isinstance(t, OAMix) and t.num_views == 2 and t.keep_orig and t.fused_output is None
    f = build_from_cfg(cfg.oamix_fused_config, PIPELINES)
    assert f.fused_output['size_divisor'] == 32 and f.fused_output['to_rgb'] is True
    built = {k: build_loss(v) for k, v in cfg.losses.items()}
    assert [type(built[k]) for k in ('rpn_cls', 'rpn_bbox', 'roi_cls', 'roi_bbox', 'cont')] == \
        [CrossEntropyLossPlus, L1LossPlus, CrossEntropyLossPlus, SmoothL1LossPlus, ContrastiveLossPlus]
    assert built['rpn_cls'].use_sigmoid and built['rpn_cls'].lambda_weight == 0.1 and built['roi_cls'].lambda_weight == 10
    assert built['cont'].temperature == 0.06 and built['cont'].loss_weight == 0.01
    assert cfg.train_pipeline[0]['type'] == 'OAMix' and cfg.random_proposal_cfg['num_bboxes'] == 10


def test_register_into_a_foreign_mmdet_registry_replaces_the_stock_classes():
    """oadg_b200.plugins against a stand-in for mmcv.utils.Registry (same register_module(name, force, module)
    contract, reference mmdet/datasets/builder.py:28, mmdet/models/builder.py:13)."""
    from oadg_b200 import plugins, OAMix, ContrastiveLossPlus, CrossEntropyLossPlus

    class Stock:
        pass

    class MmcvLikeRegistry:
        def __init__(self):
            self.module_dict = {'OAMix': Stock, 'ContrastiveLossPlus': Stock, 'CrossEntropyLossPlus': Stock}

        def register_module(self, name=None, force=False, module=None):
            if name in self.module_dict and not force:
                raise KeyError(name + ' is already registered')
            self.module_dict[name] = module

    pipes, losses = MmcvLikeRegistry(), MmcvLikeRegistry()
    done = plugins.register_into_mmdet(pipes, losses)
    assert sorted(done) == ['ContrastiveLossPlus', 'CrossEntropyLossPlus', 'L1LossPlus', 'OAMix', 'SmoothL1LossPlus']
    assert pipes.module_dict['OAMix'] is OAMix and losses.module_dict['ContrastiveLossPlus'] is ContrastiveLossPlus
    assert losses.module_dict['CrossEntropyLossPlus'] is CrossEntropyLossPlus
    only = plugins.register_into_mmdet(MmcvLikeRegistry(), MmcvLikeRegistry(), which=['OAMix'])
    assert only == ['OAMix']
    # with this repository's mmdet shim on the path the registries are our own: nothing is re-registered
    assert plugins.register_into_mmdet() == []


@pytest.mark.skipif(not os.path.isdir(REF + '/configs/OA-DG'), reason='its _base_ files only exist in the dev container')
def test_composed_dwd_oadg_config_builds_every_plugin_of_the_step():
    """BASELINE config 5 (configs/OA-DG/dwd/faster_rcnn_r101_dc5_1x_dwd_oadg.py, composed here: the reference has no
    DWD OA-DG config): R101-DC5 baseline + the OA-DG losses with 7 classes + the two-view pipeline."""
    from oadg_b200 import (Config, PIPELINES, build_from_cfg, build_loss, OAMix, ContrastiveLossPlus,
                           CrossEntropyLossPlus, SmoothL1LossPlus, L1LossPlus)
    cfg = Config.fromfile(os.path.join(ROOT, 'configs/OA-DG/dwd/faster_rcnn_r101_dc5_1x_dwd_oadg.py'),
                          base_remap={'/ws/external/': REF + '/'})
    assert cfg.model.backbone.depth == 101 and cfg.model.roi_head.type == 'ContrastiveRoIHead'
    head = cfg.model.roi_head.bbox_head
    assert head.type == 'Shared2FCContrastiveHead' and head.num_classes == 7
    built = [build_loss(head.loss_cls), build_loss(head.loss_bbox), build_loss(head.loss_cont),
             build_loss(cfg.model.rpn_head.loss_cls), build_loss(cfg.model.rpn_head.loss_bbox)]
    assert [type(b) for b in built] == [CrossEntropyLossPlus, SmoothL1LossPlus, ContrastiveLossPlus,
                                        CrossEntropyLossPlus, L1LossPlus]
    assert built[0].lambda_weight == 10 and built[3].lambda_weight == 0.1 and built[3].use_sigmoid
    assert built[2].temperature == 0.06 and built[2].loss_weight == 0.01
    assert cfg.model.train_cfg.random_proposal_cfg['bbox_from'] == 'oagrb'
    t = build_from_cfg(cfg.oamix_config, PIPELINES)
    assert isinstance(t, OAMix) and t.num_views == 2 and t.keep_orig
    assert [s['type'] for s in cfg.train_pipeline][4:7] == ['OAMix', 'Normalize', 'Pad']
    assert cfg.data.samples_per_gpu == 2 and cfg.optimizer.lr == 0.001          # inherited from the DWD baseline


@pytest.mark.parametrize('case', [('augmix', 96, 160, 3, 0, 100, {}), ('augmix.all', 120, 200, 4, 1, 201, {}),
                                  ('augmix', 64, 64, 0, 2, 302, dict(mixture_width=1)),
                                  ('augmix.all', 80, 90, 2, 3, 7, dict(mixture_width=2, mixture_depth=2))])
def test_plan_sampler_consumes_rng_like_the_oracle(case):
    from oadg_b200.oamix import OAMix
    version, h, w, n_gt, s, seed, extra = case
    cfg = sampler_cfg(dict(OAMIX_CFG, version=version, **extra))
    img, gt = synth.make_image(s, h, w, n_gt)
    np.random.seed(seed)
    plan = oamix_np.sample_plan(img, gt, **cfg)
    st_o = np.random.get_state()
    np.random.seed(seed)
    t = OAMix(**cfg)
    vp = t._sample_head(h, w, gt)
    t._sample_tail(vp, gt, plan['scores'])
    st_p = np.random.get_state()
    assert st_o[2] == st_p[2] and np.array_equal(st_o[1], st_p[1])
    assert np.array_equal(vp.ml_boxes, plan['ml_boxes']) and vp.ml_boxes.dtype == np.int64
    assert len(vp.oa_boxes) == len(plan['oa_boxes']) and all(np.array_equal(a, b) for a, b in zip(vp.oa_boxes, plan['oa_boxes']))
    assert np.array_equal(vp.ws, plan['ws']) and vp.m == plan['m'] and vp.m_oa == plan['m_oa']
    assert vp.oa_low == plan['oa_low_fg']
    for steps_p, steps_o in zip(vp.ops, plan['branches']):
        assert len(steps_p) == len(steps_o)
        for regs_p, regs_o in zip(steps_p, steps_o):
            for a, b in zip(regs_p, regs_o):
                if b['name'].startswith('bg_only'):
                    from oracle import prims_np
                    assert a[0] == 'bg_affine' and np.array_equal(np.array(a[1]), prims_np.invert_affine(b['M']))
                elif b['name'].startswith('bboxes_only'):
                    assert a[0] == 'bbo_affine' and [k for k, _ in a[1]] == [x['k'] for x in b['boxes']]
                else:
                    assert a[0] == b['name']


@pytest.mark.parametrize('case', [('augmix', 1024, 2048, 8, {}), ('augmix.all', 120, 200, 4, {}),
                                  ('augmix.all', 99, 131, 6, dict(mixture_width=4, mixture_depth=3)),
                                  ('augmix', 64, 64, 0, {}), ('augmix', 40, 36, 1, {})])
def test_native_sampler_matches_python_sampler_byte_for_byte(libpath, case):
    """libOADG's oadg_oamix_sample_plan (the product path) against the Python statement of the draw order
    (_sample_head/_sample_tail/_pack, itself pinned to the oracle above): identical plan blob, identical boxes and
    identical np.random state afterwards, image after image in one batch; same ValueError when no box fits."""
    from oadg_b200.oamix import OAMix
    version, h, w, n_gt, extra = case
    cfg = dict(OAMIX_CFG, version=version, **extra)
    for seed in range(40):
        t = OAMix(**cfg)
        gts = [synth.make_image(s, h, w, n_gt)[1] for s in (seed, seed + 1)]
        if n_gt >= 2:
            gts[0][1] = [5.3, 5.9, 7.1, 40.2]       # narrower than spatial_ratio
        table = [3.5, 20.0, -1, 9.99, 10.0, 11.0, 50.0, 0.0]
        scores = [[np.float64(table[(k + i) % 8]) for k in range(len(g))] for i, g in enumerate(gts)]
        np.random.seed(seed)
        jobs, err_py = [], None
        try:
            for i, g in enumerate(gts):
                vp = t._sample_head(h, w, g)
                t._sample_tail(vp, g, scores[i])
                jobs.append((vp, g, i))
            blob_py = t._pack(jobs)
        except ValueError as e:
            err_py = e
        st_py = np.random.get_state()
        np.random.seed(seed)
        plan, err_c = None, None
        try:
            plan = t.sample_plan([(h, w)] * 2, gts, scores)
        except ValueError as e:
            err_c = e
        st_c = np.random.get_state()
        assert (err_py is None) == (err_c is None)
        assert st_py[2] == st_c[2] and np.array_equal(st_py[1], st_c[1])
        if plan is None:
            continue
        assert np.array_equal(blob_py, plan.blob)
        for (vp, _, _), ml, oa, ds in zip(jobs, plan.ml_boxes, plan.oa_boxes, plan.depth_sums):
            assert np.array_equal(vp.ml_boxes, ml) and ml.dtype == np.int64 and oa.dtype == np.int64
            assert len(vp.oa_boxes) == len(oa) and all(np.array_equal(a, b) for a, b in zip(vp.oa_boxes, oa))
            assert ds == sum(vp.depths)


def test_native_sampler_raises_like_the_reference_when_no_box_fits(libpath):
    from oadg_b200.oamix import OAMix
    t = OAMix(version='augmix', random_box_scale=(0.9, 0.99))      # boxes nearly never fit a 8x8 frame
    hits = 0
    for seed in range(20):
        np.random.seed(seed)
        try:
            t.sample_plan([(8, 8)], [np.zeros((0, 4), np.float32)], [[]])
        except ValueError:
            hits += 1
    assert hits > 0


def test_pair_map_and_row_check():
    from oadg_b200 import reference_pair_map
    from oracle import supcon_np
    for n in (2048, 2088, 2085, 3100):
        assert np.array_equal(reference_pair_map(n), supcon_np.pair_map(n))
    with pytest.raises(RuntimeError):
        reference_pair_map(1024)


def test_view_buffer_pool_recycles_on_drop_and_shares_size_classes(monkeypatch):
    """OAMix._pinned_out (host logic of the loader loop; page-locking itself needs a GPU box, so torch.empty is
    relieved of pin_memory here): `reserve` buffers of a size class are made at first use, a buffer comes back only
    when the array AND every view of it are gone, and shapes of the same 1 MiB class share buffers."""
    import torch
    import oadg_b200.oamix as m
    real_empty = torch.empty
    monkeypatch.setattr(torch, 'empty', lambda *a, **k: real_empty(*a, **{x: y for x, y in k.items() if x != 'pin_memory'}))
    monkeypatch.setattr(m._lib, 'require_cuda', lambda: torch)
    t = m.OAMix(version='augmix')
    view, a = t._pinned_out((10, 20, 3), reserve=4)
    pool = t._host_state['out_pool']
    cls = 1 << 20
    assert pool['made'] == {cls: 4} and len(pool['free'][cls]) == 3 and a.shape == (10, 20, 3) and a.dtype == np.uint8
    assert view.data_ptr() == a.ctypes.data
    a[:] = 7
    part = a[2:4]
    del a, view
    assert len(pool['free'][cls]) == 3          # a view of the array still holds the buffer
    del part
    assert len(pool['free'][cls]) == 4
    _, b = t._pinned_out((100, 200, 3))         # another shape, same class: no new buffer
    assert pool['made'] == {cls: 4} and b.shape == (100, 200, 3)
    held = [t._pinned_out((10, 20, 3)) for _ in range(5)]   # more than reserved: grows on demand
    assert pool['made'][cls] == 6 and len(pool['free'][cls]) == 0
    del held, b
    assert len(pool['free'][cls]) == 6
    _, c = t._pinned_out((1024, 2048, 3))
    assert pool["made"][6 << 20] == 1 and c.nbytes == 1024 * 2048 * 3


def test_native_sampler_randomized_sweep(libpath):
    """40 random (frame size up to 1100x2100, 0..11 boxes, scores with rejected boxes, op set, mixture shape)
    batches of two images: the native sampler's plan blob and final np.random state equal the Python statement of the
    draw order byte for byte, and both raise together when no box fits."""
    from oadg_b200.oamix import OAMix
    rng = np.random.RandomState(9)
    for k in range(40):
        version = 'augmix.all' if k % 2 else 'augmix'
        h, w = int(rng.randint(16, 1100)), int(rng.randint(16, 2100))
        n_gt, seed = int(rng.randint(0, 12)), int(rng.randint(0, 1 << 30))
        extra = {} if k % 3 else dict(mixture_width=int(rng.randint(1, 5)), mixture_depth=int(rng.choice([-1, 1, 2, 3])))
        t = OAMix(version=version, **extra)
        gts = []
        for _ in range(2):
            g = np.zeros((n_gt, 4), np.float32)
            for i in range(n_gt):
                x1, y1 = rng.uniform(0, w - 8), rng.uniform(0, h - 8)
                g[i] = [x1, y1, min(w, x1 + rng.uniform(4, w / 2)), min(h, y1 + rng.uniform(4, h / 2))]
            gts.append(g)
        scores = [[np.float64(rng.uniform(0, 1)) if rng.rand() > 0.2 else -1 for _ in range(n_gt)] for _ in range(2)]
        np.random.seed(seed)
        err_py = blob_py = None
        try:
            jobs = []
            for i, g in enumerate(gts):
                vp = t._sample_head(h, w, g)
                t._sample_tail(vp, g, scores[i])
                jobs.append((vp, g, i))
            blob_py = t._pack(jobs)
        except ValueError as e:
            err_py = e
        st_py = np.random.get_state()
        np.random.seed(seed)
        err_c = plan = None
        try:
            plan = t.sample_plan([(h, w)] * 2, gts, scores)
        except ValueError as e:
            err_c = e
        st_c = np.random.get_state()
        assert (err_py is None) == (err_c is None), (k, err_py, err_c)
        if err_py is None:
            assert np.array_equal(blob_py, plan.blob), k
            assert st_py[2] == st_c[2] and np.array_equal(st_py[1], st_c[1]), k

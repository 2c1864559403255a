"""mmcv-free work-alikes of the MMDetection-2.x plugin surface used by this path.

``Registry`` / ``build_from_cfg`` follow mmcv.utils.registry (third-party,
mmcv-full 1.3.17-1.5.0, reference mmdet/__init__.py:19-20); ``Config.fromfile``
follows mmcv.Config for the features the reference configs use: python files,
``_base_`` lists, recursive dict merge, ``_delete_``, ``custom_imports`` and a remap
of the reference's absolute ``/ws/external/`` base paths
(configs/OA-DG/cityscapes/faster_rcnn_r50_fpn_1x_cityscapes_oadg.py:2).

When the real mmcv / mmdet are importable the product classes register into the
real registries instead; this module is only the stand-in.
"""
import copy
import importlib
import inspect
import os


class Registry:
    def __init__(self, name, parent=None):
        self._name = name
        self._module_dict = {}
        self.parent = parent

    @property
    def name(self):
        return self._name

    @property
    def module_dict(self):
        return self._module_dict

    def __len__(self):
        return len(self._module_dict)

    def __contains__(self, key):
        return self.get(key) is not None

    def __repr__(self):
        return '%s(name=%s, items=%s)' % (self.__class__.__name__, self._name, list(self._module_dict))

    def get(self, key):
        if key in self._module_dict:
            return self._module_dict[key]
        if self.parent is not None:
            return self.parent.get(key)
        return None

    def _register_module(self, module_class, module_name=None, force=False):
        if not inspect.isclass(module_class):
            raise TypeError('module must be a class, but got %s' % type(module_class))
        if module_name is None:
            module_name = module_class.__name__
        names = [module_name] if isinstance(module_name, str) else module_name
        for name in names:
            if not force and name in self._module_dict:
                raise KeyError('%s is already registered in %s' % (name, self.name))
            self._module_dict[name] = module_class

    def register_module(self, name=None, force=False, module=None):
        if not isinstance(force, bool):
            raise TypeError('force must be a boolean, but got %s' % type(force))
        if module is not None:
            self._register_module(module, name, force)
            return module

        def _register(cls):
            self._register_module(cls, name, force)
            return cls
        return _register

    def build(self, *args, **kwargs):
        return build_from_cfg(*args, **kwargs, registry=self)


def build_from_cfg(cfg, registry, default_args=None):
    if not isinstance(cfg, dict):
        raise TypeError('cfg must be a dict, but got %s' % type(cfg))
    if 'type' not in cfg and (default_args is None or 'type' not in default_args):
        raise KeyError('`cfg` or `default_args` must contain the key "type", but got %s\n%s' % (cfg, default_args))
    if not isinstance(registry, Registry):
        raise TypeError('registry must be a Registry object, but got %s' % type(registry))
    args = dict(cfg)
    if default_args is not None:
        for k, v in default_args.items():
            args.setdefault(k, v)
    obj_type = args.pop('type')
    if isinstance(obj_type, str):
        obj_cls = registry.get(obj_type)
        if obj_cls is None:
            raise KeyError('%s is not in the %s registry' % (obj_type, registry.name))
    elif inspect.isclass(obj_type):
        obj_cls = obj_type
    else:
        raise TypeError('type must be a str or valid type, but got %s' % type(obj_type))
    try:
        return obj_cls(**args)
    except Exception as e:  # same message shape as mmcv
        raise type(e)('%s: %s' % (obj_cls.__name__, e))


PIPELINES = Registry('pipeline')
MODELS = Registry('models')
LOSSES = MODELS


def build_loss(cfg):
    return LOSSES.build(cfg)


class ConfigDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def _to_cfgdict(v):
    if isinstance(v, dict):
        return ConfigDict({k: _to_cfgdict(x) for k, x in v.items()})
    if isinstance(v, list):
        return [_to_cfgdict(x) for x in v]
    if isinstance(v, tuple):
        return tuple(_to_cfgdict(x) for x in v)
    return v


def _merge(a, b):
    """mmcv Config._merge_a_into_b: a overrides b, dicts merge recursively, `_delete_` replaces."""
    b = copy.deepcopy(b)
    for k, v in a.items():
        if isinstance(v, dict) and k in b and isinstance(b[k], dict) and not v.get('_delete_', False):
            b[k] = _merge(v, b[k])
        else:
            if isinstance(v, dict):
                v = {kk: vv for kk, vv in v.items() if kk != '_delete_'}
            b[k] = copy.deepcopy(v)
    return b


class Config:
    """``Config.fromfile(path, base_remap={'/ws/external/': '<reference or repo root>/'})``."""

    def __init__(self, cfg_dict, filename=None):
        object.__setattr__(self, '_cfg_dict', _to_cfgdict(cfg_dict))
        object.__setattr__(self, 'filename', filename)

    def __getattr__(self, k):
        return getattr(self._cfg_dict, k)

    def __getitem__(self, k):
        return self._cfg_dict[k]

    def __contains__(self, k):
        return k in self._cfg_dict

    def get(self, k, default=None):
        return self._cfg_dict.get(k, default)

    @staticmethod
    def _remap(path, base_remap):
        for src, dst in (base_remap or {}).items():
            if path.startswith(src):
                return os.path.join(dst, path[len(src):])
        return path

    @staticmethod
    def _file2dict(filename, base_remap):
        filename = Config._remap(filename, base_remap)
        if not os.path.isfile(filename):
            raise FileNotFoundError(filename)
        scope = {'__file__': filename}
        with open(filename) as fh:
            code = compile(fh.read(), filename, 'exec')
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):  # some reference configs print()
            exec(code, scope)
        cfg = {k: v for k, v in scope.items()
               if not k.startswith('__') and not inspect.ismodule(v) and not inspect.isfunction(v)
               and not inspect.isclass(v)}
        base = cfg.pop('_base_', [])
        base = [base] if isinstance(base, str) else list(base)
        merged = {}
        for b in base:
            if not os.path.isabs(b):
                b = os.path.join(os.path.dirname(filename), b)
            bd = Config._file2dict(b, base_remap)
            dup = set(merged) & set(bd)
            if dup:
                raise KeyError('Duplicate key is not allowed among bases: %s' % dup)
            merged.update(bd)
        return _merge(cfg, merged)

    @staticmethod
    def fromfile(filename, base_remap=None, import_custom_modules=True):
        cfg = Config._file2dict(filename, base_remap)
        if import_custom_modules and cfg.get('custom_imports'):
            ci = cfg['custom_imports']
            for mod in ci.get('imports', []):
                try:
                    importlib.import_module(mod)
                except ImportError:
                    if not ci.get('allow_failed_imports', False):
                        raise
        return Config(cfg, filename)


class Compose:
    """mmdet Compose (reference pipelines/compose.py:9-44) over this registry."""

    def __init__(self, transforms):
        self.transforms = []
        for t in transforms:
            if isinstance(t, dict):
                self.transforms.append(build_from_cfg(t, PIPELINES))
            elif callable(t):
                self.transforms.append(t)
            else:
                raise TypeError('transform must be callable or a dict')

    def __call__(self, data):
        for t in self.transforms:
            data = t(data)
            if data is None:
                return None
        return data

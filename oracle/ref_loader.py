"""TEST INFRASTRUCTURE — loader for the *unmodified* reference (``/root/reference``).

The reference (ShiqiYu/libfacedetection.train) hard-imports ``mmcv`` (``mmdet/__init__.py:2``)
which is not installable in this image.  This module installs a ``sys.meta_path`` finder that
serves permissive stub modules for ``mmcv*`` and a few other absent third-party packages, with
*real* implementations only for the handful of objects the YuNet hot path actually executes
(``Registry``/``build_from_cfg``/``ConfigDict``/``BaseModule``/``batched_nms`` ...).  Every
hot-path file then runs verbatim from ``/root/reference`` — nothing is copied.

Only usable where ``/root/reference`` exists (the development container).  It is used by
``oracle/gen_golden.py`` to pin the portable restatement in ``oracle/yunet_oracle.py`` and to
produce the committed fixtures under ``tests/golden``.  Nothing in the product path, in
``-m gpu`` tests, ``smoke()`` or ``bench.py`` imports this file.
"""
import importlib.abc
import importlib.machinery
import inspect
import os
import runpy
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get('YUNET_REFERENCE_ROOT', '/root/reference')
_STUB_ROOTS = ('mmcv', 'terminaltables', 'pycocotools', 'onnx', 'onnxruntime', 'lvis',
               'cityscapesscripts', 'panopticapi', 'albumentations', 'imagecorruptions',
               'onnxsim', 'tensorrt')


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'mmdet'))


# --------------------------------------------------------------------------- real objects
class ConfigDict(dict):
    """dict with attribute access (mmcv.utils.ConfigDict behaviour needed by the path)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value


def to_config(obj):
    if isinstance(obj, dict):
        return ConfigDict({k: to_config(v) for k, v in obj.items()})
    if isinstance(obj, list):
        return [to_config(v) for v in obj]
    if isinstance(obj, tuple):
        return tuple(to_config(v) for v in obj)
    return obj


def build_from_cfg(cfg, registry, default_args=None):
    args = dict(cfg)
    if default_args is not None:
        for k, v in default_args.items():
            args.setdefault(k, v)
    obj_type = args.pop('type')
    if isinstance(obj_type, str):
        obj_cls = registry.get(obj_type)
        if obj_cls is None:
            raise KeyError(f'{obj_type} is not in the {registry.name} registry')
    else:
        obj_cls = obj_type
    return obj_cls(**args)


class Registry:

    def __init__(self, name, build_func=None, parent=None, scope=None):
        self._name = name
        self._module_dict = {}
        self._children = {}
        self.parent = parent
        self.scope = scope
        if build_func is None:
            build_func = parent.build_func if parent is not None else build_from_cfg
        self.build_func = build_func

    name = property(lambda self: self._name)
    module_dict = property(lambda self: self._module_dict)

    def __len__(self):
        return len(self._module_dict)

    def __contains__(self, key):
        return self.get(key) is not None

    def get(self, key):
        if key in self._module_dict:
            return self._module_dict[key]
        if self.parent is not None:
            return self.parent.get(key)
        return None

    def build(self, *args, **kwargs):
        return self.build_func(*args, **kwargs, registry=self)

    def _register(self, module, name=None, force=False):
        names = [name or module.__name__] if not isinstance(name, (list, tuple)) else name
        for n in names:
            self._module_dict[n] = module

    def register_module(self, name=None, force=False, module=None):
        if module is not None:
            self._register(module, name, force)
            return module

        def deco(cls):
            self._register(cls, name, force)
            return cls

        return deco


class BaseModule(nn.Module):

    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg

    def init_weights(self):
        pass


def _passthrough_decorator_factory(*dargs, **dkwargs):
    if len(dargs) == 1 and callable(dargs[0]) and not dkwargs:
        return dargs[0]

    def deco(fn):
        return fn

    return deco


def batched_nms(boxes, scores, idxs, nms_cfg, class_agnostic=False):
    """mmcv.ops.nms.batched_nms semantics (mmcv-full 1.3.17..1.6.0) on torchvision.ops.nms:
    class-offset trick, IoU without +1, suppress IoU > thr, keep in score order."""
    import torchvision
    nms_cfg_ = dict(nms_cfg)
    class_agnostic = nms_cfg_.pop('class_agnostic', class_agnostic)
    iou_thr = nms_cfg_.get('iou_threshold', nms_cfg_.get('iou_thr'))
    if class_agnostic:
        boxes_for_nms = boxes
    else:
        max_coordinate = boxes.max()
        offsets = idxs.to(boxes) * (max_coordinate + torch.tensor(1).to(boxes))
        boxes_for_nms = boxes + offsets[:, None]
    keep = torchvision.ops.nms(boxes_for_nms, scores, float(iou_thr))
    boxes = boxes[keep]
    scores = scores[keep]
    return torch.cat([boxes, scores[:, None]], -1), keep


def digit_version(version_str, length=4):
    out = []
    for p in str(version_str).split('+')[0].split('.')[:length]:
        num = ''.join(ch for ch in p if ch.isdigit())
        out.append(int(num) if num else 0)
    while len(out) < length:
        out.append(0)
    return tuple(out)


def is_tuple_of(seq, expected_type):
    return isinstance(seq, tuple) and all(isinstance(x, expected_type) for x in seq)


def to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


_MODELS = Registry('model')
_REAL = {
    '__version__': '1.6.0',
    'Registry': Registry,
    'build_from_cfg': build_from_cfg,
    'ConfigDict': ConfigDict,
    'Config': ConfigDict,
    'BaseModule': BaseModule,
    'ModuleList': nn.ModuleList,
    'Sequential': nn.Sequential,
    'ModuleDict': nn.ModuleDict,
    'force_fp32': _passthrough_decorator_factory,
    'auto_fp16': _passthrough_decorator_factory,
    'jit': _passthrough_decorator_factory,
    'batched_nms': batched_nms,
    'get_dist_info': lambda: (0, 1),
    'print_log': lambda *a, **k: None,
    'digit_version': digit_version,
    'TORCH_VERSION': torch.__version__,
    'is_tuple_of': is_tuple_of,
    'to_2tuple': to_2tuple,
    'MODELS': _MODELS,
}
for _n in ('HOOKS', 'PLUGIN_LAYERS', 'CONV_LAYERS', 'POSITIONAL_ENCODING', 'TRANSFORMER_LAYER',
           'TRANSFORMER_LAYER_SEQUENCE', 'ATTENTION', 'FEEDFORWARD_NETWORK', 'DROPOUT_LAYERS',
           'ACTIVATION_LAYERS', 'NORM_LAYERS', 'PADDING_LAYERS', 'UPSAMPLE_LAYERS', 'RUNNERS',
           'OPTIMIZERS', 'OPTIMIZER_BUILDERS', 'PIPELINES', 'DATASETS'):
    _REAL[_n] = Registry(_n.lower())


class _PlaceholderMeta(type):

    def __getattr__(cls, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return _Placeholder


class _Placeholder(metaclass=_PlaceholderMeta):
    """Subclassable / instantiable / usable-as-decorator stand-in for anything unused."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return self

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return _Placeholder()

    @staticmethod
    def register_module(*a, **k):
        return lambda cls: cls


class _StubModule(types.ModuleType):

    def __getattr__(self, name):
        if name in _REAL:
            return _REAL[name]
        if name.startswith('__'):
            raise AttributeError(name)
        return _Placeholder


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split('.')[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


_installed = False


def install():
    """Install the stub finder and put the reference tree on ``sys.path`` (idempotent)."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f'reference tree not found at {REFERENCE_ROOT}; the reference loader '
                           'only works in the development container')
    sys.meta_path.insert(0, _StubFinder())
    sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def load_config(name):
    """``configs/yunet_{n,s}.py`` → nested ConfigDict (plain-Python config files)."""
    cfg = runpy.run_path(os.path.join(REFERENCE_ROOT, 'configs', f'{name}.py'))
    return to_config({k: v for k, v in cfg.items() if not k.startswith('_') and
                      not inspect.ismodule(v)})


def build_reference_model(name='yunet_n', pretrained=True, seed=0):
    """Build the reference detector exactly as ``tools/train.py:207`` does."""
    install()
    import warnings
    warnings.filterwarnings('ignore')
    from mmdet.models import build_detector
    cfg = load_config(name)
    torch.manual_seed(seed)
    model = build_detector(cfg.model)
    if pretrained:
        ck = torch.load(os.path.join(REFERENCE_ROOT, 'weights', f'{name}.pth'),
                        map_location='cpu', weights_only=False)
        model.load_state_dict(ck['state_dict'], strict=True)
    return model, cfg

"""ctypes binding of ``libyunet_b200.so`` (C ABI declared in ``include/yunet_b200.h``).

The product path has no CPU fallback: if the shared library is missing, importing this module
raises; if a compute entry point is called without a CUDA device it raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libyunet_b200.so')

MAX_STAGES = 8
PRED_CH = 16
GT_ROW = 19


class ArchCfg(C.Structure):
    _fields_ = [('num_stages', C.c_int),
                ('stage_channels', (C.c_int * 3) * MAX_STAGES),
                ('downsample_mask', C.c_int),
                ('out_idx', C.c_int * 3),
                ('shared_stacked_convs', C.c_int),
                ('feat_channels', C.c_int),
                ('num_classes', C.c_int),
                ('kps_num', C.c_int),
                ('strides', C.c_int * 3)]


class LossCfg(C.Structure):
    _fields_ = [('center_radius', C.c_float), ('candidate_topk', C.c_int),
                ('iou_weight', C.c_float), ('cls_weight', C.c_float),
                ('loss_cls_weight', C.c_float), ('loss_bbox_weight', C.c_float),
                ('loss_obj_weight', C.c_float), ('loss_kps_weight', C.c_float),
                ('eiou_smooth_point', C.c_float), ('eiou_eps', C.c_float),
                ('smooth_l1_beta', C.c_float)]


class UnitDesc(C.Structure):
    _fields_ = [('name', C.c_char * 96), ('cin', C.c_int), ('cout', C.c_int), ('mode', C.c_int),
                ('in_a', C.c_int), ('in_b', C.c_int), ('out', C.c_int), ('div', C.c_int),
                ('has_bn', C.c_int), ('acc_a', C.c_int), ('acc_b', C.c_int),
                ('bn_out', C.c_int), ('bn_a', C.c_int), ('bn_b', C.c_int),
                ('pred_level', C.c_int),
                ('w1', C.c_longlong), ('b1', C.c_longlong), ('w2', C.c_longlong),
                ('b2', C.c_longlong), ('gamma', C.c_longlong), ('beta', C.c_longlong)]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; '
            f'g.build()"` (or `make -C libfacedetection/train_b200/csrc`). There is no CPU fallback.')
    lib = C.CDLL(LIB_PATH)
    vp, ci, cf, cs, ll = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_longlong
    P = C.POINTER
    sigs = {
        'yunet_ctx_create': (ci, [P(ArchCfg), P(vp)]),
        'yunet_ctx_destroy': (None, [vp]),
        'yunet_last_error': (C.c_char_p, [vp]),
        'yunet_version': (C.c_char_p, []),
        'yunet_num_params': (ll, [vp]),
        'yunet_param_count': (ci, [vp]),
        'yunet_param_info': (ci, [vp, ci, C.c_char_p, ci, P(ll), P(ci), P(ci)]),
        'yunet_num_bn_channels': (ll, [vp]),
        'yunet_bn_count': (ci, [vp]),
        'yunet_bn_info': (ci, [vp, ci, C.c_char_p, ci, P(ll), P(ci)]),
        'yunet_num_priors': (ci, [vp, ci, ci]),
        'yunet_workspace_bytes': (cs, [vp, ci, ci, ci, ci]),
        'yunet_grid_priors': (ci, [vp, ci, ci, vp, vp]),
        'yunet_forward': (ci, [vp, vp, vp, vp, ci, ci, ci, ci, cf, vp, vp, cs, vp]),
        'yunet_assign_workspace_bytes': (cs, [vp, ci, ci, ci]),
        'yunet_simota_assign': (ci, [vp, P(LossCfg), vp, vp, vp, ci, ci, ci, vp, vp, vp, vp, cs, vp]),
        'yunet_simota_assign_ext': (ci, [vp, P(LossCfg), ci, vp, vp, vp, vp, vp, vp, vp, vp, vp, cs, vp]),
        'yunet_loss_grad': (ci, [vp, P(LossCfg), vp, vp, vp, vp, vp, vp, vp, P(cf), ci, ci, ci,
                                 vp, vp, vp]),
        'yunet_backward': (ci, [vp, vp, vp, vp, ci, ci, ci, vp, vp, cs, vp]),
        'yunet_sgd_step': (ci, [vp, vp, vp, vp, ll, cf, cf, cf, cf, vp]),
        'yunet_sgd_step_dev': (ci, [vp, vp, vp, vp, ll, vp, cf, cf, cf, vp]),
        'yunet_nms_workspace_bytes': (cs, [vp, ci, ci, ci]),
        'yunet_decode_nms': (ci, [vp, vp, ci, ci, ci, cf, cf, vp, ci, vp, vp, vp, vp, cs, vp]),
        'yunet_preprocess_u8': (ci, [vp, vp, vp, vp, vp, ci, ci, cf, vp, vp]),
        'yunet_unit_count': (ci, [vp]),
        'yunet_unit_get': (ci, [vp, ci, P(UnitDesc)]),
        'yunet_read_activation': (ci, [vp, ci, vp, vp, ci, ci, ci, ci, vp, vp, vp]),
        'yunet_launch_count': (ll, [vp]),
        'yunet_set_option': (ci, [vp, C.c_char_p, ci]),
        'yunet_ws_offset': (ll, [vp, ci, ci, ci, ci, ci, ci]),
        'yunet_profile_begin': (ci, [vp]),
        'yunet_profile_end': (ci, [vp]),
        'yunet_profile_get': (ci, [vp, ci, C.c_char_p, ci, P(cf), P(C.c_double)]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)      # AttributeError here == header/library mismatch
        fn.restype = res
        fn.argtypes = args
    return lib, sorted(sigs)


lib, EXPORTED = _load()


class YuNetError(RuntimeError):
    pass


def check(ctx, code, what):
    if code != 0:
        msg = lib.yunet_last_error(ctx)
        raise YuNetError(f'{what} failed ({code}): {msg.decode() if msg else "?"}')


def default_loss_cfg():
    """configs/yunet_n.py:122-138 + sim_ota_assigner.py:28-36 + iou_loss.py:535-544."""
    return LossCfg(2.5, 10, 3.0, 1.0, 1.0, 5.0, 1.0, 0.1, 0.1, 1e-6, 0.1111111111111111)


def make_arch_cfg(stage_channels, downsample_idx, out_idx, shared_stacked_convs, feat_channels=64,
                  num_classes=1, kps_num=5, strides=(8, 16, 32)):
    cfg = ArchCfg()
    cfg.num_stages = len(stage_channels)
    for i, sc in enumerate(stage_channels):
        for j in range(3):
            cfg.stage_channels[i][j] = sc[j] if j < len(sc) else 0
    mask = 0
    for i in downsample_idx:
        mask |= 1 << i
    cfg.downsample_mask = mask
    for k in range(3):
        cfg.out_idx[k] = out_idx[k]
        cfg.strides[k] = strides[k]
    cfg.shared_stacked_convs = shared_stacked_convs
    cfg.feat_channels = feat_channels
    cfg.num_classes = num_classes
    cfg.kps_num = kps_num
    return cfg


class Ctx:
    """Owns a ``yunet_ctx*`` and exposes the (host-only) plan queries."""

    def __init__(self, arch_cfg):
        self._h = C.c_void_p()
        code = lib.yunet_ctx_create(C.byref(arch_cfg), C.byref(self._h))
        if code != 0:
            msg = lib.yunet_last_error(self._h) if self._h else b'allocation failed'
            if self._h:
                lib.yunet_ctx_destroy(self._h)
            self._h = None
            raise YuNetError(f'yunet_ctx_create failed ({code}): {msg.decode()}')
        self.num_params = lib.yunet_num_params(self._h)
        self.num_bn_channels = lib.yunet_num_bn_channels(self._h)

    def __del__(self):
        h = getattr(self, '_h', None)
        if h and lib is not None:      # module globals may already be gone at interpreter exit
            lib.yunet_ctx_destroy(h)
            self._h = None

    @property
    def handle(self):
        return self._h

    def params(self):
        """[(state_dict key, offset, shape)] in bucket order."""
        out = []
        name = C.create_string_buffer(160)
        off = C.c_longlong()
        nd = C.c_int()
        shape = (C.c_int * 4)()
        for i in range(lib.yunet_param_count(self._h)):
            lib.yunet_param_info(self._h, i, name, 160, C.byref(off), C.byref(nd), shape)
            out.append((name.value.decode(), off.value, tuple(shape[k] for k in range(nd.value))))
        return out

    def bns(self):
        """[(state_dict prefix, channel offset, channels)]."""
        out = []
        name = C.create_string_buffer(160)
        off = C.c_longlong()
        ch = C.c_int()
        for i in range(lib.yunet_bn_count(self._h)):
            lib.yunet_bn_info(self._h, i, name, 160, C.byref(off), C.byref(ch))
            out.append((name.value.decode(), off.value, ch.value))
        return out

    def units(self, include_stem=False):
        out = []
        for i in range(-1 if include_stem else 0, lib.yunet_unit_count(self._h)):
            d = UnitDesc()
            lib.yunet_unit_get(self._h, i, C.byref(d))
            out.append(d)
        return out

    def num_priors(self, H, W):
        return lib.yunet_num_priors(self._h, H, W)

    def workspace_bytes(self, B, H, W, train):
        return lib.yunet_workspace_bytes(self._h, B, H, W, 1 if train else 0)

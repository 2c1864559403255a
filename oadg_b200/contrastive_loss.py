"""OA-Loss (instance-level contrastive loss) on B200.

Drop-in for the reference ``@LOSSES.register_module() class ContrastiveLossPlus``
(mmdet/models/losses/oadg/contrastive_loss_plus.py:10-50) and the ``supcontrast``
function it calls (contrastive_loss.py:170-232): same constructor keys
(``loss_weight, temperature, num_views, normalized_input, min_samples, **kwargs``),
same call ``loss_cont(cont_feats, labels)``, same quirks.
The arithmetic is the fused CUDA path behind ``oadg_supcon_forward/backward``
(include/oadg.h): no N x N tensor is materialised and there is no host sync.
"""
import ctypes

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .registry import LOSSES

ORI_SIZE = 1024  # contrastive_loss.py:190 hard-codes 512 RoIs x 2 images per view


def reference_pair_map(n, ori_size=ORI_SIZE):
    """pair[i] = row of the other view of RoI i under the reference's hard-wired layout
    (contrastive_loss.py:190-208): rows [0,1024) <-> [1024,2048); random-proposal block
    [2048, 2048+rp) <-> [2048+rp, 2048+2rp), rp = (n % 1024)//2; -1 elsewhere."""
    if n < 2 * ori_size:
        # the reference fails at contrastive_loss.py:205 assigning a 1024x1024 eye into a smaller slice
        raise RuntimeError(
            'The expanded size of the tensor must match the existing size: supcontrast needs at least '
            '%d rows (512 RoIs x 2 images x 2 views), got %d' % (2 * ori_size, n))
    rp = (n % ori_size) // 2
    pair = np.full(n, -1, np.int32)
    i = np.arange(ori_size)
    pair[i] = i + ori_size
    pair[i + ori_size] = i
    j = np.arange(rp)
    pair[2 * ori_size + j] = 2 * ori_size + rp + j
    pair[2 * ori_size + rp + j] = 2 * ori_size + j
    return pair


def yolo_pair_map(n):
    """The two-view layout of ``supcontrast_yolo`` (contrastive_loss.py:234-299): ``ori_size = n // 2`` rows per view,
    row i <-> row i + n // 2; a trailing odd row has no other view (its ``rp_size`` is 0)."""
    half = n // 2
    pair = np.full(n, -1, np.int32)
    i = np.arange(half)
    pair[i] = i + half
    pair[i + half] = i
    return pair


MIN_TEMPERATURE = 0.025
_PAIR_CACHE = {}


def _pair_tensor(n, device):
    key = (n, str(device))
    if key not in _PAIR_CACHE:
        _PAIR_CACHE[key] = torch.from_numpy(reference_pair_map(n)).to(device)
    return _PAIR_CACHE[key]


class _SupConFn(torch.autograd.Function):
    """loss = w * supcontrast(normalize(x), y); backward through libOADG."""

    @staticmethod
    def forward(ctx, feats, labels, pair, temperature, loss_weight, min_samples, normalized_input, stats):
        lib = _lib.load()
        feats = feats.contiguous()
        n, c = feats.shape
        need = ctypes.c_size_t(0)
        _lib.check(lib.oadg_supcon_workspace_bytes(n, c, ctypes.byref(need)))
        ws = torch.empty(need.value + 256, dtype=torch.uint8, device=feats.device)
        base = (ws.data_ptr() + 255) // 256 * 256
        loss = torch.empty((), dtype=torch.float32, device=feats.device)
        nl = ctypes.c_int(0)
        s = _lib.raw_stream(feats.device)
        _lib.check(lib.oadg_supcon_forward(feats.data_ptr(), labels.data_ptr(), pair.data_ptr(), n, c,
                                           float(temperature), float(loss_weight), int(min_samples),
                                           int(bool(normalized_input)), loss.data_ptr(), base, need.value,
                                           ctypes.byref(nl), s))
        stats['launches'] = stats.get('launches', 0) + nl.value
        ctx.save_for_backward(feats, labels, pair, ws)
        ctx.cfg = (float(temperature), float(loss_weight), int(bool(normalized_input)), need.value, stats)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        feats, labels, pair, ws = ctx.saved_tensors
        temperature, loss_weight, normalized_input, need, stats = ctx.cfg
        n, c = feats.shape
        base = (ws.data_ptr() + 255) // 256 * 256
        g = grad_out.to(torch.float32).contiguous()
        gx = torch.empty_like(feats)
        nl = ctypes.c_int(0)
        s = _lib.raw_stream(feats.device)
        _lib.check(lib.oadg_supcon_backward(feats.data_ptr(), labels.data_ptr(), pair.data_ptr(), n, c,
                                            temperature, loss_weight, normalized_input, g.data_ptr(),
                                            gx.data_ptr(), base, need, ctypes.byref(nl), s))
        stats['launches'] = stats.get('launches', 0) + nl.value
        return gx, None, None, None, None, None, None, None


def supcontrast(logits_clean, labels=None, num_views=2, lambda_weight=0.1, temper=0.07, min_samples=10,
                pair=None, loss_weight=1.0, normalized_input=False, stats=None):
    """Reference ``supcontrast`` (contrastive_loss.py:170-232) on CUDA features.

    ``pair`` (int32 [N], optional) generalises the reference's hard-wired two-view layout."""
    assert num_views == 2, "Only num_views 2 and batch_size 2 case are supported."
    _lib.require_cuda()
    if not logits_clean.is_cuda:
        raise _lib.OADGError('supcontrast: features must live on a CUDA device (no CPU fallback)')
    if not temper >= MIN_TEMPERATURE:
        raise ValueError('supcontrast: temperature %r is below %g, the bound of the fixed-shift exponent of the '
                         'tcgen05 forward (oaloss.cu kMinTemperatureTc); the reference configs use 0.06 / 0.07'
                         % (temper, MIN_TEMPERATURE))
    n = logits_clean.shape[0]
    # the kernels read raw pointers: labels and pair must be dense, on the features' device, of the declared dtype
    labels = labels.reshape(-1).to(device=logits_clean.device, dtype=torch.int64).contiguous()
    if labels.shape[0] != n:
        raise ValueError('supcontrast: %d labels for %d rows' % (labels.shape[0], n))
    if pair is None:
        pair = _pair_tensor(n, logits_clean.device)
    else:
        if not torch.is_tensor(pair):
            pair = torch.as_tensor(np.asarray(pair))
        if pair.dim() != 1 or pair.shape[0] != n or pair.is_floating_point():
            raise ValueError('supcontrast: pair must be an integer vector of %d entries' % n)
        pair = pair.to(device=logits_clean.device, dtype=torch.int32).contiguous()
    feats = logits_clean if logits_clean.dtype == torch.float32 else logits_clean.float()
    return _SupConFn.apply(feats, labels, pair, temper, loss_weight, min_samples, normalized_input,
                           stats if stats is not None else {})


def supcontrast_yolo(logits_clean, labels=None, num_views=2, lambda_weight=0.1, temper=0.07, min_samples=10):
    """Reference ``supcontrast_yolo`` (contrastive_loss.py:234-299; called by dense_heads/yolo_head_cont.py:463 on the
    sampled grid cells of both views): the same masked InfoNCE with the view boundary at N // 2.  Same kernels, other
    pair map."""
    n = logits_clean.shape[0]
    key = ('yolo', n, str(logits_clean.device))
    if key not in _PAIR_CACHE:
        _PAIR_CACHE[key] = torch.from_numpy(yolo_pair_map(n)).to(logits_clean.device)
    return supcontrast(logits_clean, labels, num_views=num_views, lambda_weight=lambda_weight, temper=temper,
                       min_samples=min_samples, pair=_PAIR_CACHE[key])


@LOSSES.register_module()
class ContrastiveLossPlus(nn.Module):

    def __init__(self,
                 loss_weight=1,
                 temperature=0.07,
                 num_views=2,
                 normalized_input=True,
                 min_samples=10,
                 **kwargs):
        """ContrastiveLossPlus (contrastive_loss_plus.py:13-29)."""
        super(ContrastiveLossPlus, self).__init__()
        self.loss_weight = loss_weight
        self.temperature = temperature
        self.num_views = num_views
        self.normalized_input = normalized_input
        self.min_samples = min_samples
        self.kwargs = kwargs
        self.loss = supcontrast
        self.stats = {}

    def forward(self, cont_feats, labels, pair=None):
        """cont_feats [N, 256] float, labels [M, 1] int64 (M <= N) -> 0-dim loss tensor."""
        if len(cont_feats) == 0:
            return torch.zeros(1)  # contrastive_loss_plus.py:38-39 (CPU tensor, shape [1])
        if len(cont_feats) != len(labels):  # random proposal case, contrastive_loss_plus.py:44-47
            random_proposal_len = len(cont_feats) - len(labels)
            random_proposal_targets = labels[-1:, :].expand(random_proposal_len, -1)   # a view: one kernel (the cat)
            labels = torch.cat([labels, random_proposal_targets], dim=0)
        # the two F.normalize calls (contrastive_loss_plus.py:41, contrastive_loss.py:155) are fused
        # into the kernel; loss_weight is applied inside it as well
        return self.loss(cont_feats, labels, num_views=self.num_views, temper=self.temperature,
                         min_samples=self.min_samples, pair=pair, loss_weight=self.loss_weight,
                         normalized_input=self.normalized_input, stats=self.stats)

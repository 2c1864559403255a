"""The JSD consistency half of OA-Loss and the first-view regression losses.

Drop-ins for the reference's ``@LOSSES.register_module()`` classes, same constructor keys (so that
``configs/OA-DG/**`` build them unchanged) and same call signatures:

* ``CrossEntropyLossPlus`` (mmdet/models/losses/oadg/cross_entropy_loss_plus.py:322-500): the classification term sees
  only the FIRST view's chunk of the predictions / labels / weights (:40-56 softmax, :115-128 sigmoid) but is averaged
  by the all-views ``avg_factor`` (``avg='1.0'``); ``additional_loss='jsdv1_3_2aug'`` (:264-319) adds
  ``lambda_weight`` x the Jensen-Shannon divergence between the two views' class distributions -- softmax for the RoI
  head, (sigmoid, 1 - sigmoid) for the single-logit RPN head -- summed over rows and classes (the ``/ len(p_aug1)`` at
  :310 acts on a tensor reshaped to [1, n, C], so it divides by 1, not by the rows of a view), and then (the reference
  routes the scalar through ``weight_reduce_loss``, :316-317) divided by ``avg_factor``.
* ``SmoothL1LossPlus`` / ``L1LossPlus`` (smooth_l1_loss_plus.py:12-62,350-552, losses/utils.py:106-151): element-wise loss
  on the first view's chunk only, weighted by the first chunk of the weights, reduced with the all-views ``avg_factor``.

The JSD term runs as one fused CUDA kernel pair behind ``oadg_jsd2_forward`` when the logits live on a GPU (probabilities,
clamped mixture, both KL terms, row sum and the gradient with respect to both views' logits in one pass); the same
closed form in torch serves CPU tensors (the reference's own op sequence) and is the test oracle for the kernel.
"""
import ctypes

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .registry import LOSSES


def _weight_reduce(loss, weight=None, reduction='mean', avg_factor=None):
    """losses/utils.py:30-57."""
    if weight is not None:
        loss = loss * weight
    if avg_factor is None:
        if reduction == 'mean':
            return loss.mean()
        return loss.sum() if reduction == 'sum' else loss
    if reduction == 'mean':
        return loss.sum() / avg_factor
    if reduction != 'none':
        raise ValueError('avg_factor can not be used with reduction="sum"')
    return loss


def _first_chunk(t, num_views):
    return torch.chunk(t, num_views)[0] if t is not None else None


def jsd_two_views_torch(pred):
    """cross_entropy_loss_plus.py:288-310: the divergence summed over rows and classes (the reference's
    ``/ len(p_aug1)`` divides by the leading 1 of its [1, n, C] reshape): a 0-dim tensor."""
    a, b = torch.chunk(pred, 2)
    if a.shape[-1] == 1:
        sa, sb = torch.sigmoid(a), torch.sigmoid(b)
        p, q = torch.cat((sa, 1 - sa), dim=1), torch.cat((sb, 1 - sb), dim=1)
    else:
        p, q = F.softmax(a, dim=1), F.softmax(b, dim=1)
    log_m = torch.clamp((p + q) / 2., 1e-7, 1).log()
    loss = (F.kl_div(log_m, p, reduction='none') + F.kl_div(log_m, q, reduction='none')) / 2.
    return loss.sum()


_JSD_SCRATCH = {}


class _Jsd2Fn(torch.autograd.Function):
    """The same number from one fused kernel (oadg_jsd2_forward): loss and d loss / d logits in a single pass."""

    @staticmethod
    def forward(ctx, pred):
        lib = _lib.load()
        pred = pred.contiguous()
        n2, c = pred.shape
        loss = torch.empty((), dtype=torch.float32, device=pred.device)
        grad = torch.empty_like(pred)
        key = str(pred.device)
        scratch = _JSD_SCRATCH.get(key)
        if scratch is None:
            scratch = _JSD_SCRATCH[key] = torch.zeros(lib.oadg_jsd2_scratch_bytes(), dtype=torch.uint8, device=pred.device)
        _lib.check(lib.oadg_jsd2_forward(pred.data_ptr(), n2 // 2, c, loss.data_ptr(), grad.data_ptr(),
                                         scratch.data_ptr(), _lib.raw_stream(pred.device)))
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g


def jsd_two_views(pred):
    if pred.is_cuda and pred.dtype == torch.float32 and pred.dim() == 2 and pred.shape[0] % 2 == 0 and \
            1 <= pred.shape[1] <= 32 and pred.shape[0] > 0:
        return _Jsd2Fn.apply(pred)
    return jsd_two_views_torch(pred)


def jsdv1_3_2aug(pred, label=None, weight=None, reduction='mean', avg_factor=None, **kwargs):
    """cross_entropy_loss_plus.py:264-319.  ``weight`` (first chunk) multiplies the SCALAR, as in the reference."""
    loss = jsd_two_views(pred)
    if weight is not None:
        weight = torch.chunk(weight, 2)[0].float()
    return _weight_reduce(loss, weight=weight, reduction=reduction, avg_factor=avg_factor)


def _expand_onehot(labels, weights, channels, ignore_index):
    """mmdet losses/cross_entropy_loss.py _expand_onehot_labels (the sigmoid path's label layout)."""
    bin_labels = labels.new_full((labels.size(0), channels), 0)
    valid = (labels >= 0) & (labels != ignore_index)
    inds = torch.nonzero(valid & (labels < channels), as_tuple=False)
    if inds.numel() > 0:
        bin_labels[inds, labels[inds]] = 1
    valid = valid.view(-1, 1).expand(labels.size(0), channels).float()
    bin_weights = valid if weights is None else weights.view(-1, 1).repeat(1, channels) * valid
    return bin_labels, bin_weights


@LOSSES.register_module()
class CrossEntropyLossPlus(nn.Module):

    def __init__(self, use_sigmoid=False, use_mask=False, reduction='mean', class_weight=None, ignore_index=None,
                 loss_weight=1.0, additional_loss='jsd', additional_loss_weight_reduce=False, lambda_weight=0.0001,
                 additional_loss2=None, lambda_weight2=0.0001, kpositive=3, classes=9, temper=1, temper_ratio=1.0,
                 analysis=False, add_act=None, wandb_name=None, use_cls_weight=False, add_class_weight=None,
                 num_views=3, avg='1.0', **kwargs):
        super().__init__()
        assert (use_sigmoid is False) or (use_mask is False)
        if use_mask:
            raise NotImplementedError('mask_cross_entropy is outside the OA-DG hot path')
        self.use_sigmoid, self.reduction, self.class_weight = use_sigmoid, reduction, class_weight
        self.ignore_index, self.loss_weight = ignore_index, loss_weight
        self.additional_loss, self.additional_loss_weight_reduce = additional_loss, additional_loss_weight_reduce
        self.lambda_weight, self.num_views, self.avg, self.wandb_name = lambda_weight, num_views, avg, wandb_name
        self.kwargs = kwargs
        # only the variant the shipped configs use is on the path; 'jsdv1_3' (three views) is not
        self.cls_additional = jsdv1_3_2aug if additional_loss == 'jsdv1_3_2aug' else None
        self.wandb_features = {}

    def forward(self, cls_score, label, weight=None, avg_factor=None, reduction_override=None, ignore_index=None,
                **kwargs):
        assert reduction_override in (None, 'none', 'mean', 'sum')
        reduction = reduction_override if reduction_override else self.reduction
        ignore_index = self.ignore_index if ignore_index is None else ignore_index
        ignore_index = -100 if ignore_index is None else ignore_index
        cw = cls_score.new_tensor(self.class_weight) if self.class_weight is not None else None
        af = kwargs.get('original_avg_factor', avg_factor)
        if af is not None and self.avg == '1.1':
            af = af / self.num_views
        nv = self.num_views
        if self.use_sigmoid:                                  # :84-131
            lab, w = label, weight
            if cls_score.dim() != lab.dim():
                lab, w = _expand_onehot(lab, w, cls_score.size(-1), ignore_index)
            w = _first_chunk(w, nv)
            w = w.float() if w is not None else None
            el = F.binary_cross_entropy_with_logits(_first_chunk(cls_score, nv), _first_chunk(lab, nv).float(),
                                                    pos_weight=cw, reduction='none')
        else:                                                 # :12-58
            el = F.cross_entropy(_first_chunk(cls_score, nv), _first_chunk(label, nv), weight=cw, reduction='none',
                                 ignore_index=ignore_index)
            w = _first_chunk(weight, nv)
            w = w.float() if w is not None else None
        loss_cls = self.loss_weight * _weight_reduce(el, weight=w, reduction=reduction, avg_factor=af)
        loss_additional = 0
        if self.cls_additional is not None:
            w2 = weight if self.additional_loss_weight_reduce else None
            loss_additional = self.cls_additional(cls_score, label, w2, reduction=reduction, avg_factor=avg_factor)
        self.wandb_features = {'ce_loss(%s)' % self.wandb_name: loss_cls,
                               'additional_loss(%s)' % self.wandb_name: loss_additional}
        return loss_cls + self.lambda_weight * loss_additional


class _FirstViewRegLoss(nn.Module):
    """smooth_l1_loss_plus.py: element-wise loss on the first view's chunk (losses/utils.py:106-151 weighted_loss2)."""

    def __init__(self, reduction='mean', loss_weight=1.0, additional_loss='jsd', lambda_weight=0.0001, wandb_name=None,
                 analysis=False, num_views=3, **kwargs):
        super().__init__()
        self.reduction, self.loss_weight, self.num_views = reduction, loss_weight, num_views
        self.additional_loss, self.lambda_weight, self.wandb_name = additional_loss, lambda_weight, wandb_name
        if additional_loss not in (None, 'None', 'none'):
            raise NotImplementedError('regression consistency terms are not used by the OA-DG configs')

    def element(self, pred, target):
        raise NotImplementedError

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None, **kwargs):
        assert reduction_override in (None, 'none', 'mean', 'sum')
        reduction = reduction_override if reduction_override else self.reduction
        p, t = _first_chunk(pred, self.num_views), _first_chunk(target, self.num_views)
        el = p.sum() * 0 if t.numel() == 0 else self.element(p, t)
        w = _first_chunk(weight, self.num_views)
        return self.loss_weight * _weight_reduce(el, w, reduction, avg_factor)


@LOSSES.register_module()
class SmoothL1LossPlus(_FirstViewRegLoss):

    def __init__(self, beta=1.0, **kwargs):
        super().__init__(**kwargs)
        assert beta > 0
        self.beta = beta

    def element(self, pred, target):
        diff = torch.abs(pred - target)
        return torch.where(diff < self.beta, 0.5 * diff * diff / self.beta, diff - 0.5 * self.beta)


@LOSSES.register_module()
class L1LossPlus(_FirstViewRegLoss):

    def element(self, pred, target):
        return torch.abs(pred - target)

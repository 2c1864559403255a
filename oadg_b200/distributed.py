"""OA-Loss across ranks: one all-gather of the RoI embeddings before the contrastive loss.

New capability: the reference computes ``loss_cont`` per rank on its own 2048(+rp) rows.  Here every
rank's rows are anchors against the embeddings of ALL ranks:

    fhat_r  = normalize(normalize(x_r))                      local kernel
    F_all   = all_gather(fhat_r), y_all = all_gather(y_r)    one collective over NVLink (2 MB per rank at N=2048)
    loss_r, stats_r = forward(F_all; anchors = rows of r)    tcgen05 similarity, rectangular [N, W*N]
    loss    = all_reduce_sum(loss_r)                         = mean over all W*N anchors
    stats   = all_gather(stats_r)                            16 bytes per row
    dx_r    = W * dL/dx_r,  dL/dfhat_r = (G + G^T)[rows of r, :] F_all / T

Positives: foreground rows of the same class on ANY rank; background rows only with their own other view (same
rank).  With the per-row statistics of every rank at hand, (G + G^T) restricted to the local rows is exact, so no
reduce-scatter of column gradients is needed.  The factor W compensates DDP's gradient averaging: every rank
reports the same global-mean loss, and sum_r J_r^T (W dL/dx_r) / W is the true gradient.

The collectives go through ``torch.distributed`` (NCCL on GPUs; gloo in the CPU tests, where the local compute is
replaced by a numpy stand-in to exercise exactly this plumbing).
"""
import ctypes

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .contrastive_loss import reference_pair_map


def gathered_pair_map(pair_local, world):
    """Global other-view index of every gathered row: rank r's block is offset by r * N."""
    pair_local = np.asarray(pair_local, dtype=np.int64)
    n = len(pair_local)
    out = [np.where(pair_local >= 0, pair_local + r * n, -1) for r in range(world)]
    return np.concatenate(out).astype(np.int32)


def _all_gather_cat(t, group):
    world = dist.get_world_size(group)
    if world == 1:
        return t
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t.contiguous(), group=group)
    return torch.cat(parts, dim=0)


class CudaBackend:
    """Local compute through libOADG (tcgen05 kernels)."""

    def __init__(self):
        self.launches = 0

    def _ws(self, n_total, c, device):
        lib = _lib.load()
        need = ctypes.c_size_t(0)
        _lib.check(lib.oadg_supcon_workspace_bytes(n_total, c, ctypes.byref(need)))
        key = (n_total, c, str(device))
        if getattr(self, '_ws_key', None) != key:   # one workspace per shape, reused every step
            self._ws_buf = torch.empty(need.value + 256, dtype=torch.uint8, device=device)
            self._ws_key = key
        ws = self._ws_buf
        return ws, (ws.data_ptr() + 255) // 256 * 256, need.value

    def normalize(self, x, n_total, normalized_input):
        _lib.require_cuda()
        lib = _lib.load()
        self.ws = self._ws(n_total, x.shape[1], x.device)
        fhat = torch.empty_like(x)
        s = _lib.raw_stream(x.device)
        _lib.check(lib.oadg_supcon_normalize(x.data_ptr(), x.shape[0], n_total, x.shape[1], int(normalized_input),
                                             fhat.data_ptr(), self.ws[1], self.ws[2], s))
        self.launches += 1
        return fhat

    def forward(self, f_all, labels_all, pair_all, row0, n_rows, temperature, loss_weight, min_samples):
        lib = _lib.load()
        loss = torch.empty((), dtype=torch.float32, device=f_all.device)
        stats = torch.empty(n_rows, 4, dtype=torch.float32, device=f_all.device)
        nl = ctypes.c_int(0)
        s = _lib.raw_stream(f_all.device)
        _lib.check(lib.oadg_supcon_forward_gathered(f_all.data_ptr(), labels_all.data_ptr(), pair_all.data_ptr(),
                                                    f_all.shape[0], row0, n_rows, f_all.shape[1], float(temperature),
                                                    float(loss_weight), int(min_samples), loss.data_ptr(),
                                                    stats.data_ptr(), self.ws[1], self.ws[2], ctypes.byref(nl), s))
        self.launches += nl.value
        return loss, stats

    def backward(self, x, f_all, labels_all, pair_all, stats_all, row0, temperature, normalized_input, grad):
        lib = _lib.load()
        gx = torch.empty_like(x)
        nl = ctypes.c_int(0)
        s = _lib.raw_stream(x.device)
        _lib.check(lib.oadg_supcon_backward_gathered(x.data_ptr(), labels_all.data_ptr(), pair_all.data_ptr(),
                                                     stats_all.data_ptr(), f_all.shape[0], row0, x.shape[0], x.shape[1],
                                                     float(temperature), int(normalized_input), grad.data_ptr(),
                                                     gx.data_ptr(), self.ws[1], self.ws[2], ctypes.byref(nl), s))
        self.launches += nl.value
        return gx


_PAIR_CACHE = {}


def _pair_all_on(device, pair_local, world):
    """The global pair map as a device tensor, cached per (layout, world, device): no per-step H2D copy."""
    pl = np.asarray(pair_local, dtype=np.int64)
    key = (pl.tobytes(), world, str(device))
    t = _PAIR_CACHE.get(key)
    if t is None:
        if len(_PAIR_CACHE) > 16:
            _PAIR_CACHE.clear()
        t = _PAIR_CACHE[key] = torch.from_numpy(gathered_pair_map(pl, world)).to(device)
    return t


def _all_gather_rows(t, group):
    """[n, k] -> [W * n, k] with ONE collective (all_gather_into_tensor where the backend has it)."""
    world = dist.get_world_size(group)
    if world == 1:
        return t
    t = t.contiguous()
    out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    try:
        dist.all_gather_into_tensor(out, t, group=group)
    except (RuntimeError, NotImplementedError, AttributeError):   # e.g. gloo builds without the tensor variant
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t, group=group)
        out = torch.cat(parts, dim=0)
    return out


class _GatheredSupCon(torch.autograd.Function):
    """Two collectives in the forward (embeddings + labels in one buffer; row statistics + the rank's loss part in
    another), none in the backward."""

    @staticmethod
    def forward(ctx, x, labels, pair_local, cfg, backend, group):
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        n, c = x.shape
        temperature, loss_weight, min_samples, normalized_input = cfg
        x = x.contiguous()
        fhat = backend.normalize(x, world * n, normalized_input)
        # one gather for [fhat | labels]: the int64 labels travel as two float32-sized columns of raw bits
        lab = labels.contiguous().view(-1).to(torch.int64).view(n, 1).view(fhat.dtype)   # 2 columns of f32, 1 of f64
        packed = torch.empty(n, c + lab.shape[1], dtype=fhat.dtype, device=x.device)
        packed[:, :c] = fhat
        packed[:, c:] = lab
        g_all = _all_gather_rows(packed, group)
        f_all = g_all[:, :c].contiguous()
        labels_all = g_all[:, c:].contiguous().view(torch.int64).view(-1)
        pair_all = _pair_all_on(x.device, pair_local, world)
        loss_part, stats = backend.forward(f_all, labels_all, pair_all, rank * n, n, temperature, loss_weight,
                                           min_samples)
        # one gather for [row statistics ; loss part]: the loss is the sum of the W parts
        tail = torch.zeros(n + 1, stats.shape[1], dtype=stats.dtype, device=x.device)
        tail[:n] = stats
        tail[n, 0] = loss_part
        t_all = _all_gather_rows(tail, group).view(world, n + 1, stats.shape[1])
        stats_all = t_all[:, :n].reshape(world * n, stats.shape[1]).contiguous()
        loss = t_all[:, n, 0].sum()
        ctx.save_for_backward(x, f_all, labels_all, pair_all, stats_all)
        ctx.meta = (rank * n, temperature, normalized_input, world, backend)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        x, f_all, labels_all, pair_all, stats_all = ctx.saved_tensors
        row0, temperature, normalized_input, world, backend = ctx.meta
        g = (grad_out.to(torch.float32) * float(world)).contiguous()   # undo DDP's 1/W gradient averaging
        gx = backend.backward(x, f_all, labels_all, pair_all, stats_all, row0, temperature, normalized_input, g)
        return gx, None, None, None, None, None


def gathered_contrastive_loss(cont_feats, labels, temperature=0.07, loss_weight=1.0, min_samples=10,
                              normalized_input=True, pair_local=None, backend=None, group=None):
    """``ContrastiveLossPlus`` semantics with the contrast set all-gathered over ``group``.

    cont_feats [N, 256] (N equal on every rank), labels [M, 1] or [M] with M <= N (padded with the last label like
    contrastive_loss_plus.py:44-47).  With one rank (or no process group) it equals the local loss."""
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError('gathered_contrastive_loss needs an initialised torch.distributed process group')
    labels = labels.view(-1)
    n = cont_feats.shape[0]
    if labels.shape[0] != n:
        labels = torch.cat([labels, labels[-1:].expand(n - labels.shape[0])])
    if pair_local is None:
        pair_local = reference_pair_map(n)
    backend = backend or CudaBackend()
    return _GatheredSupCon.apply(cont_feats, labels, pair_local,
                                 (temperature, loss_weight, min_samples, normalized_input), backend, group)

"""OA-Loss across ranks: one all-gather of the RoI embeddings before the contrastive loss.

New capability: the reference computes ``loss_cont`` per rank on its own 2048(+rp) rows.  Here every
rank's rows are anchors against the embeddings of ALL ranks:

    rows_r  = [normalize(normalize(x_r)) | y_r]              local kernel, packed rows of 260 floats
    ROWS    = gather(rows_r)                                 exchange 1 over NVLink (2.2 MB per rank at N=2088)
    tail_r  = forward(ROWS; anchors = rows of r)             tcgen05 similarity, rectangular [N, W*N]: row statistics
                                                             (16 bytes per row) + the rank's loss part
    TAIL    = gather(tail_r)                                 exchange 2 (33 KB per rank)
    loss    = sum of the loss parts in rank order            = mean over all W*N anchors, same bits on every rank
    dx_r    = W * dL/dx_r,  dL/dfhat_r = (G + G^T)[rows of r, :] F_all / T

Positives: foreground rows of the same class on ANY rank; background rows only with their own other view (same
rank).  With the per-row statistics of every rank at hand, (G + G^T) restricted to the local rows is exact, so no
reduce-scatter of column gradients is needed.  The factor W compensates DDP's gradient averaging: every rank
reports the same global-mean loss, and sum_r J_r^T (W dL/dx_r) / W is the true gradient.

The two exchanges run either as ``all_gather_into_tensor`` collectives (``exchange='nccl'``; gloo in the CPU tests,
where the local compute is replaced by a numpy stand-in to exercise exactly this plumbing) or -- the default on GPUs
-- with no collective at all (``exchange='peer'``, class PeerExchange): the pack kernel stores its rows straight into
every rank's gather buffer over NVLink and raises a flag there, the forward's tail is scattered the same way, and each
rank's stream waits on its own flags.  A collective kernel needs a free SM on BOTH ranks at the same moment, which the
persistent OA-Mix kernel grants only between its launches; a one-sided store needs nothing on the receiving side.
"""
import ctypes
import os

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


PACK_PAD = 4   # floats appended to a packed row: the int64 label as two 32-bit words + padding (oadg_supcon_pack_width)


class UnpackedBackend:
    """The packed protocol `_GatheredSupCon` speaks, expressed through a backend's plain `normalize / forward /
    backward` (the numpy stand-in of the CPU tests derives from this; the CUDA backend overrides all four with
    direct library calls and no tensor glue)."""

    def pack(self, x, labels, n_total, normalized_input):
        fhat = self.normalize(x, n_total, normalized_input)
        n, c = fhat.shape
        lab = labels.contiguous().view(-1).to(torch.int64)
        if lab.shape[0] != n:
            lab = torch.cat([lab, lab[-1:].expand(n - lab.shape[0])])
        send = torch.zeros(n, c + PACK_PAD, dtype=fhat.dtype, device=fhat.device)
        send[:, :c] = fhat
        w32 = lab.view(n, 1).view(torch.int32).view(n, 2)            # the raw bits travel, never their values
        if fhat.dtype == torch.float32:
            send[:, c:c + 2] = w32.view(torch.float32)
        else:                                                        # float64 stand-in: one 64-bit column
            send[:, c] = lab.view(torch.float64)
        return send

    def forward_packed(self, recv, pair_all, row0, n_rows, temperature, loss_weight, min_samples):
        c = recv.shape[1] - PACK_PAD
        f_all = recv[:, :c].contiguous()
        if recv.dtype == torch.float32:
            labels_all = recv[:, c:c + 2].contiguous().view(torch.int32).view(-1, 2).view(torch.int64).view(-1)
        else:
            labels_all = recv[:, c].contiguous().view(torch.int64)
        self._saved = (f_all, labels_all)
        loss_part, stats = self.forward(f_all, labels_all, pair_all, row0, n_rows, temperature, loss_weight, min_samples)
        tail = torch.zeros(n_rows + 1, stats.shape[1], dtype=stats.dtype, device=stats.device)
        tail[:n_rows] = stats
        tail[n_rows, 0] = loss_part
        return tail

    def finish(self, tail_all, world, n_rows):
        t = tail_all.view(world, n_rows + 1, tail_all.shape[1])
        self._stats_all = t[:, :n_rows].reshape(world * n_rows, tail_all.shape[1]).contiguous()
        return t[:, n_rows, 0].sum()

    def backward_packed(self, x, pair_all, row0, temperature, normalized_input, grad):
        f_all, labels_all = self._saved
        return self.backward(x, f_all, labels_all, pair_all, self._stats_all, row0, temperature, normalized_input, grad)


class CudaBackend(UnpackedBackend):
    """Local compute through libOADG (tcgen05 kernels).  The packed protocol is four library calls on buffers the
    collectives read / write directly: no pack, slice or copy kernels between them."""

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

    # ---- packed protocol
    def pack(self, x, labels, n_total, normalized_input):
        _lib.require_cuda()
        lib = _lib.load()
        n, c = x.shape
        self.ws = self._ws(n_total, c, x.device)
        labels = labels.reshape(-1)
        if labels.dtype != torch.int64 or labels.device != x.device or not labels.is_contiguous():
            labels = labels.to(device=x.device, dtype=torch.int64).contiguous()
        send = torch.empty(n, lib.oadg_supcon_pack_width(c), dtype=torch.float32, device=x.device)
        _lib.check(lib.oadg_supcon_gather_pack(x.data_ptr(), labels.data_ptr(), labels.shape[0], n, n_total, c,
                                               int(normalized_input), send.data_ptr(), self.ws[1], self.ws[2],
                                               _lib.raw_stream(x.device)))
        self.launches += 1
        self._c = c
        return send

    def forward_packed(self, recv, pair_all, row0, n_rows, temperature, loss_weight, min_samples, tail=None):
        lib = _lib.load()
        if tail is None:
            tail = torch.empty(n_rows + 1, 4, dtype=torch.float32, device=recv.device)
        nl = ctypes.c_int(0)
        _lib.check(lib.oadg_supcon_forward_packed(recv.data_ptr(), pair_all.data_ptr(), recv.shape[0], row0, n_rows,
                                                  self._c, float(temperature), float(loss_weight), int(min_samples),
                                                  tail.data_ptr(), self.ws[1], self.ws[2], ctypes.byref(nl),
                                                  _lib.raw_stream(recv.device)))
        self.launches += nl.value
        self._n_total = recv.shape[0]
        return tail

    def finish(self, tail_all, world, n_rows):
        lib = _lib.load()
        loss = torch.empty((), dtype=torch.float32, device=tail_all.device)
        _lib.check(lib.oadg_supcon_finish_packed(tail_all.data_ptr(), world, n_rows, self._c, loss.data_ptr(),
                                                 self.ws[1], self.ws[2], _lib.raw_stream(tail_all.device)))
        self.launches += 1
        return loss

    def backward_packed(self, x, pair_all, row0, temperature, normalized_input, grad):
        lib = _lib.load()
        gx = torch.empty_like(x)
        nl = ctypes.c_int(0)
        _lib.check(lib.oadg_supcon_backward_packed(x.data_ptr(), pair_all.data_ptr(), self._n_total, row0, x.shape[0],
                                                   x.shape[1], float(temperature), int(normalized_input),
                                                   grad.data_ptr(), gx.data_ptr(), self.ws[1], self.ws[2],
                                                   ctypes.byref(nl), _lib.raw_stream(x.device)))
        self.launches += nl.value
        return gx

    # ---- the same protocol with the exchanges done by the kernels themselves (PeerExchange)
    def peer_exchange(self, group, device, n, c):
        """The mapped buffers for steps of up to a capacity of rows (RoI counts vary from step to step with the random
        proposals); a new, larger exchange is set up -- collectively, every rank sees the same n -- only when a step
        exceeds it."""
        key = (id(group), str(device), c)
        px = getattr(self, '_px', None)
        if px is None or getattr(self, '_px_key', None) != key or n > px.n:
            cap = max(4096, (n + 1023) // 1024 * 1024)
            self._px = PeerExchange(group, device, cap, _lib.load().oadg_supcon_pack_width(c))
            self._px_key = key
            if px is not None:               # every rank is past the new exchange's barrier: retire the old buffers
                torch.cuda.synchronize(device)
                dist.barrier(group=group)
                px.close()
        return self._px

    def forward_peers(self, x, labels, pair_all, px, temperature, loss_weight, min_samples, normalized_input):
        lib = _lib.load()
        n, c = x.shape
        n_total = px.world * n
        self.ws = self._ws(n_total, c, x.device)
        labels = labels.reshape(-1)
        if labels.dtype != torch.int64 or labels.device != x.device or not labels.is_contiguous():
            labels = labels.to(device=x.device, dtype=torch.int64).contiguous()
        seq, half = px.begin_step(n)
        s = _lib.raw_stream(x.device)
        _lib.check(lib.oadg_supcon_gather_pack_peers(x.data_ptr(), labels.data_ptr(), labels.shape[0], n, n_total, c,
                                                     int(normalized_input), ctypes.byref(px.peers), px.off_rows[half],
                                                     px.off_flag_rows, px.off_counter, seq, self.ws[1], self.ws[2], s))
        px.wait(px.off_flag_rows, seq, n, s)
        self.launches += 2
        self._c = c
        self.forward_packed(px.rows(half, n), pair_all, px.rank * n, n, temperature, loss_weight, min_samples,
                            tail=px.tail_own(half, n))
        px.scatter_tail(half, seq, n, s)
        px.wait(px.off_flag_tail, seq, n, s)
        self.launches += 2
        return self.finish(px.tail_all(half, n), px.world, n)

    # ---- plain entry points (one rank playing several, tests)
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


PEER_MAX = 16   # OADG_PEER_MAX


class _Peers(ctypes.Structure):
    """oadg_peers_t (include/oadg.h)."""
    _fields_ = [('base', ctypes.c_void_p * PEER_MAX), ('world', ctypes.c_int32), ('rank', ctypes.c_int32)]


class _Raw:
    """A region of the exchange buffer, with the two attributes the backend's library calls read off a tensor."""

    def __init__(self, ptr, shape, device):
        self._ptr, self.shape, self.device = ptr, tuple(shape), device

    def data_ptr(self):
        return self._ptr


class PeerExchange:
    """The step's two exchanges as stores into peer memory (include/oadg.h, `oadg_peer_*`): each rank owns one
    exportable buffer holding, twice (steps alternate between the halves), the gathered packed rows [W * n, width] and
    the gathered tails [W, n + 1, 4], plus one flag word per (exchange, source rank) and a ticket word.  The buffers
    are mapped across the ranks of the node once (CUDA IPC; the 64-byte handles travel through
    ``all_gather_object``); from then on the pack kernel stores a rank's rows into every buffer and raises its flag,
    the forward's tail is scattered the same way, and a rank's kernels wait on its own flags.  No collective kernel
    runs in a step, so nothing of another library has to find a free SM next to the persistent OA-Mix kernel.

    Why two halves are enough: a peer overwrites my rows half of step s at its step s + 2, after it has seen my tail of
    step s + 1, which my stream produced after the forward of step s; my tail half of step s is overwritten by a
    peer's forward of step s + 2, which waited for my rows of s + 2, packed after my `finish` of step s.  The backward
    reads the workspace only."""

    TIMEOUT_MS = 20000

    @staticmethod
    def layout(world, n_rows, width):
        """Byte offsets inside a rank's exchange buffer: two halves of gathered rows, two halves of gathered tails, one
        flag word per (exchange, source rank), the ticket word of the last-block-done protocol; 256-byte sections."""
        a256 = lambda v: (v + 255) // 256 * 256
        rows_bytes = a256(world * n_rows * width * 4)
        tail_bytes = a256(world * (n_rows + 1) * 16)
        off_flag_rows = 2 * rows_bytes + 2 * tail_bytes
        return dict(rows_bytes=rows_bytes, tail_bytes=tail_bytes, off_rows=(0, rows_bytes),
                    off_tail=(2 * rows_bytes, 2 * rows_bytes + tail_bytes), off_flag_rows=off_flag_rows,
                    off_flag_tail=off_flag_rows + 256, off_counter=off_flag_rows + 512, bytes=off_flag_rows + 768)

    def __init__(self, group, device, n_rows, width):
        """``n_rows``: the CAPACITY in rows per rank; a step may exchange any row count up to it (the same on every
        rank: senders tag their flag with it and a waiter that expected another count raises)."""
        lib = _lib.load()
        self.group, self.device = group, device
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > PEER_MAX:
            raise ValueError('PeerExchange: at most %d ranks (one NVLink domain)' % PEER_MAX)
        self.n, self.width = n_rows, width
        for k, v in self.layout(self.world, n_rows, width).items():
            setattr(self, k, v)
        own = ctypes.c_void_p()
        _lib.check(lib.oadg_peer_alloc(self.bytes, ctypes.byref(own)))
        self.own = own.value
        handle = (ctypes.c_ubyte * 64)()
        _lib.check(lib.oadg_peer_export(self.own, handle))
        gathered = [None] * self.world
        dist.all_gather_object(gathered, (bytes(handle), self.bytes, n_rows, width), group=group)
        self.peers = _Peers()
        self.peers.world, self.peers.rank = self.world, self.rank
        for r, (h, nbytes, n_r, w_r) in enumerate(gathered):
            if (nbytes, n_r, w_r) != (self.bytes, n_rows, width):
                raise ValueError('PeerExchange: rank %d has %d rows of %d floats, this rank %d x %d (the gathered loss '
                                 'needs the same number of rows on every rank)' % (r, n_r, w_r, n_rows, width))
            if r == self.rank:
                self.peers.base[r] = self.own
            else:
                p = ctypes.c_void_p()
                _lib.check(lib.oadg_peer_import((ctypes.c_ubyte * 64).from_buffer_copy(h), ctypes.byref(p)))
                self.peers.base[r] = p.value
        fault = ctypes.POINTER(ctypes.c_uint32)()
        _lib.check(lib.oadg_peer_fault_alloc(ctypes.byref(fault)))
        self.fault = fault
        self.seq = 0
        dist.barrier(group=group)            # every buffer is mapped everywhere before the first store

    def begin_step(self, n):
        if self.fault[0] == 2:
            raise _lib.OADGError('gathered OA-Loss: the ranks disagree on the number of rows of a step (every rank '
                                 'must pass the same N)')
        if self.fault[0]:
            raise _lib.OADGError('gathered OA-Loss: a peer rank did not deliver its rows / statistics within %d s'
                                 % (self.TIMEOUT_MS // 1000))
        if n > self.n:
            raise ValueError('PeerExchange: %d rows exceed the capacity of %d' % (n, self.n))
        self.seq += 1
        return self.seq, self.seq & 1

    # a step's regions are packed for ITS row count n (<= capacity): rank r's rows start at row r * n
    def rows(self, half, n):
        return _Raw(self.own + self.off_rows[half], (self.world * n, self.width), self.device)

    def tail_own(self, half, n):
        return _Raw(self.own + self.off_tail[half] + self.rank * (n + 1) * 16, (n + 1, 4), self.device)

    def tail_all(self, half, n):
        return _Raw(self.own + self.off_tail[half], (self.world, n + 1, 4), self.device)

    def wait(self, flag_offset, seq, n, stream):
        _lib.check(_lib.load().oadg_peer_wait(self.own + flag_offset, self.world, seq, n, self.TIMEOUT_MS, self.fault,
                                              stream))

    def scatter_tail(self, half, seq, n, stream):
        off = self.off_tail[half] + self.rank * (n + 1) * 16
        _lib.check(_lib.load().oadg_peer_scatter(ctypes.byref(self.peers), off, (n + 1) * 16, self.off_flag_tail,
                                                 self.off_counter, seq, n, stream))

    def nvlink_bytes(self, n):
        """Bytes this rank stores into its peers' buffers in a step of n rows."""
        return (self.world - 1) * (n * self.width * 4 + (n + 1) * 16)

    def close(self):
        """Unmap the peers' buffers and free this rank's (collective in effect: call it on every rank, after a
        barrier, when no step is in flight)."""
        lib = _lib.load()
        for r in range(self.world):
            if r != self.rank and self.peers.base[r]:
                lib.oadg_peer_release(self.peers.base[r])
                self.peers.base[r] = None
        if self.own:
            lib.oadg_peer_free(self.own)
            self.own = None


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
    """Two exchanges in the forward -- the packed [embeddings | labels] rows, then the packed [row statistics ;
    loss part] tail -- none in the backward; between them only the backend's calls.  ``exchange='nccl'``: two
    ``all_gather_into_tensor`` collectives; ``'peer'``: the kernels store into the peers' buffers (PeerExchange)."""

    @staticmethod
    def forward(ctx, x, labels, pair_local, cfg, backend, group, exchange):
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        n = x.shape[0]
        temperature, loss_weight, min_samples, normalized_input = cfg
        x = x.contiguous()
        pair_all = _pair_all_on(x.device, pair_local, world)
        if exchange == 'peer':
            px = backend.peer_exchange(group, x.device, n, x.shape[1])
            loss = backend.forward_peers(x, labels, pair_all, px, temperature, loss_weight, min_samples,
                                         normalized_input)
            ctx.save_for_backward(x, pair_all)
        else:
            send = backend.pack(x, labels, world * n, normalized_input)
            recv = _all_gather_rows(send, group)
            tail = backend.forward_packed(recv, pair_all, rank * n, n, temperature, loss_weight, min_samples)
            tail_all = _all_gather_rows(tail, group)
            loss = backend.finish(tail_all, world, n)
            ctx.save_for_backward(x, pair_all, recv, tail_all)   # stay alive: the backend reads them again
        # the backward reads the backend's workspace, which the next forward overwrites
        backend.fwd_seq = getattr(backend, 'fwd_seq', 0) + 1
        ctx.meta = (rank * n, temperature, normalized_input, world, backend, backend.fwd_seq)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        x, pair_all = ctx.saved_tensors[:2]
        row0, temperature, normalized_input, world, backend, seq = ctx.meta
        if seq != backend.fwd_seq:
            raise RuntimeError('gathered_contrastive_loss: backward of step %d after the forward of step %d -- the '
                               'loss keeps one step of statistics; call backward before the next forward (or use '
                               'one backend per loss that is alive at the same time)' % (seq, backend.fwd_seq))
        g = (grad_out.to(torch.float32) * float(world)).contiguous()   # undo DDP's 1/W gradient averaging
        gx = backend.backward_packed(x, pair_all, row0, temperature, normalized_input, g)
        return gx, None, None, None, None, None, None


_DEFAULT_BACKEND = None


def _default_backend():
    """One CUDA backend per process: it owns the workspace and the mapped peer buffers."""
    global _DEFAULT_BACKEND
    if _DEFAULT_BACKEND is None:
        _DEFAULT_BACKEND = CudaBackend()
    return _DEFAULT_BACKEND


def gathered_contrastive_loss(cont_feats, labels, temperature=0.07, loss_weight=1.0, min_samples=10,
                              normalized_input=True, pair_local=None, backend=None, group=None, exchange=None):
    """``ContrastiveLossPlus`` semantics with the contrast set all-gathered over ``group``.

    cont_feats [N, 256] (N equal on every rank), labels [M, 1] or [M] with M <= N (padded with the last label like
    contrastive_loss_plus.py:44-47).  In a world of one rank it equals the local loss; without an initialised process
    group it raises.  ``exchange``: 'peer' (default on CUDA, env OADG_EXCHANGE) or 'nccl'; both give the same bits.
    The loss keeps ONE step of statistics per backend: call ``backward`` before the next forward."""
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError('gathered_contrastive_loss needs an initialised torch.distributed process group '
                           '(a world of one rank is fine: the result then equals the local loss)')
    from .contrastive_loss import MIN_TEMPERATURE
    if not temperature >= MIN_TEMPERATURE:
        raise ValueError('gathered_contrastive_loss: temperature %r is below %g (the bound of the fixed-shift exponent '
                         'of the tcgen05 forward)' % (temperature, MIN_TEMPERATURE))
    labels = labels.reshape(-1)
    n = cont_feats.shape[0]
    if labels.shape[0] > n or labels.shape[0] < 1:
        raise ValueError('labels: between 1 and %d entries expected, got %d' % (n, labels.shape[0]))
    if pair_local is None:
        pair_local = reference_pair_map(n)
    backend = backend or _default_backend()
    if exchange is None:
        exchange = os.environ.get('OADG_EXCHANGE', 'peer' if hasattr(backend, 'forward_peers') else 'nccl')
    if exchange not in ('peer', 'nccl'):
        raise ValueError("exchange: 'peer' (stores into the peers' buffers over NVLink) or 'nccl' (two all-gathers)")
    if exchange == 'peer' and (not hasattr(backend, 'forward_peers') or dist.get_world_size(group) == 1):
        exchange = 'nccl'                  # a world of one rank has nobody to store to; stand-in backends have no kernels
    return _GatheredSupCon.apply(cont_feats, labels, pair_local,
                                 (temperature, loss_weight, min_samples, normalized_input), backend, group, exchange)

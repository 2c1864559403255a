"""Stock-torch restatement of the reference OA-Loss, for timing the reference's own GPU path.

ORACLE / TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): imported by tests/ and by bench.py's
`oaloss.torch_reference_ms` leg, never by the product.

The reference computes its loss with eager ATen ops on the training device
(``mmdet/models/losses/oadg/contrastive_loss_plus.py:31-50`` -> ``contrastive_loss.py:170-232`` ``supcontrast`` ->
``:147-167`` ``supcontrast_mask``): two ``F.normalize``, dense N x N float masks (background, same-instance, eye,
same-class, foreground), one ``matmul``, row max, ``exp``, ``log``, masked row sums, a ``.item()`` host sync on the mask
check.  This module issues the same sequence of dense ops on the same shapes (so that its CUDA-event time is the
reference's), written from the closed form in SURVEY.md App. B; it is checked against ``oracle/supcon_np.py`` and the
reference-made goldens in tests/test_oracle_golden.py.
"""
import torch
import torch.nn.functional as F

ORI_SIZE = 1024   # contrastive_loss.py:190: 512 RoIs x 2 images per view, hard-coded


def supcontrast_torch(feats, labels, temper=0.07, min_samples=10, host_sync=True):
    """feats [N, C] float32 (already normalised once, as ContrastiveLossPlus hands them over), labels [N] or [N, 1].
    Returns the 0-dim loss (before loss_weight)."""
    dev = feats.device
    n = feats.shape[0]
    rp = (n % ORI_SIZE) // 2
    y = labels.contiguous().view(-1, 1)
    is_bg = (y == y.max())
    fg, bg = (~is_bg).float(), is_bg.float()
    n_fg = (~is_bg).nonzero(as_tuple=True)[0]                       # :194  (a device->host sync through its size)
    both_bg = bg @ bg.T                                             # :198
    other_view = torch.zeros(n, n, dtype=torch.float32, device=dev)  # :199-207 the hard-wired two-view layout
    eye_o = torch.eye(ORI_SIZE, dtype=torch.float32, device=dev)
    eye_r = torch.eye(rp, dtype=torch.float32, device=dev)
    other_view[:ORI_SIZE, ORI_SIZE:2 * ORI_SIZE] = eye_o
    other_view[ORI_SIZE:2 * ORI_SIZE, :ORI_SIZE] = eye_o
    other_view[2 * ORI_SIZE + rp:2 * ORI_SIZE + 2 * rp, 2 * ORI_SIZE:2 * ORI_SIZE + rp] = eye_r
    other_view[2 * ORI_SIZE:2 * ORI_SIZE + rp, 2 * ORI_SIZE + rp:2 * ORI_SIZE + 2 * rp] = eye_r
    pos_bg = other_view * both_bg                                   # :208
    if n_fg.size(0) <= min_samples:                                 # :210
        return torch.zeros((), dtype=torch.float32, device=dev)
    both_fg = fg @ fg.T                                             # :212
    eye = torch.eye(n, dtype=torch.float32, device=dev)
    same = torch.eq(y, y.T).float()
    pos = ((same - eye) * both_fg + pos_bg).detach()                # :214-219
    if host_sync:                                                   # :221 the reference asserts on a .item()
        assert ((pos != 0) & (pos != 1)).float().sum().item() == 0
    not_self = (torch.ones(n, n, dtype=torch.float32, device=dev) - eye).detach()   # :222-224
    # supcontrast_mask, :147-167
    a = F.normalize(feats, dim=1)
    z = torch.div(torch.matmul(a, a.T), temper)
    z = z - z.max(dim=1, keepdim=True)[0].detach()
    e = torch.exp(z) * not_self
    logp = z - torch.log(e.sum(1, keepdim=True))
    row = (pos * logp).sum(1) / (pos.sum(1) + 1e-8)
    return (-row).mean()


def contrastive_loss_plus_torch(cont_feats, labels, loss_weight=1.0, temperature=0.07, min_samples=10, host_sync=True):
    """ContrastiveLossPlus.forward (contrastive_loss_plus.py:31-50): normalise, pad the labels for the
    random-proposal rows with the last label, weight."""
    if len(cont_feats) == 0:
        return torch.zeros(1)
    f = F.normalize(cont_feats, dim=1)
    if len(f) != len(labels):
        pad = labels[-1, :].repeat(len(f) - len(labels), 1)
        labels = torch.cat([labels, pad], dim=0)
    return loss_weight * supcontrast_torch(f, labels, temper=temperature, min_samples=min_samples, host_sync=host_sync)

import os, sys
sys.path.insert(0, '/root/repo'); os.environ['OADG_DEBUG']='1'
import torch, numpy as np
import bench
from oadg_b200 import ContrastiveLossPlus
from oracle import supcon_np
x, labels = bench.make_roi_set()
xd = x.cuda().requires_grad_(True)
fn = ContrastiveLossPlus(**bench.LOSS_CFG)
loss = fn(xd, labels.cuda())
print('loss', loss.item())
try:
    loss.backward()
    ref, gref = supcon_np.supcon_loss(x.numpy(), labels.numpy(), 0.06, 10, 0.01, want_grad=True)
    g = xd.grad.cpu().numpy()
    print('grad rel err', np.linalg.norm(g-gref)/np.linalg.norm(gref), 'loss rel', abs(loss.item()-ref)/ref)
except Exception as e:
    print('ERR', e)

"""BASELINE config 3: one OA-DG training step of a random-init Faster R-CNN R50-FPN (8 classes) on a synthetic
1024x2048 batch of 2 frames, 1 x B200 (GPU box).

    python scripts/step_bench.py [--steps 10] [--backbone resnet50]                                   # config 3
    python scripts/step_bench.py --arch dc5 --backbone resnet101 --classes 7 --hw 600 1067            # config 5
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/step_bench.py  # config 4: DDP,
                                                        # the contrastive loss over the RoI embeddings of all ranks

The detector (backbone, FPN, RPN, RoIAlign) is stock torch / torchvision, as BASELINE.json's north_star prescribes; the
OA-DG parts are this repo's: OA-Mix produces view 2 on the GPU and its mix kernel writes the Normalize + Pad + CHW
float32 tensors of both views (f3), `integrate_data` stacks them, RoIs are sampled on view 1 and replicated, random
proposals are drawn on the device, and the head's three losses (CrossEntropyLossPlus with the JSD kernel,
SmoothL1LossPlus, ContrastiveLossPlus on [2048 + rp, 256]) run through the registry classes.  Prints one JSON line with
CUDA-event times per phase.  There is no reference number to put beside it: the reference needs mmcv-full, which this
image does not have."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from oadg_b200 import OAMix  # noqa: E402
from oadg_b200.two_view import TwoViewFasterRCNN  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--backbone', default='resnet50')
    ap.add_argument('--classes', type=int, default=8)
    ap.add_argument('--arch', default='fpn', choices=['fpn', 'dc5'])
    ap.add_argument('--hw', type=int, nargs=2, default=[bench.H, bench.W], help='frame height and width')
    args = ap.parse_args()
    import torch.distributed as dist
    world, rank = int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    torch.manual_seed(0)
    norm = dict(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375], to_rgb=True, size_divisor=32)
    mix = OAMix(fused_output=norm, **bench.OAMIX_CFG)
    frames = [bench.make_image(8 * rank + s, args.hw[0], args.hw[1]) for s in range(8)]
    imgs = [torch.from_numpy(f).to(dev) for f, _ in frames]
    gts = [g for _, g in frames]
    rng = np.random.RandomState(0)
    labels = [torch.from_numpy(rng.randint(0, args.classes, len(g))).to(dev) for g in gts]
    gts_dev = [torch.from_numpy(g).to(dev) for g in gts]
    model = TwoViewFasterRCNN(num_classes=args.classes, backbone=args.backbone, arch=args.arch,
                              random_proposal_cfg=dict(num_bboxes=10, scales=(0.01, 0.3), ratios=(0.3, 1 / 0.3),
                                                       iou_max=0.7, iou_min=0.0),
                              loss_cont=dict(loss_weight=0.01, num_views=2, temperature=0.06),
                              gather=world > 1).to(dev).train()
    net = model
    if world > 1:   # every rank starts from rank 0's weights; gradients are averaged by DDP
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local])
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    gen = torch.Generator(device=dev).manual_seed(rank)
    np.random.seed(1000 + rank)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    phases = {'oamix': 0.0, 'forward': 0.0, 'backward': 0.0, 'optimizer': 0.0}
    losses_seen, launches = {}, 0

    def step(i, timed):
        nonlocal launches
        j = (2 * i) % len(imgs)
        batch, gt = [imgs[j], imgs[j + 1]], [gts[j], gts[j + 1]]
        e = [ev() for _ in range(5)]
        e[0].record()
        views, oamix_boxes, ml_boxes = mix.oamix_batch(batch, gt)
        view_f32, src_f32 = mix.last_fused
        launches += mix.last_launches
        data = dict(img=torch.stack(src_f32), img2=torch.stack(view_f32),
                    gt_bboxes=[gts_dev[j], gts_dev[j + 1]], gt_labels=[labels[j], labels[j + 1]],
                    multilevel_boxes=[torch.as_tensor(np.asarray(b, dtype=np.float32), device=dev) for b in ml_boxes],
                    oamix_boxes=[torch.as_tensor(np.asarray(b, dtype=np.float32), device=dev) for b in oamix_boxes])
        e[1].record()
        out = net(data, generator=gen)
        total = sum(out.values())
        e[2].record()
        opt.zero_grad(set_to_none=True)
        total.backward()
        e[3].record()
        opt.step()
        e[4].record()
        torch.cuda.synchronize()
        if timed:
            for k, a, b in (('oamix', 0, 1), ('forward', 1, 2), ('backward', 2, 3), ('optimizer', 3, 4)):
                phases[k] += e[a].elapsed_time(e[b])
            for k, v in out.items():
                losses_seen[k] = float(v)
        return float(total)

    for i in range(args.warmup):
        step(i, False)
    cont = model.roi_head.bbox_head.loss_cont
    cont.stats['launches'] = 0
    launches = 0

    def gathered_launches():
        from oadg_b200 import distributed as D
        return D._DEFAULT_BACKEND.launches - base_launches if D._DEFAULT_BACKEND is not None else 0
    from oadg_b200 import distributed as _D
    base_launches = _D._DEFAULT_BACKEND.launches if _D._DEFAULT_BACKEND is not None else 0
    t0, t1 = ev(), ev()
    torch.cuda.synchronize()
    t0.record()
    last = None
    for i in range(args.steps):
        last = step(args.warmup + i, True)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / args.steps
    if world > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
        if rank != 0:
            dist.barrier()
            dist.destroy_process_group()
            return
    print(json.dumps({
        'n_gpus': world, 'loss_cont': 'gathered over all ranks (peer stores)' if world > 1 else 'local',
        'workload': 'OA-DG two-view training step, torchvision %s-%s Faster R-CNN (random init, %d classes), 2 frames '
                    '%dx%d -> 4 padded float32 images per step, 512 RoIs/img + random proposals, SGD' %
                    (args.backbone, args.arch.upper(), args.classes, args.hw[0], args.hw[1]),
        'steps': args.steps, 'ms_per_step': ms, 'images_per_s': 2 * world * 1e3 / ms,
        'phase_ms_per_step': {k: v / args.steps for k, v in phases.items()},
        'oamix_share_of_step': phases['oamix'] / args.steps / ms,
        'oamix_launches_per_step': launches / args.steps,
        'loss_cont_launches_per_step': (cont.stats.get('launches', 0) + gathered_launches()) / args.steps,
        'rois_per_step': int(model.roi_head.last_rois.shape[0]),
        'losses_last_step': losses_seen, 'total_last_step': last,
        'peak_mem_gb': torch.cuda.max_memory_allocated() / 2 ** 30}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

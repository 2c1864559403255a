#!/usr/bin/env python
"""bench.py -- OA-Mix + OA-Loss images/sec @1024x2048, bs=2/GPU (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One *step* = the per-GPU hot path of one training iteration of
configs/OA-DG/cityscapes/faster_rcnn_r50_fpn_1x_cityscapes_oadg.py on synthetic data:
  OA-Mix  : 2 source frames 1024x2048x3 u8 (8 gt boxes each) -> 2 generated views
  OA-Loss : ContrastiveLossPlus forward + backward on [2088, 256] two-view RoI embeddings
images/sec = source images consumed per second (2 per step per GPU), whole job.

value  : inputs resident in HBM, CUDA-event timed: OAMix.iter_batches over CUDA frames (device views out; includes
         the saliency read-back, host plan sampling and the plan upload -- they are part of the path) with the loss
         forward + backward of each step enqueued as its views arrive.  The loader loop keeps batch k + 1's chain
         in flight while step k's loss runs.
e2e    : the registered plugins called with HOST buffers (pinned) in a loader loop: OAMix.iter_batches over the
         steps' sample dicts (numpy frames in, numpy views out; the pipelined form of OAMix.call_batch, same values)
         and ContrastiveLossPlus on a host tensor with loss.item() every step: H2D of frames / embeddings and D2H of
         the generated views / loss value inside the timed region.
roofline: the OA-Mix chain kernel (one persistent launch per batch), achieved = algorithmic bytes (2 * 3HW per lane
         step) / CUDA-event kernel time from a second, event-instrumented pass over the same seeded plans.
cpu_baseline / --impl reference: the oracle port of the reference's CPU path (oracle/oamix_np.py +
         oracle/supcon_np.py: same cv2 / Pillow / NumPy calls as the reference; the reference itself is
         Python under /root/reference and does not exist on the GPU box) on the host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

H, W, N_GT, BS = 1024, 2048, 8, 2
N_ROI, C_ROI = 2088, 256
OAMIX_CFG = dict(version='augmix', num_views=2, keep_orig=True, severity=10,
                 random_box_ratio=(3, 1 / 3), random_box_scale=(0.01, 0.1),
                 oa_random_box_scale=(0.005, 0.1), oa_random_box_ratio=(3, 1 / 3),
                 spatial_ratio=4, sigma_ratio=0.3)
LOSS_CFG = dict(loss_weight=0.01, num_views=2, temperature=0.06)
POOL = 24  # distinct source frames cycled through: 24 x 6.3 MB = 151 MB > 126 MB L2


def make_image(seed, h=H, w=W, n_gt=N_GT):
    """SURVEY.md 8d generator (same as oracle/synth.py; duplicated so the product arm never imports oracle)."""
    import cv2
    rng = np.random.RandomState(seed)
    base = rng.randint(0, 256, (max(h // 32, 2), max(w // 32, 2), 3)).astype(np.uint8)
    img = cv2.resize(base, (w, h), interpolation=cv2.INTER_CUBIC)
    img = np.clip(img.astype(np.int16) + rng.randint(-12, 13, (h, w, 3)), 0, 255).astype(np.uint8)
    bw = rng.randint(max(w * 32 // 2048, 2), max(w * 400 // 2048, 4), n_gt)
    bh = rng.randint(max(h * 32 // 1024, 2), max(h * 300 // 1024, 4), n_gt)
    x1 = rng.randint(0, w - bw)
    y1 = rng.randint(0, h - bh)
    return img, np.stack([x1, y1, x1 + bw, y1 + bh], axis=1).astype(np.float32)


def make_roi_set(n=N_ROI, c=C_ROI, seed=0, n_fg=200, n_cls=8):
    import torch
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, c, generator=g, dtype=torch.float32)
    base = torch.full((1024,), n_cls, dtype=torch.int64)
    idx = torch.randperm(1024, generator=g)[:n_fg]
    base[idx] = torch.randint(0, n_cls, (n_fg,), generator=g)
    return x, torch.cat([base, base]).view(-1, 1)


def workload_config(n_gpus):
    return {'workload': 'oamix(2x1024x2048x3 u8, 8 gt/img, version=augmix) + '
                        'contrastive_loss_plus fwd+bwd([2088,256] f32, T=0.06) per GPU step',
            'imgs_per_gpu': BS, 'frame': [H, W, 3], 'gt_per_img': N_GT, 'rois': [N_ROI, C_ROI],
            'sharding': ('by image, %d rank(s); OA-Mix: no collective; OA-Loss: the RoI embeddings of all ranks are '
                         'gathered before the contrastive loss (anchors local, contrasts global)'
                         % n_gpus) if n_gpus > 1 else 'single rank',
            'l2': 'inputs larger than L2: %d distinct source frames (%.0f MB) cycled' % (POOL, POOL * H * W * 3 / 1e6),
            'pipeline': 'OAMix.iter_batches: steps travel in groups (1, 2, then 4 steps per plan and chain launch); '
                        'saliency two groups ahead, kernel chain one group ahead of the step that consumes it'}


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons DURING the timed region: an in-process NVML poller (every 1 ms, so that even a
    10 ms region gets several samples), or `nvidia-smi -lms 50` when NVML is not importable."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc, self.thread = index, None, None
        self.sm, self.reasons, self.max_mhz, self.stop_flag = [], set(), None, False

    def _nvml_loop(self, nv, h):
        bits = {'hw_slowdown': nv.nvmlClocksThrottleReasonHwSlowdown,
                'hw_thermal_slowdown': nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                'sw_thermal_slowdown': nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                'sw_power_cap': nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:   # noqa: BLE001  (a failed poll is just a missing sample)
                pass
            time.sleep(0.001)

    def start(self):
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            idx = self.index
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            if vis:   # NVML numbers physical devices
                try:
                    idx = int(vis.split(',')[self.index])
                except (ValueError, IndexError):
                    idx = self.index
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:   # noqa: BLE001
            self.thread = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            return {'sm_mhz': statistics.median(self.sm) if self.sm else None, 'sm_max_mhz': self.max_mhz,
                    'samples': len(self.sm), 'reasons': sorted(self.reasons), 'source': 'nvml, 1 ms poll'}
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons), 'source': 'nvidia-smi -lms 50'}


# ------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference path on the host cores
# ------------------------------------------------------------------------------------------
def _cpu_oamix_worker(job):
    import cv2
    cv2.setNumThreads(1)
    from oracle import oamix_np
    seed, plan_seed = job
    img, gt = make_image(seed)
    np.random.seed(plan_seed)
    t0 = time.perf_counter()
    oamix_np.oamix_view(img, gt, **{k: v for k, v in OAMIX_CFG.items() if k not in ('num_views', 'keep_orig', 'severity')})
    return time.perf_counter() - t0


def cpu_reference_round(pool, cores, round_idx, loss_inputs):
    """One bounded sample: `cores` frames through the CPU OA-Mix (one per worker) + one CPU OA-Loss
    forward/backward.  Returns (images, oamix_wall_s, loss_s)."""
    from oracle import supcon_np
    jobs = [((round_idx * cores + i) % 4, 1000 + round_idx * cores + i) for i in range(cores)]
    t0 = time.perf_counter()
    pool.map(_cpu_oamix_worker, jobs)
    t_mix = time.perf_counter() - t0
    x, labels = loss_inputs
    t0 = time.perf_counter()
    supcon_np.supcon_loss(x, labels, LOSS_CFG['temperature'], 10, LOSS_CFG['loss_weight'], dtype=np.float32, want_grad=True)
    t_loss = time.perf_counter() - t0
    return cores, t_mix, t_loss


def cpu_images_per_sec(n_img, t_mix, t_loss_per_step):
    """A step consumes BS frames and one loss evaluation: time per frame = 1/rate_mix + t_loss/BS."""
    per_img = t_mix / n_img + t_loss_per_step / BS
    return 1.0 / per_img


def run_cpu_baseline(rounds, cores):
    """The CPU arm runs in a fresh interpreter (never forked from the CUDA process): (n_img, t_mix, t_loss)."""
    out = subprocess.run([sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--steps', str(rounds),
                          '--warmup', '0', '--cores', str(cores)], capture_output=True, text=True, timeout=600)
    for line in out.stdout.splitlines():
        if line.startswith('{'):
            d = json.loads(line)['cpu_baseline']
            n_img = rounds * d['cores']
            return n_img, d['oamix_s_per_img_per_core'] * n_img / d['cores'], d['loss_s']
    raise RuntimeError('cpu baseline subprocess failed: ' + out.stderr[-2000:])


def reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = args.cores if args.cores > 0 else min(os.cpu_count() or 1, 64)
    import multiprocessing as mp
    x, labels = make_roi_set()
    loss_inputs = (x.numpy(), labels.numpy())
    ctx = mp.get_context('spawn')
    with ctx.Pool(cores) as pool:
        for r in range(args.warmup):
            cpu_reference_round(pool, min(cores, 4), r, loss_inputs)
        n_img, t_mix, t_loss = 0, 0.0, []
        t_all = time.perf_counter()
        for r in range(args.steps):
            n, tm, tl = cpu_reference_round(pool, cores, r, loss_inputs)
            n_img += n
            t_mix += tm
            t_loss.append(tl)
        t_all = time.perf_counter() - t_all
    value = cpu_images_per_sec(n_img, t_mix, statistics.median(t_loss))
    sample = ('%d rounds x %d frames (1 per worker process, cv2 1 thread each) through the CPU OA-Mix port + '
              '1 CPU OA-Loss fwd+bwd (numpy f32) per round; images/s = 1/(t_mix/frames + t_loss/%d)' %
              (args.steps, cores, BS))
    line = {'impl': 'reference', 'metric': 'oamix+oaloss images/sec @1024x2048 bs=2/GPU', 'value': value,
            'unit': 'images/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': 1e3 * BS / value, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'u8/f32', 'data': 'synthetic', 'config': workload_config(args.gpus),
            'cpu_baseline': {'value': value, 'unit': 'images/s', 'cores': cores, 'kind': 'port', 'sample': sample,
                             'oamix_s_per_img_per_core': t_mix * cores / n_img, 'loss_s': statistics.median(t_loss),
                             'wall_s': t_all},
            'e2e': {'value': value, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), file=_REAL_STDOUT, flush=True)


# ------------------------------------------------------------------------------------------
# product arm
# ------------------------------------------------------------------------------------------
def product_arm(args):
    import torch
    import torch.distributed as dist
    from oadg_b200 import OAMix, ContrastiveLossPlus, build
    build.build()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- inputs (synthetic, resident in HBM) and their pinned host twins for the e2e leg
    frames = [make_image(s) for s in range(rank * POOL, rank * POOL + POOL)]
    gts = [g for _, g in frames]
    host_frames = [torch.from_numpy(f).pin_memory() for f, _ in frames]
    dev_frames = [t.to(dev) for t in host_frames]
    x, labels = make_roi_set(seed=rank)
    x_host = x.pin_memory()
    x_dev = x.to(dev).requires_grad_(True)
    labels_dev = labels.to(dev)
    mix = OAMix(**OAMIX_CFG)
    loss_fn = ContrastiveLossPlus(**LOSS_CFG)
    gather = world > 1 and not args.no_gather
    if gather:
        from oadg_b200.distributed import gathered_contrastive_loss, CudaBackend
        gbe = CudaBackend()
        exchange = args.exchange

    def run_loss(xin):
        if gather:   # one all-gather of the RoI embeddings over NVLink before the contrastive loss
            return gathered_contrastive_loss(xin, labels_dev, temperature=LOSS_CFG['temperature'],
                                             loss_weight=LOSS_CFG['loss_weight'], backend=gbe, exchange=exchange)
        return loss_fn(xin, labels_dev)
    stream = torch.cuda.current_stream(dev)

    def dev_batches(n):
        for i in range(n):
            j = (i * BS) % POOL
            yield [dict(img=dev_frames[(j + b) % POOL], gt_bboxes=gts[(j + b) % POOL]) for b in range(BS)]

    def run_steps(n):
        # the step loop with frames resident in HBM: the registered transform's loader loop (device views out), then
        # the loss forward + backward of the step
        loss = None
        for _ in mix.iter_batches(dev_batches(n)):
            x_dev.grad = None
            loss = run_loss(x_dev)
            loss.backward()
        return loss

    def host_batches(n):
        for i in range(n):
            j = (i * BS) % POOL
            yield [dict(img=host_frames[(j + b) % POOL].numpy(), gt_bboxes=gts[(j + b) % POOL]) for b in range(BS)]

    def e2e_steps(n):
        # the loader loop a user writes: host numpy in, host numpy out, every step's loss read back
        marks = [time.perf_counter()]
        mix.pipe_profile = {}
        for results in mix.iter_batches(host_batches(n)):
            views = [res['img2'] for res in results]
            xd = x_host.to(dev, non_blocking=True).requires_grad_(True)
            loss = run_loss(xd)
            loss.backward()
            last = float(loss.item()), views
            marks.append(time.perf_counter())
        dt = np.diff(np.array(marks)) * 1e3
        log('e2e step wall ms: median %.3f  p90 %.3f  max %.3f  (first %.3f)' % (
            np.median(dt), np.percentile(dt, 90), dt.max(), dt[0]))
        log('e2e pipeline host phases, avg / max us: ' + ', '.join(
            '%s %.0f/%.0f' % (k, v / n * 1e6, mix.pipe_profile[k + '.max'] * 1e6)
            for k, v in mix.pipe_profile.items() if not k.endswith('.max')))
        mix.pipe_profile = None
        return last

    def log(msg):
        if rank == 0:
            print('[bench] ' + msg, file=sys.stderr, flush=True)

    log('inputs ready')
    # ---- warm-up
    np.random.seed(7 + rank)
    run_steps(max(args.warmup, 3))
    barrier()

    log('warm-up done')
    # ---- timed region (device-resident inputs)
    clocks = ClockSampler(local) if rank == 0 else None   # rank 0 reports the line; one sampler, not one per rank
    if clocks is not None:
        clocks.start()
    np.random.seed(1000 + rank)
    loss_fn.stats['launches'] = 0
    if gather:
        gbe.launches = 0
    launches = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    run_steps(args.steps)
    e1.record(stream)
    barrier()
    launches += mix.pipe_launches
    ms = e0.elapsed_time(e1)
    clk = clocks.stop() if clocks is not None else None
    launches += loss_fn.stats['launches'] + (gbe.launches if gather else 0)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = BS * args.steps * world / (ms_max / 1e3)

    log('timed region done: %.3f ms/step' % (ms_max / args.steps))
    # ---- e2e through the registered plugins with host buffers
    e2e_value = None
    if not args.no_e2e:     # (--no-e2e: diagnostics runs only; the driver's command line never passes it)
        np.random.seed(7 + rank)
        e2e_steps(3)
        np.random.seed(1000 + rank)
        barrier()
        e0.record(stream)
        e2e_steps(args.steps)
        e1.record(stream)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_value = BS * args.steps * world / (float(t.item()) / 1e3)
    frame_bytes = H * W * 3
    h2d = BS * frame_bytes + N_ROI * C_ROI * 4
    d2h = BS * frame_bytes + 4

    log('e2e done')
    # ---- N > 1, outside every timed region: all ranks report the same loss bits, and rank 0 checks them against the
    # oracle's closed form on the concatenated batch (oracle/supcon_np.py, float64)
    parity = None
    if gather:
        x_dev.grad = None
        lg = run_loss(x_dev)
        lg.backward()
        mine = torch.stack([lg.detach().double(), x_dev.grad.double().norm()])
        every = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        xs = [torch.empty_like(x_dev) for _ in range(world)]
        dist.all_gather(xs, x_dev.detach())
        ys = [torch.empty_like(labels_dev) for _ in range(world)]
        dist.all_gather(ys, labels_dev)
        if rank == 0:
            from oadg_b200 import reference_pair_map
            from oadg_b200.distributed import gathered_pair_map
            from oracle import supcon_np
            vals = [float(e[0]) for e in every]
            assert all(v == vals[0] for v in vals), 'ranks disagree on the gathered loss: %r' % (vals,)
            y_all = np.concatenate([supcon_np.pad_labels(y.cpu().numpy(), N_ROI) for y in ys])
            x_all = np.concatenate([t.cpu().numpy() for t in xs])
            ref = supcon_np.supcon_loss(x_all, y_all, LOSS_CFG['temperature'], 10, LOSS_CFG['loss_weight'],
                                        pair=gathered_pair_map(reference_pair_map(N_ROI), world))
            rel = abs(vals[0] - ref) / abs(ref)
            assert rel <= 1e-5, 'gathered loss %.9g vs oracle %.9g' % (vals[0], ref)
            parity = {'ranks_equal_bits': True, 'loss_rel_vs_oracle_on_concatenated_batch': rel,
                      'oracle': 'oracle/supcon_np.py float64, %d rows' % (world * N_ROI)}
            log('gathered-loss parity: rel %.2e vs the oracle on %d rows' % (rel, world * N_ROI))
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel: event-instrumented replay of the same seeded plans (rank 0)
    prof = {}
    np.random.seed(1000 + rank)
    gb = max(1, int(mix.group_batches))
    for i0 in range(0, args.steps, gb):   # the loader loop's launches: the views of `gb` consecutive steps per plan
        idx = [((i * BS) % POOL + b) % POOL for i in range(i0, min(i0 + gb, args.steps)) for b in range(BS)]
        mix.oamix_batch([dev_frames[j] for j in idx], [gts[j] for j in idx], profile=prof, inputs_ready=True)
    torch.cuda.synchronize()
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    else:
        peak, peak_src = 6650.0, 'fallback (B200_PROFILING.md)'
    chain_ms, mix_ms, n_l = prof.get('chain_ms', 0.0), prof.get('mix_ms', 0.0), max(prof.get('chain_n', 1), 1)
    achieved = prof.get('step_bytes', 0) / (chain_ms / 1e3) / 1e9 if chain_ms > 0 else 0.0
    total_kernel_ms = chain_ms + mix_ms
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, 'profiles', 'chain_kernel_traffic.json')
    if os.path.exists(tpath):   # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture (profiles/)
        tj = json.load(open(tpath))
        traffic, traffic_src = tj['dram_read_bytes'] + tj['dram_write_bytes'], tj.get('source')
    roofline = {'kernel': 'oadg::oamix_chain_kernel', 'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                'frac': achieved / peak, 'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': peak_src,
                'launches': prof.get('chain_n', 0),
                'algorithmic_bytes_per_launch': prof.get('step_bytes', 0) / n_l,
                'algorithmic_bytes': 'sum over the launch\'s lane steps of 2 * 3HW (one read + one write of a frame per '
                                     'depth step, SURVEY 8d); the launch also does the masks, histograms, LUTs and '
                                     'bboxes-only chains, which count as zero useful bytes',
                'avg_launch_ms': chain_ms / n_l,
                'share_of_oamix_kernel_time': chain_ms / total_kernel_ms if total_kernel_ms else None,
                'work_items_per_launch': prof.get('items', 0) / n_l, 'tiles_per_launch': prof.get('tiles', 0) / n_l,
                'chain_cta_busy_us_per_launch_by_kind': {k: round(v[0] / n_l, 1) for k, v in
                                                         prof.get('kind_busy_us_and_tiles', {}).items()},
                'mix_kernel_ms': mix_ms, 'mix_launches': prof.get('mix_n', 0),
                'oamix_whole_view_gbs': prof.get('view_bytes', 0) / (total_kernel_ms / 1e3) / 1e9 if total_kernel_ms else None,
                'oamix_whole_view_frac': (prof.get('view_bytes', 0) / (total_kernel_ms / 1e3) / 1e9 / peak)
                if total_kernel_ms else None}

    # the timed region runs these launches through the loader loop: the same seeded plans again, OA-Mix only, CUDA events around the whole loop (saliency + chain + mix kernels)
    np.random.seed(1000 + rank)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in mix.iter_batches(dev_batches(args.steps)):
        pass
    e1.record(stream)
    torch.cuda.synchronize()
    loop_ms = e0.elapsed_time(e1)
    roofline['as_run_in_timed_region'] = {
        'mode': 'OAMix.iter_batches; the views of up to %d consecutive steps per plan / chain launch' % gb,
        'oamix_ms_per_batch': loop_ms / args.steps,
        'achieved': prof.get('step_bytes', 0) / (loop_ms / 1e3) / 1e9, 'unit': 'GB/s',
        'frac': prof.get('step_bytes', 0) / (loop_ms / 1e3) / 1e9 / peak,
        'note': 'algorithmic bytes of the chain launches / wall time of the whole OA-Mix loop (its saliency and mix '
                'kernels included); `achieved` above is the chain kernel alone, one full-width launch at a time'}
    log('profiled replay done')
    # OA-Loss alone (CUDA events), for the record
    for _ in range(3):
        x_dev.grad = None
        loss_fn(x_dev, labels_dev).backward()
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(20):
        x_dev.grad = None
        loss_fn(x_dev, labels_dev).backward()
    e1.record(stream)
    torch.cuda.synchronize()
    loss_ms = e0.elapsed_time(e1) / 20
    oaloss = {'fwd_bwd_ms': loss_ms, 'algorithmic_gflop': 6.0 * N_ROI * N_ROI * C_ROI / 1e9,
              'tflops': 6.0 * N_ROI * N_ROI * C_ROI / (loss_ms / 1e3) / 1e12}
    # the reference's own GPU path for the loss: its eager ATen op sequence (oracle/supcon_torch.py, the checker --
    # timed beside the product, never part of it), f32, TF32 off, same inputs, same box, CUDA events; parity first
    from oracle import supcon_torch
    tf32_was = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    xt = x_dev.detach().clone().requires_grad_(True)

    def torch_step():
        xt.grad = None
        l_ = supcon_torch.contrastive_loss_plus_torch(xt, labels_dev, LOSS_CFG['loss_weight'], LOSS_CFG['temperature'])
        l_.backward()
        return l_
    x_dev.grad = None
    ours = loss_fn(x_dev, labels_dev)
    ours.backward()
    ref_l = torch_step()
    rel_l = abs(float(ours) - float(ref_l)) / abs(float(ref_l))
    rel_g = float((x_dev.grad - xt.grad).norm() / xt.grad.norm())
    assert rel_l <= 1e-4 and rel_g <= 1e-4, ('OA-Loss differs from the stock-torch path', rel_l, rel_g)
    for _ in range(3):
        torch_step()
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(20):
        torch_step()
    e1.record(stream)
    torch.cuda.synchronize()
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32_was
    oaloss.update(torch_reference_ms=e0.elapsed_time(e1) / 20, torch_reference='oracle/supcon_torch.py: the '
                  'reference\'s eager ATen op sequence (contrastive_loss.py:147-232), f32, allow_tf32=False, CUDA '
                  'events, same inputs', vs_torch_reference=e0.elapsed_time(e1) / 20 / loss_ms,
                  parity_vs_torch={'loss_rel': rel_l, 'grad_rel': rel_g})

    # ---- the single-sample transform call a stock config makes (OAMix.__call__ on one host frame: H2D, kernels, D2H,
    # one synchronisation per image), for reference next to the batched / pipelined paths
    single_ms = None
    if world == 1:
        np.random.seed(3)
        ts = []
        for k in range(6):
            f, g = frames[k % POOL]
            t0 = time.perf_counter()
            mix(dict(img=f, gt_bboxes=g))
            ts.append((time.perf_counter() - t0) * 1e3)
        single_ms = float(np.median(ts[1:]))
    # ---- CPU baseline (rank 0, N=1 only): bounded sample on the host cores
    log('loss timing done')
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = args.cores if args.cores > 0 else min(os.cpu_count() or 1, 64)   # the same rule as --impl reference
        n_img, t_mix, t_loss = run_cpu_baseline(1, cores)
        cpu = {'value': cpu_images_per_sec(n_img, t_mix, t_loss), 'unit': 'images/s', 'cores': cores, 'kind': 'port',
               'sample': '%d frames (1 per worker process, cv2 1 thread each) through oracle/oamix_np.py + 1 OA-Loss '
                         'fwd+bwd (oracle/supcon_np.py, numpy f32); images/s = 1/(t_mix/frames + t_loss/%d)' % (n_img, BS),
               'oamix_s_per_img_per_core': t_mix * cores / n_img, 'loss_s': t_loss}

    line = {'metric': 'oamix+oaloss images/sec @1024x2048 bs=2/GPU', 'value': value, 'unit': 'images/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_max / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8/f32', 'data': 'synthetic',
            'config': workload_config(world), 'clocks': clk,
            'e2e': {'value': e2e_value, 'unit': 'images/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
            'gpu_launches': launches, 'roofline': roofline, 'oaloss': oaloss, 'cpu_baseline': cpu,
            'single_sample_call_ms': single_ms}
    if gather:
        px = getattr(gbe, '_px', None)
        line['config']['exchange'] = (
            'peer: the pack kernel stores the rows into every rank\'s gather buffer over NVLink and raises a flag, the '
            'row statistics are scattered the same way; no collective kernel in the step' if exchange == 'peer' and px
            else 'nccl: two all_gather_into_tensor per step')
        line['comm'] = {'collectives_per_step': 0 if exchange == 'peer' and px else 2,
                        'nvlink_bytes_out_per_rank_per_step': px.nvlink_bytes(N_ROI) if exchange == 'peer' and px else
                        (world - 1) * (N_ROI * (C_ROI + 4) * 4 + (N_ROI + 1) * 16)}
        line['parity'] = parity
    print(json.dumps(line), file=_REAL_STDOUT, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = sys.stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cores', type=int, default=0, help='worker processes of the CPU arm (0 = min(cpu_count, 64))')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-gather', action='store_true', help='N>1: keep the reference\'s per-rank local loss')
    ap.add_argument('--exchange', default=os.environ.get('OADG_EXCHANGE', 'peer'), choices=['peer', 'nccl'],
                    help='N>1: how the RoI rows travel (peer stores over NVLink without a collective | NCCL all-gathers)')
    args = ap.parse_args()
    # stdout carries exactly one JSON line: anything a library writes to fd 1 meanwhile (e.g. NCCL's version banner
    # under NCCL_DEBUG) goes to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    if args.impl == 'reference':
        reference_arm(args)
    else:
        product_arm(args)


if __name__ == '__main__':
    main()

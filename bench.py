"""bench.py - RoIs/sec of the HTD RoI head (fwd+bwd) on B200, with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): HTD R-50-FPN RoI head fwd+bwd, bf16, 2 images/GPU x 512 RoIs
(128 positives/image), synthetic 800x1333 pyramid (P2..P6, 256 ch), random-init weights.  A step =
``HTDRoIHead.forward_train`` (SFA -> stage 0 -> box refinement -> stage 1 with BA + PGraph) with
the synthetic positives-first sampling of SURVEY.md 8(d), the 7 losses, and the backward pass down
to the pyramid gradients and all 47.2M head-parameter gradients.  N > 1: one process per GPU
(torchrun), each GPU its own images (weak scaling), NCCL all-reduce of the head gradients.

value   device-resident inputs (fp32 NCHW pyramid already in HBM), CUDA-event timed, max over ranks
e2e     same step through the plugin API from HOST buffers: pinned-host pyramid/proposals copied
        H2D every step (double-buffered on a copy stream) and the losses read back D2H
roofline  dominant own kernel (roi_align_bwd), algorithmic bytes of SURVEY 8(d) / live event time
cpu_baseline  the CPU oracle port of the reference path on a bounded sample, rank 0, N = 1 only
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'RoIs/sec HTD RoI head fwd+bwd'
UNIT = 'RoIs/s'
IMGS, ROIS, POS = 2, 512, 128
IMG_H, IMG_W = 800, 1333
# BASELINE.json configs[1]; the SAME string in both arms (the driver compares them)
WORKLOAD = ('HTD R-50-FPN RoI head fwd+bwd, 2 images/GPU x 512 RoIs (128 positives/img), '
            '800x1333 pyramid P2-P6 x 256 ch, random init')


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='htd_b200', choices=['htd_b200', 'reference'])
    ap.add_argument('--dtype', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--imgs-per-gpu', type=int, default=IMGS,
                    help='images per GPU (BASELINE configs[1]: 2; configs[2] = C3: 8)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='eager steps instead of CUDA-graph replay')
    ap.add_argument('--overlap', action='store_true',
                    help='N > 1: all-reduce the head gradients on a side stream, started by an '
                         'in-graph event once they are complete, while the replay still runs the '
                         'backward gather and the layout passes')
    ap.add_argument('--nccl-channels', type=int, default=0,
                    help='with --overlap: NCCL_MAX_NCHANNELS (fewer SMs taken from the gather)')
    ap.add_argument('--no-static', action='store_true',
                    help='skip the extra figure for the step with device-side assign + sample')
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
        return float(p['hbm_gbs']), float(p.get('bf16_tflops_sustained', p['bf16_tflops'])), 'measured'
    except Exception:
        return 6650.0, 1400.0, 'fallback'


# --------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path (the reference itself is pure Python on
# mmcv, absent here and on the GPU box - DESIGN.md "Oracle")
# --------------------------------------------------------------------------------------------
CPU_SAMPLE = dict(imgs=IMGS, rois=ROIS, pos=POS)      # the whole workload: a few seconds per step


def cpu_step_factory():
    import torch
    from htd_b200 import synth
    from oracle import restate
    torch.set_num_threads(os.cpu_count() or 1)
    s = CPU_SAMPLE
    head = restate.HTDRoIHead()
    synth.fill_params_(head, 'init', 0)
    x = [t.requires_grad_(True) for t in synth.make_pyramid(s['imgs'], IMG_H, IMG_W)]
    props = synth.make_proposals(s['imgs'], s['rois'], IMG_H, IMG_W)
    gts = synth.make_gt(s['imgs'], props, num_pos=s['pos'])
    shapes = [(IMG_H, IMG_W, 3)] * s['imgs']

    def step():
        for p in head.parameters():
            p.grad = None
        for t in x:
            t.grad = None
        losses = head.forward_train_sampled(x, props, gts, shapes, s['pos'])
        sum(v for k, v in losses.items() if 'loss' in k).backward()
        return float(losses['s1.loss_cls'])
    return step, s['imgs'] * s['rois']


def cpu_sample_desc():
    s = CPU_SAMPLE
    return (f"oracle port (oracle/restate.py + C RoIAlign, fp32, torch CPU) of the same step on "
            f"{s['imgs']} image x {s['rois']} RoIs ({s['pos']} positives), 800x1333 pyramid")


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    step, rois_per_step = cpu_step_factory()
    for _ in range(max(args.warmup, 0)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = rois_per_step * args.steps / dt
    cores = os.cpu_count() or 1
    line = dict(metric=METRIC, value=val, unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * dt / args.steps, higher_is_better=True,
                scaling='weak', vs_baseline=None, dtype='f32', data='synthetic', impl='reference',
                config=dict(workload=WORKLOAD,
                            global_rois_per_step=rois_per_step,
                            execution='CPU oracle port (oracle/restate.py + C RoIAlign), fp32, '
                                      'torch CPU, every step = the whole workload of one GPU'),
                cpu_baseline=dict(value=val, unit=UNIT, cores=cores, kind='port',
                                  sample=cpu_sample_desc()),
                e2e=dict(value=val, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed regions: an NVML polling thread (2 ms period;
    the timed regions last 0.1-0.3 s, `nvidia-smi -lms` delivered a single sample there), with the
    nvidia-smi loop as the fallback when NVML cannot be loaded."""
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
             'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        import threading
        self.sm, self.mx, self.reasons = [], [], set()
        self.proc = self.thread = None
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = int(vis.split(',')[gpu_index]) if vis and vis.split(',')[gpu_index].isdigit() \
                else gpu_index
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            bits = {'hw_slowdown': pynvml.nvmlClocksThrottleReasonHwSlowdown,
                    'hw_thermal_slowdown': pynvml.nvmlClocksThrottleReasonHwThermalSlowdown,
                    'sw_thermal_slowdown': pynvml.nvmlClocksThrottleReasonSwThermalSlowdown,
                    'sw_power_cap': pynvml.nvmlClocksThrottleReasonSwPowerCap}

            def poll():
                while not self._stop.is_set():
                    try:
                        self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        self.mx.append(mx)
                        r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        for nm, b in bits.items():
                            if r & b:
                                self.reasons.add(nm)
                    except Exception:
                        pass
                    time.sleep(0.002)
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            self.how = 'nvml'
        except Exception:
            self.how = 'nvidia-smi'
            self.path = tempfile.mktemp(suffix='.csv')
            try:
                self.proc = subprocess.Popen(
                    ['nvidia-smi', f'--id={gpu_index}', f'--query-gpu={self.QUERY}',
                     '--format=csv,noheader,nounits', '-lms', '10'],
                    stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
            except Exception:
                self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0, how=self.how)
        if self.thread is not None:
            self._stop.set()
            self.thread.join(timeout=2)
        elif self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
            try:
                for ln in open(self.path):
                    f = [t.strip() for t in ln.split(',')]
                    if len(f) < 9:
                        continue
                    try:
                        self.sm.append(float(f[1]))
                        self.mx.append(float(f[2]))
                    except ValueError:
                        continue
                    for nm, v in zip(names, f[5:9]):
                        if v.lower().startswith('active'):
                            self.reasons.add(nm)
                os.unlink(self.path)
            except Exception:
                pass
        if self.sm:
            sm = sorted(self.sm)
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(self.mx), samples=len(sm))
        out['reasons'] = sorted(self.reasons)
        return out



# --------------------------------------------------------------------------------------------
# north_star evidence measured in the same run (N = 1): tensor-pipe roofline of the PGraph
# aggregation, the RoIAlign / BA gather sweep (BASELINE config 5), the same-box GPU comparator
# --------------------------------------------------------------------------------------------
def _median_ms(fn, iters, flush):
    import torch
    ts = []
    for i in range(iters + 3):
        flush.fill_(float(i))                      # > L2: every timed launch starts cold
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def gather_kernel_times(dev, flush):
    """Device time of the three gather launches of a step at the bench size (2 x 512 RoIs, 128
    positives / image, bf16), each captured in a small CUDA graph and replayed with a cold L2 -
    eager event timing of these launches is dominated by host launch gaps (the plan is three tiny
    kernels).  forward figures include the sampling plan (footprints, scan, axis-weight tables)."""
    import torch
    from htd_b200 import ops, synth
    scales = [0.25, 0.125, 0.0625, 0.03125]
    C = 256
    x = [ops.to_channels_last(t.to(dev), torch.bfloat16) for t in synth.make_pyramid(IMGS)[:4]]
    shapes = [tuple(t.shape) for t in x]
    props = synth.make_proposals(IMGS, ROIS)
    rois = torch.cat([torch.cat([p.new_full((p.size(0), 1), i), p], 1) for i, p in enumerate(props)]).to(dev)
    pos = torch.cat([torch.cat([p.new_full((POS, 1), i), p[:POS]], 1) for i, p in enumerate(props)]).to(dev)
    lv = ops.level_assign(rois, 4)
    K, P = rois.shape[0], pos.shape[0]
    out = torch.empty(K, 7, 7, C, device=dev, dtype=torch.bfloat16)
    outb = torch.empty(4, P, 7, 7, C, device=dev, dtype=torch.bfloat16)
    plan = ops.RoIPlan(x, scales, rois, lv, 7, 0)
    planb = ops.RoIPlan(x, scales, pos, None, 7, 0)
    g = torch.randn(K, 7, 7, C, device=dev).to(torch.bfloat16)
    gp = torch.randn(P, 7, 7, C, device=dev).to(torch.bfloat16)
    wts = torch.rand(4, P, device=dev)
    dm = torch.randn(4 * P, C, device=dev)
    srcs = [dict(rois=rois, plan=plan.tensors(), dy=g, dy_per_level=False),
            dict(rois=rois, plan=plan.tensors(), dy=g, dy_per_level=False),
            dict(rois=pos, plan=planb.tensors(), dy=gp, dy_per_level=False, scale=wts, ring_edge=1,
                 addvec=dm)]

    def fwd_single():
        pl = ops.RoIPlan(x, scales, rois, lv, 7, 0)
        ops._fwd_launch('f', x, scales, rois, lv, 7, 0, None, out, plan=pl)

    def fwd_ba():
        pl = ops.RoIPlan(x, scales, pos, None, 7, 0)
        ops._fwd_launch('f', x, scales, pos, None, 7, 0, None, outb, plan=pl)

    def bwd_fused():
        ops._bwd_multi(shapes, torch.bfloat16, False, scales, srcs, 7)

    res = {}
    for name, fn in (('roi_align_fwd(single)', fwd_single), ('roi_align_fwd(BA)', fwd_ba),
                     ('roi_align_bwd(fused)', bwd_fused)):
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            fn()
        res[name] = _median_ms(gr.replay, 15, flush)
    return res


def pgraph_tensor_roofline(dev, flush):
    """The PGraph aggregation contraction A[n,n] x X[n,1024] (htd_bbox_head.py:213,216) of ONE dense
    group of n RoIs (BASELINE config 4 stress) on the tcgen05 kernel: 2 n^2 d flops / CUDA-event
    time of the launch / measured bf16 peak (burst: the kernel is timed alone)."""
    import torch
    from htd_b200 import pgraph
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            pk = json.load(f)
        burst, sust = float(pk['bf16_tflops']), float(pk.get('bf16_tflops_sustained', pk['bf16_tflops']))
    except Exception:
        burst, sust = 1667.0, 1408.7
    out = []
    d = 1024
    for n in (1000, 4096):
        npad = (n + 63) // 64 * 64
        A = torch.zeros(npad, npad, device=dev, dtype=torch.bfloat16)
        A[:n, :n] = torch.randn(n, n, device=dev).to(torch.bfloat16)
        XT = torch.zeros(d, npad, device=dev, dtype=torch.bfloat16)
        XT[:, :n] = torch.randn(d, n, device=dev).to(torch.bfloat16)
        D = torch.empty(npad, d, device=dev, dtype=torch.bfloat16)
        grp = [dict(M=n, N=d, K=n)]
        ms = _median_ms(lambda: pgraph._gemm(A, XT, grp, D=D, ldd=d), 20, flush)
        tf = 2.0 * n * n * d / (ms * 1e-3) / 1e12
        out.append(dict(n=n, flops=2 * n * n * d, ms=ms, achieved=tf, unit='TFLOP/s', peak=burst,
                        frac=tf / burst, frac_of_sustained=tf / sust))
    return dict(bound='tensor', kernel='pgraph_gemm_kernel (tcgen05, A_local.X / A_global.Xm)',
                peak_kind='measured burst (MEASURED_PEAKS.json bf16_tflops)', sizes=out)


def gather_sweep(dev, flush, hbm_peak):
    """BASELINE config 5: RoIAlign / BA gather sweep over the RoI count (2 images, 256 ch, 7x7 bins,
    P2-P5, bf16): algorithmic bytes of SURVEY 8(d) / event time / measured HBM peak per kernel."""
    import torch
    from htd_b200 import ops, synth
    scales = [0.25, 0.125, 0.0625, 0.03125]
    C, PP, bs = 256, 49, 2
    x = [ops.to_channels_last(t.to(dev), torch.bfloat16) for t in synth.make_pyramid(2)[:4]]
    shapes = [tuple(t.shape) for t in x]
    rows = []
    for per in (128, 512, 2048):
        props = synth.make_proposals(2, per)
        rois = torch.cat([torch.cat([p.new_full((p.size(0), 1), i), p], 1)
                          for i, p in enumerate(props)]).to(dev)
        pos = torch.cat([torch.cat([p.new_full((per // 4, 1), i), p[:per // 4]], 1)
                         for i, p in enumerate(props)]).to(dev)
        lv = ops.level_assign(rois, 4)
        K, P = rois.shape[0], pos.shape[0]
        plan = ops.RoIPlan(x, scales, rois, lv, 7, 0)
        px = plan.pixels()
        out = torch.empty(K, 7, 7, C, device=dev, dtype=torch.bfloat16)
        g = torch.randn(K, 7, 7, C, device=dev).to(torch.bfloat16)

        def fwd_single():
            pl = ops.RoIPlan(x, scales, rois, lv, 7, 0)         # plan time is inside the figure
            ops._fwd_launch('f', x, scales, rois, lv, 7, 0, None, out, plan=pl)
        by = px * C * bs + K * PP * C * bs + 20 * K
        ms = _median_ms(fwd_single, 10, flush)
        rows.append(dict(rois=K, kernel='fwd single (plan included)', ms=ms,
                         frac=by / ms / 1e6 / hbm_peak))
        by = K * PP * C * bs + px * C * 4
        ms = _median_ms(lambda: ops._roi_align_bwd(shapes, torch.bfloat16, scales, rois,
                                                   plan.tensors(), 7, g, False), 10, flush)
        rows.append(dict(rois=K, kernel='bwd single', ms=ms, frac=by / ms / 1e6 / hbm_peak))
        planb = ops.RoIPlan(x, scales, pos, None, 7, 0)
        pxb = planb.pixels()
        outb = torch.empty(4, P, 7, 7, C, device=dev, dtype=torch.bfloat16)

        def fwd_ba():
            pl = ops.RoIPlan(x, scales, pos, None, 7, 0)
            ops._fwd_launch('f', x, scales, pos, None, 7, 0, None, outb, plan=pl)
        by = pxb * C * bs + 4 * P * PP * C * bs + 20 * P
        ms = _median_ms(fwd_ba, 10, flush)
        rows.append(dict(rois=P, kernel='fwd BA all levels (plan included)', ms=ms,
                         frac=by / ms / 1e6 / hbm_peak))
        gp = torch.randn(P, 7, 7, C, device=dev).to(torch.bfloat16)
        wts = torch.rand(4, P, device=dev)
        dm = torch.randn(4 * P, C, device=dev)
        by = P * PP * C * bs + pxb * C * 4
        ms = _median_ms(lambda: ops._roi_align_bwd(shapes, torch.bfloat16, scales, pos,
                                                   planb.tensors(), 7, gp, False, scale=wts,
                                                   ring_edge=1, addvec=dm), 10, flush)
        rows.append(dict(rois=P, kernel='bwd BA all levels', ms=ms, frac=by / ms / 1e6 / hbm_peak))
    return dict(what='BASELINE config 5: gather sweep, bf16, 2 images, fraction of the measured HBM '
                     'roofline (algorithmic bytes of SURVEY 8d; L2 flushed before every launch)',
                rows=[dict(r, ms=round(r['ms'], 4), frac=round(r['frac'], 3)) for r in rows])


def gpu_comparator(dev, pyr_host, props_host, gts, shapes):
    """Same-box GPU comparator: the reference's algorithm as PyTorch-eager CUDA with
    torchvision's CUDA roi_align (the generic thread-per-output / atomicAdd kernel mmcv's is derived
    from), fp32 as the reference runs it.  A reported baseline like the CPU arm: it executes the
    oracle's restatement (the one other leg of bench.py that may), never the product."""
    import torch
    from torchvision.ops import roi_align
    from htd_b200 import synth
    from oracle import restate
    ref = restate.HTDRoIHead()
    synth.fill_params_(ref, 'init', 0)
    ref = ref.to(dev)
    orig = restate.RoIAlign.forward
    restate.RoIAlign.forward = lambda self, x, rois: roi_align(
        x, rois.to(x.dtype), self.output_size, self.spatial_scale, self.sampling_ratio, self.aligned)
    try:
        xr = [t.to(dev).requires_grad_(True) for t in pyr_host]
        props = [p.to(dev) for p in props_host]

        def step():
            for p in ref.parameters():
                p.grad = None
            for t in xr:
                t.grad = None
            losses = ref.forward_train_sampled(xr, props, gts, shapes, POS)
            sum(v for k, v in losses.items() if 'loss' in k).backward()
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        reps = 5
        for _ in range(reps):
            step()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
    finally:
        restate.RoIAlign.forward = orig
    return dict(what='reference algorithm as PyTorch-eager CUDA (oracle restatement on the GPU, '
                     'torchvision CUDA roi_align with atomicAdd backward), fp32, same workload',
                ms_per_step=ms, rois_per_s=IMGS * ROIS / (ms * 1e-3))


def run_gpu(args):
    import torch
    import torch.distributed as dist
    if args.nccl_channels > 0:
        os.environ['NCCL_MAX_NCHANNELS'] = str(args.nccl_channels)
    if os.environ.get('NCCL_DEBUG'):        # keep NCCL's log (rank / transport evidence), but on
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')   # stderr: stdout is the JSON line
    import htd_b200
    from htd_b200 import _lib, synth
    from htd_b200.parallel import GradAllReducer

    # stdout carries exactly ONE line (the JSON): libraries that print there (NCCL's version banner
    # at communicator creation, whatever NCCL_DEBUG says) are sent to stderr for the whole run
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit('launch N > 1 with: python -m torch.distributed.run --nnodes=1 '
                             f'--nproc-per-node {args.gpus} --master-addr 127.0.0.1 bench.py ...')
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: htd_b200 has no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    dtype = torch.bfloat16 if args.dtype == 'bf16' else torch.float32
    hbm_peak, tc_peak, peak_kind = peaks()

    # ---- model + synthetic inputs (each rank its own images: weak scaling) -------------------
    head = htd_b200.build_htd_roi_head()
    synth.fill_params_(head, 'init', 0)
    head = head.to(dev).to(dtype)
    head.compute_dtype = dtype
    head.train()
    imgs = max(1, args.imgs_per_gpu)
    img0 = rank * imgs
    pyr_host = [t.pin_memory() for t in synth.make_pyramid(imgs, IMG_H, IMG_W, seed=1000 + img0)]
    props_host = [p.pin_memory() for p in synth.make_proposals(imgs, ROIS, IMG_H, IMG_W,
                                                               seed=1234 + img0)]
    gts = synth.make_gt(imgs, props_host, num_pos=POS, seed=4321 + img0)
    gts = [{k: v.to(dev) for k, v in g.items()} for g in gts]
    shapes = [(IMG_H, IMG_W, 3)] * imgs
    x_dev = [t.to(dev).requires_grad_(True) for t in pyr_host]
    props_dev = [p.to(dev) for p in props_host]
    reducer = GradAllReducer(head.parameters(), world) if world > 1 else None
    rois_per_step = imgs * ROIS

    def eager_step(x, props):
        for p in head.parameters():
            p.grad = None
        for t in x:
            t.grad = None
        losses = synth.sampled_forward_train(head, x, props, gts, shapes, POS)
        total = sum(v for k, v in losses.items() if 'loss' in k)
        total.backward()
        if reducer is not None:
            reducer.allreduce()
        return losses

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (eager), one accounting step, one instrumented pass for per-kernel times -----
    for _ in range(max(args.warmup, 3)):
        eager_step(x_dev, props_dev)
    barrier()
    _lib.ACCOUNT = []
    l0 = _lib.LAUNCHES['total']
    eager_step(x_dev, props_dev)
    torch.cuda.synchronize()
    launches_per_step = _lib.LAUNCHES['total'] - l0
    account = _lib.ACCOUNT
    _lib.ACCOUNT = None
    launches_per_step -= 0          # accounting adds no launches of the library
    pg_flops = head.bbox_head[1].last_plan.flops()
    alg, alg_flops = {}, {}
    for name, amount in account:
        if name.endswith(':flops'):
            alg_flops.setdefault(name[:-6], []).append(amount)
        else:
            alg.setdefault(name, []).append(amount)
    ksteps = min(args.steps, 10)
    _lib.TIMER = _lib.KernelTimer()
    ek0, ek1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ek0.record()
    for _ in range(ksteps):
        eager_step(x_dev, props_dev)
    ek1.record()
    torch.cuda.synchronize()
    ksum = _lib.TIMER.summary()
    _lib.TIMER = None
    ms_eager = ek0.elapsed_time(ek1) / ksteps

    # ---- the step as one CUDA graph (forward + losses + backward), replayed per step ---------
    use_graph = not args.no_graph
    gstep = None
    if use_graph:
        if reducer is not None:
            reducer.remove()
        from htd_b200.graphed import GraphedTrainStep
        try:
            gstep = GraphedTrainStep(head, x_dev, props_dev, gts, shapes, POS, flat_grads=world > 1,
                                     early_modules=[head.bbox_head[0], head.bbox_head[1],
                                                    head.bbox_roi_extractor[1]]
                                     if world > 1 and args.overlap else None, flat_inputs=True)
        except Exception as e:                     # never lose the measurement to a capture problem
            print(f'[bench] CUDA-graph capture failed ({type(e).__name__}: {e}); running eager',
                  file=sys.stderr)
            use_graph = False
            torch.cuda.synchronize()
            if world > 1:
                reducer = GradAllReducer(head.parameters(), world)
    if use_graph:

        comm_stream = torch.cuda.Stream(device=dev) if world > 1 else None

        def allreduce_grads():
            if world == 1:
                return
            if gstep.early_event is None:
                dist.all_reduce(gstep.flat_grad, op=dist.ReduceOp.AVG)   # one NCCL call, in place
                return
            # the stage-1 head / BA gradients (2/3 of the bytes) are complete long before the
            # replay ends: their all-reduce starts on a side stream as soon as the in-graph event
            # fires and overlaps the stage-0 backward and the RoIAlign gather; the rest follows
            cur = torch.cuda.current_stream()
            with torch.cuda.stream(comm_stream):
                comm_stream.wait_event(gstep.early_event)
                dist.all_reduce(gstep.early_grad, op=dist.ReduceOp.AVG)
            dist.all_reduce(gstep.late_grad, op=dist.ReduceOp.AVG)
            cur.wait_stream(comm_stream)

        def step(x=None, props=None):
            losses = gstep(x, props)
            allreduce_grads()
            return losses
    else:
        def step(x=None, props=None):
            return eager_step(x if x is not None else x_dev, props if props is not None else props_dev)

    for _ in range(3):
        step()
    # ---- timed region 1: device-resident inputs ----------------------------------------------
    clocks = ClockSampler(local)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = launches_per_step * args.steps

    # ---- timed region 2: end to end from host buffers (H2D double-buffered, D2H of the losses)
    copy_stream = torch.cuda.Stream(device=dev)
    h2d_bytes = sum(t.numel() * t.element_size() for t in pyr_host) + \
        sum(p.numel() * p.element_size() for p in props_host)

    # one pinned staging buffer (pyramid levels + proposals back to back): ONE H2D copy per step
    parts = list(pyr_host) + list(props_host)
    sizes = [t.numel() for t in parts]
    host_flat = torch.empty(sum(sizes), dtype=torch.float32).pin_memory()
    off = 0
    for t, n in zip(parts, sizes):
        host_flat[off:off + n].copy_(t.reshape(-1))
        off += n
    direct = use_graph and gstep.input_flat is not None

    def upload():
        """Eager fallback: H2D into a fresh tensor on the copy stream (then copied into the step)."""
        with torch.cuda.stream(copy_stream):
            flat = host_flat.to(dev, non_blocking=True)
            views, off = [], 0
            for t, n in zip(parts, sizes):
                views.append(flat[off:off + n].view(t.shape))
                off += n
            evt = torch.cuda.Event()
            evt.record(copy_stream)
        return views[:len(pyr_host)], views[len(pyr_host):], evt, flat

    def e2e_loop(n):
        """Per step: H2D of the step's inputs from pinned host memory, the step itself, and an
        asynchronous D2H of the 7 losses into pinned memory; the host reads step i-1's losses while
        step i runs (every step's result still reaches the host inside the timed region).
        Graph path: the copy stream writes STRAIGHT into the static buffers the graph reads (one
        183 MB H2D, no staging tensor, no device-to-device copy); it may start as soon as the
        running step has consumed its inputs (an event recorded inside the graph ~0.1 ms into
        the step), so the upload of step i+1 overlaps the rest of step i."""
        pinned = [torch.empty(7, dtype=torch.float32).pin_memory() for _ in range(2)]
        done = [torch.cuda.Event() for _ in range(2)]
        seen = 0.0
        cur = torch.cuda.current_stream()
        if direct:
            copy_stream.wait_stream(cur)
            up = torch.cuda.Event()
            for i in range(n):
                with torch.cuda.stream(copy_stream):
                    if i > 0:
                        copy_stream.wait_event(gstep.inputs_consumed)   # step i-1 has read its inputs
                    gstep.input_flat.copy_(host_flat, non_blocking=True)
                    up.record(copy_stream)
                cur.wait_event(up)
                gstep.graph.replay()
                allreduce_grads()
                pinned[i & 1].copy_(gstep.loss_vec, non_blocking=True)
                done[i & 1].record()
                if i > 0:
                    done[(i - 1) & 1].synchronize()
                    seen += float(pinned[(i - 1) & 1][0])
            done[(n - 1) & 1].synchronize()
            seen += float(pinned[(n - 1) & 1][0])
            return pinned[0].numel() * pinned[0].element_size()
        nxt = upload()
        for i in range(n):
            xs, ps, evt, flat = nxt
            if i + 1 < n:
                nxt = upload()                      # next step's H2D overlaps this step's compute
            cur.wait_event(evt)
            flat.record_stream(cur)
            if not use_graph:
                xs = [t.detach().requires_grad_(True) for t in xs]
            losses = step(xs, ps)
            # the captured step leaves the 7 losses stacked in one device vector (part of the graph)
            dev_l = gstep.loss_vec if use_graph else \
                torch.stack([v.detach().float().reshape(()) for v in losses.values()])
            pinned[i & 1].copy_(dev_l, non_blocking=True)
            done[i & 1].record()
            if i > 0:
                done[(i - 1) & 1].synchronize()
                seen += float(pinned[(i - 1) & 1][0])
        done[(n - 1) & 1].synchronize()
        seen += float(pinned[(n - 1) & 1][0])
        return pinned[0].numel() * pinned[0].element_size()

    e2e_loop(2)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    d2h_bytes = e2e_loop(args.steps)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    clk = clocks.stop()                     # sampled over both timed regions

    # ---- the COMPLETE step: assign + sample on the device, then the same two stages, one graph --
    static = None
    if world == 1 and use_graph and not args.no_static:
        try:
            from htd_b200.graphed import GraphedStaticTrainStep
            nprop = 2000                               # configs/htd/htd_resnet50_1x.py:115-121
            sp, sg, sl, sn = (t.to(dev) for t in synth.make_detection_batch(imgs, nprop))
            metas = [dict(img_shape=s_, scale_factor=1.0) for s_ in shapes]
            sstep = GraphedStaticTrainStep(head, x_dev, metas, sp, sg, sl, sn)   # keys drawn in-graph
            for _ in range(3):
                sstep()
            torch.cuda.synchronize()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for _ in range(args.steps):
                sstep()
            s1.record()
            torch.cuda.synchronize()
            sms = s0.elapsed_time(s1) / args.steps
            cnt = [S.counts.tolist() for S in head.last_static]
            static = dict(what='forward_train_static as one CUDA graph: MaxIoUAssigner + RandomSampler '
                               'on the device (htd_assign_sample, random keys drawn in-graph), both '
                               'stages, losses, backward', proposals_per_img=nprop,
                          sampled_rois_per_step=rois_per_step, ms_per_step=sms,
                          rois_per_s=rois_per_step / (sms * 1e-3),
                          last_counts_pos_neg_per_img=[[c[:2] for c in st] for st in cnt])
        except Exception as e:
            print(f'[bench] static step failed ({type(e).__name__}: {e})', file=sys.stderr)

    # ---- the producer (FPN) emitting channels-last maps in the compute dtype -------
    nhwc = None
    if world == 1 and use_graph and not args.no_static and dtype == torch.bfloat16:
        try:
            from htd_b200.graphed import GraphedTrainStep
            x_cl = [t.detach().to(dtype).contiguous(memory_format=torch.channels_last)
                    .requires_grad_(True) for t in x_dev[:4]] + [x_dev[4]]     # P6: SFA input
            cstep = GraphedTrainStep(head, x_cl, props_dev, gts, shapes, POS)
            for _ in range(3):
                cstep()
            torch.cuda.synchronize()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for _ in range(args.steps):
                cstep()
            c1.record()
            torch.cuda.synchronize()
            cms = c0.elapsed_time(c1) / args.steps
            nhwc = dict(what='same step, pyramid handed over channels-last in bf16 (no layout / cast '
                             'pass in either direction; dX returned channels-last bf16)',
                        ms_per_step=cms, rois_per_s=rois_per_step / (cms * 1e-3))
        except Exception as e:
            print(f'[bench] channels-last pyramid step failed ({type(e).__name__}: {e})',
                  file=sys.stderr)

    # ---- inference: 1 image x 1000 proposals, both stages + decode + NMS ----
    infer = None
    if world == 1 and not args.no_static:
        try:
            head.eval()
            ip = [synth.make_proposals(1, 1000, IMG_H, IMG_W, seed=999)[0].to(dev)]
            ix = [t[:1].detach() for t in x_dev]
            imeta = [dict(img_shape=shapes[0], scale_factor=1.0)]
            with torch.no_grad():
                for _ in range(3):
                    head.simple_test(ix, ip, imeta)
                torch.cuda.synchronize()
                i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                nrep = min(args.steps, 20)
                i0.record()
                for _ in range(nrep):
                    head.simple_test(ix, ip, imeta)
                i1.record()
                torch.cuda.synchronize()
            ims = i0.elapsed_time(i1) / nrep
            infer = dict(what='HTDRoIHead.simple_test, 1 image x 1000 proposals: SFA, both stages (BA on '
                              'all 1000 RoIs, PGraph), decode, multi-class NMS (htd_multiclass_nms), '
                              'results to numpy; eager', ms_per_image=ims,
                         rois_per_s=1000 / (ims * 1e-3))
            head.train()
        except Exception as e:
            head.train()
            print(f'[bench] inference figure failed ({type(e).__name__}: {e})', file=sys.stderr)

    # ---- producers of the path: FPN neck + RPN proposals at the same image size ----
    producers = None
    if world == 1 and not args.no_static and dtype == torch.bfloat16:
        try:
            from htd_b200.dense_heads import RPNHead
            from htd_b200.necks import FPN
            fpn = FPN([256, 512, 1024, 2048], 256, 5).to(dev).to(dtype)
            rpn = RPNHead(256, 256).to(dev).to(dtype).to(memory_format=torch.channels_last)
            gb = torch.Generator().manual_seed(5)
            cs = [(torch.randn(imgs, c, t.shape[2], t.shape[3], generator=gb) * 0.5).to(dev)
                  .to(dtype).contiguous(memory_format=torch.channels_last).requires_grad_(True)
                  for c, t in zip((256, 512, 1024, 2048), x_dev)]
            pmeta = [dict(img_shape=shapes[0], scale_factor=1.0)] * imgs
            pcfg = dict(nms_across_levels=False, nms_pre=2000, nms_post=2000, max_num=2000, nms_thr=0.7,
                        min_bbox_size=0)

            def fpn_step():
                outs = fpn(cs)
                torch.autograd.backward(outs, [torch.ones_like(o) for o in outs])
                return outs

            def timed(fn, n=5):
                for _ in range(2):
                    r = fn()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(n):
                    r = fn()
                b.record()
                torch.cuda.synchronize()
                return a.elapsed_time(b) / n, r
            f_ms, outs = timed(fpn_step)
            with torch.no_grad():
                feats = [o.detach() for o in outs]
                r_ms, props = timed(lambda: rpn.get_bboxes(*rpn(feats), pmeta, pcfg))
            producers = dict(what='htd_b200.necks.FPN forward+backward (C2-C5 channels-last bf16 in, five '
                                  'channels-last bf16 maps out: the head reads them without a layout pass) '
                                  'and htd_b200.dense_heads.RPNHead forward + get_bboxes (nms_pre 2000, '
                                  'nms 0.7, nms_post 2000) for the same images; eager, random weights',
                             fpn_fwd_bwd_ms=f_ms, rpn_proposals_ms=r_ms,
                             proposals_per_img=[int(p.shape[0]) for p in props])
            del fpn, rpn, cs, outs, feats
        except Exception as e:
            print(f'[bench] producer figure failed ({type(e).__name__}: {e})', file=sys.stderr)

    # ---- device times of the gather launches (plan included), CUDA-graph replays ---------------
    gtimes = None
    if dtype == torch.bfloat16 and imgs == IMGS:
        try:
            fl_ = torch.empty(256 * 1024 * 1024 // 4, device=dev)
            gtimes = gather_kernel_times(dev, fl_)
            del fl_
        except Exception as e:
            print(f'[bench] gather_kernel_times failed ({type(e).__name__}: {e})', file=sys.stderr)

    # ---- north_star evidence in the driver-visible line (N = 1 only) ----------------------------
    roofline_tensor = sweep = comparator = None
    if world == 1 and not args.no_static and imgs == IMGS:
        flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
        for name, fn in (('roofline_tensor', lambda: pgraph_tensor_roofline(dev, flush)),
                         ('sweep', lambda: gather_sweep(dev, flush, hbm_peak)),
                         ('comparator', lambda: gpu_comparator(dev, pyr_host, props_host, gts, shapes))):
            try:
                val = fn()
            except Exception as e:
                print(f'[bench] {name} failed ({type(e).__name__}: {e})', file=sys.stderr)
                val = None
            if name == 'roofline_tensor':
                roofline_tensor = val
            elif name == 'sweep':
                sweep = val
            else:
                comparator = val
        del flush

    # ---- max over ranks ------------------------------------------------------------------------
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
    value = rois_per_step * world * args.steps / (ms * 1e-3)
    value_e2e = rois_per_step * world * args.steps / (ms_e2e * 1e-3)

    # ---- roofline of the own kernels -----------------------------------------------------------
    kernels = {}
    for name, (n, tot_ms) in ksum.items():
        per_step = alg.get(name)
        if not per_step or n == 0:
            continue
        total_bytes = sum(per_step) * ksteps
        gbs = total_bytes / (tot_ms * 1e-3) / 1e9
        kernels[name] = dict(launches=n, avg_ms=tot_ms / n, alg_MB_per_launch=sum(per_step) / len(per_step) / 1e6,
                             achieved_GBs=gbs, frac=gbs / hbm_peak,
                             share_of_step=(tot_ms / ksteps) / (ms / args.steps),
                             timing='CUDA events around the launch in an eager pass of the step')
        if gtimes and name in gtimes:
            # device time of plan + gather from a CUDA-graph replay with a cold L2 (the eager figure
            # above is dominated by host launch gaps between the plan's three small kernels)
            k_ = kernels[name]
            per_launch = sum(per_step) / len(per_step)
            launches_per_step = len(per_step)
            k_.update(eager_avg_ms=k_['avg_ms'], avg_ms=gtimes[name],
                      achieved_GBs=per_launch / (gtimes[name] * 1e-3) / 1e9,
                      frac=per_launch / (gtimes[name] * 1e-3) / 1e9 / hbm_peak,
                      share_of_step=gtimes[name] * launches_per_step / (ms / args.steps),
                      timing='CUDA-graph replay of the launch (forward: sampling plan included), '
                             'L2 flushed before every replay, median of 15')
    # own dense tcgen05 kernels (FC stacks, conv tower): flops / event time / measured bf16 peak
    dense_kernels = {}
    for name, (n, tot_ms) in ksum.items():
        fl = alg_flops.get(name)
        if not fl or n == 0:
            continue
        tf = sum(fl) * ksteps / (tot_ms * 1e-3) / 1e12
        dense_kernels[name] = dict(launches=n, avg_ms=tot_ms / n, gflop_per_step=sum(fl) / 1e9,
                                   achieved_TFLOPs=tf, frac_of_sustained_peak=tf / tc_peak,
                                   share_of_step=(tot_ms / ksteps) / (ms / args.steps))
    dom = max(kernels, key=lambda k: kernels[k]['share_of_step']) if kernels else None
    roofline = None
    traffic = traffic_src = None
    for fn_ in ('r02_ncu_traffic.json', 'r01_ncu_traffic.json'):
        try:
            with open(os.path.join(ROOT, 'profiles', fn_)) as f:
                traffic = json.load(f).get(dom)
            if traffic is not None:
                traffic_src = (f'profiles/{fn_}: dram__bytes_read.sum + dram__bytes_write.sum of one '
                               'ncu --set full capture of this kernel on this workload (not '
                               're-measured in this run: ncu cannot run inside the bench)')
                break
        except Exception:
            pass
    if dom:
        k = kernels[dom]
        roofline = dict(bound='hbm', kernel=dom, achieved=k['achieved_GBs'], peak=hbm_peak,
                        unit='GB/s', frac=k['frac'], traffic=traffic, traffic_source=traffic_src,
                        peak_kind=peak_kind,
                        alg_bytes_per_launch=k['alg_MB_per_launch'] * 1e6,
                        avg_ms=k['avg_ms'], share_of_step=k['share_of_step'])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cstep, crois = cpu_step_factory()
        cstep()                                   # warm-up (page-in, thread pools)
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            cstep()
        cdt = (time.perf_counter() - t0) / reps
        cpu = dict(value=crois / cdt, unit=UNIT, cores=os.cpu_count() or 1, kind='port',
                   sample=cpu_sample_desc() + f', {reps} timed passes of {cdt:.1f} s')
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps,
                warmup=max(args.warmup, 3), ms_per_step=ms / args.steps, higher_is_better=True,
                scaling='weak', vs_baseline=None, dtype=args.dtype, data='synthetic',
                config=dict(workload=WORKLOAD if imgs == IMGS else WORKLOAD.replace(
                                '2 images/GPU', f'{imgs} images/GPU'),
                            imgs_per_gpu=imgs, global_rois_per_step=rois_per_step * world,
                            parallelism=f'dp{world}' + (' + NCCL grad all-reduce' + (
                                ' (head gradients reduced next to the backward gather)'
                                if gstep is not None and gstep.early_event is not None else '')
                                if world > 1 else ''),
                            l2='inputs larger than L2: 183 MB fp32 pyramid + 94 MB bf16 weights per step',
                            pgraph_fwd_gflop=pg_flops / 1e9,
                            execution=('CUDA graph replay of forward+losses+backward' if use_graph
                                       else 'eager'),
                            eager_ms_per_step=ms_eager,
                            full_step_with_sampling=static,
                            channels_last_bf16_pyramid=nhwc,
                            inference=infer,
                            producers=producers,
                            kernel_timing='CUDA events around each own launch in an eager pass of '
                                          'the same step (events cannot be placed inside a graph)'),
                e2e=dict(value=value_e2e, unit=UNIT, h2d_bytes_per_step=h2d_bytes,
                         d2h_bytes_per_step=d2h_bytes, ms_per_step=ms_e2e / args.steps),
                gpu_launches=launches, clocks=clk, roofline=roofline, kernels=kernels,
                dense_kernels=dense_kernels, roofline_tensor=roofline_tensor, sweep_config5=sweep,
                gpu_comparator=comparator,
                cpu_baseline=cpu)
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + '\n').encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_gpu(a)

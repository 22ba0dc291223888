#!/usr/bin/env python
"""bench.py — train-step frames/sec of the AG2Vid generation hot path at CATER 256x256.

    python bench.py --gpus N --steps K --warmup W             (ours, B200)
    python bench.py --impl reference --steps K --warmup W     (CPU arm: the oracle port)

One "step" = one training iteration of the reference (scripts/train.py:440-493) on one batch of
synthetic CATER-shaped clips: generator step (both action-graph models K1, all layouts K2 fused
into their consumer convolutions, flow warp, SPADE generator K3 for the F-1 generated frames,
GAN + feature-matching + flow-warp loss through the action discriminator, backward, Adam),
discriminator step (hinge loss on fake / real, backward, Adam) and graph step (acts_to_boxes on a
16-frame graph batch, masked smooth-L1, backward, Adam).  The VGG perceptual loss is off
(--no_vgg_loss: its weights are a download).  Workload = BASELINE.json configs[2]: 256x256,
batch 2 clips per GPU, frames_per_action 4 (8 frames per GPU step).  --generator-only times
the generator step alone with an L1 surrogate loss (the round-1 bench line).

Prints ONE JSON line (rank 0).  value = frames/s with inputs resident in HBM;
e2e = the same through the public API with the pinned-host -> device copy of
every step's batch and the device -> host read of the loss inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--size', type=int, default=256)
    ap.add_argument('--batch', type=int, default=2, help='clips per GPU')
    ap.add_argument('--frames', type=int, default=4)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--conv-impl', type=int, default=0, help='0 auto, 1 mma.sync, 2 tcgen05')
    ap.add_argument('--no-cudnn-benchmark', action='store_true', help='keep cuDNN heuristics for the out-of-path convolutions')
    ap.add_argument('--generator-only', action='store_true', help='generator fwd+bwd+Adam with an L1 surrogate loss (no discriminator / graph step)')
    ap.add_argument('--no-graph', action='store_true', help='run the step eagerly instead of replaying a CUDA graph')
    ap.add_argument('--skip-peak', action='store_true', help='do not measure the cuBLAS TF32 peak (ncu launch lists); fractions then use bf16 / 2')
    return ap.parse_args()


def peaks():
    p = {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'source': 'fallback'}
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            p.update(json.load(f))
            p['source'] = 'measured'
    except Exception:
        pass
    return p


def ncu_traffic():
    """DRAM bytes of ONE launch (dram__bytes_read.sum + dram__bytes_write.sum) per (kernel, shape) key, taken
    from this round's `ncu --set full` capture by tools/ncu_summary.py and committed as
    profiles/ncu_traffic.json ({"source": file, "launches": {key: bytes}}).  bench.py cannot run ncu, so
    `roofline.traffic` cites that file; a shape without a capture reports null."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')) as f:
            return json.load(f)
    except Exception:
        return {'source': None, 'launches': {}}


def so_sha256():
    import hashlib
    from ag2video_b200 import _lib as L
    try:
        with open(L.LIB_PATH, 'rb') as f:
            return hashlib.sha256(f.read()).hexdigest()
    except Exception:
        return None


def measure_tf32_peak(dev, seconds=1.5):
    """cuBLAS TF32 GEMM 8192^3 with the protocol of MEASURED_PEAKS.json (which holds bf16 only): best of 10
    (burst) and back to back for `seconds` (sustained, the regime of a kernel inside a long step)."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn(n, n, device=dev)
        b = torch.randn(n, n, device=dev)
        c = torch.empty(n, n, device=dev)
        for _ in range(3):
            torch.matmul(a, b, out=c)
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(10):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); torch.matmul(a, b, out=c); e.record()
            torch.cuda.synchronize()
            best = min(best, s.elapsed_time(e))
        reps = max(10, int(seconds * 1e3 / best))
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            torch.matmul(a, b, out=c)
        e.record()
        torch.cuda.synchronize()
        fl = 2.0 * n ** 3
        return {'tf32_tflops': fl / (best / 1e3) / 1e12, 'tf32_tflops_sustained': fl * reps / (s.elapsed_time(e) / 1e3) / 1e12,
                'how': 'torch.matmul fp32 with allow_tf32 (cuBLAS TF32) 8192^3: best of 10 / %d back to back' % reps}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(',')])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except Exception:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx or None, 'reasons': sorted(reasons),
                'samples': len(sm)}


def surrogate_loss(out, batch):
    imgs_pred, boxes_pred = out[0], out[1]
    return (imgs_pred - batch['imgs']).abs().mean() + 10.0 * (boxes_pred - batch['boxes'])[:, 1:].abs().mean()


# ------------------------------------------------------------------ CPU arm ---
METRIC = 'train-step frames/sec at CATER 256x256'


def workload_text(args):
    if args.generator_only:
        return ('AG2Vid CATER %dx%d generator train step (GCN -> layout -> SPADE, fwd+bwd+Adam), batch %d clips/GPU x %d frames'
                % (args.size, args.size, args.batch, args.frames))
    return ('AG2Vid CATER %dx%d train iteration (generator step + discriminator step + graph step, scripts/train.py:440-493, '
            '--no_vgg_loss), batch %d clips/GPU x %d frames + graph batch %d clips x %d frames'
            % (args.size, args.size, args.batch, args.frames, args.batch, 4 * args.frames))


def bench_config(args):
    """The `config` object - IDENTICAL in both arms (`--impl ours` / `--impl reference`): it names the workload
    (BASELINE.json configs[2]); what each arm actually executes per step is in its own `cpu_baseline.sample`
    / `impl_detail`."""
    return {'workload': workload_text(args),
            'loss': ('L1 image + box surrogate (generator step only)' if args.generator_only else
                     'reference losses: GAN hinge + feature matching + flow warp (G), hinge (D), masked smooth-L1 (graph); no VGG'),
            'l2': 'per-step working set (GBs of activations) far exceeds the 126 MB L2; no explicit flush',
            'parallelism': 'dp%d (clips sharded per rank, gradient all-reduce per optimiser, SyncBN statistics)' % args.gpus}


def cpu_step_fn(size, generator_only=False, frames=4):
    """One step of the oracle (the CPU restatement of the reference, oracle/) on a bounded sample of the
    workload: ONE clip, TWO frames (one generated frame) at the full resolution - the whole iteration
    (generator, discriminator and graph step on a 1-clip x 4*frames-frame graph batch), or the generator
    step alone.  Returns (step, description, frames_counted_per_step)."""
    from ag2video_b200.config import make_opt, synthetic_batch
    from oracle import losses as oloss
    from oracle import networks as onet
    torch.manual_seed(0)
    opt = make_opt(size, batch_size=1)
    model = onet.AG2VideoModel(opt).train()
    kw = dict(lr=opt.learning_rate, betas=(opt.beta1, 0.999))
    b = synthetic_batch(B=1, F=2, image_size=size, seed=1234)
    if generator_only:
        optim = torch.optim.Adam(model.parameters(), **kw)

        def step():
            out = model(b['imgs'], b['objs'], b['triplets'], b['actions'], boxes_gt=b['boxes'], use_gt=True)
            loss = surrogate_loss(out, b)
            optim.zero_grad(set_to_none=True)
            loss.backward()
            optim.step()
            return float(loss.detach())
        what = 'generator fwd+bwd+Adam'
    else:
        netD = oloss.MultiscaleActionDiscriminator(opt).train()
        lm = oloss.LossModel(opt, netD)
        graph_ids = {id(p) for p in model.acts_to_boxes.parameters()}
        o_graph = torch.optim.Adam(model.acts_to_boxes.parameters(), **kw)
        o_gen = torch.optim.Adam([p for p in model.parameters() if id(p) not in graph_ids], **kw)
        o_d = torch.optim.Adam(netD.parameters(), **kw)
        bg = synthetic_batch(B=1, F=4 * frames, image_size=size, seed=4321, with_images=False)

        def step():
            out = model(b['imgs'], b['objs'], b['triplets'], b['actions'], boxes_gt=b['boxes'], use_gt=True)
            G = lm.compute_generator_loss(b, out)
            o_gen.zero_grad(set_to_none=True)
            G['total_loss'].backward()
            o_gen.step()
            D = lm.compute_discriminator_loss(b, out)
            o_d.zero_grad(set_to_none=True)
            D['total_img_loss'].backward()
            o_d.step()
            boxes_pred = model(None, bg['objs'], bg['triplets'], bg['actions'], boxes_gt=bg['boxes'], graph_only=True)
            GG = lm.compute_graph_loss(bg, boxes_pred)
            o_graph.zero_grad(set_to_none=True)
            GG['total_loss'].backward()
            o_graph.step()
            return float(G['total_loss'].detach())
        what = 'full iteration (G + D + graph step on a 1 clip x %d frames graph batch)' % (4 * frames)
    desc = '1 clip x 2 frames (1 generated) at %dx%d, %s, oracle/ torch CPU fp32' % (size, size, what)
    return step, desc, 2.0


def cpu_run(step, n_warm, n_steps):
    for _ in range(n_warm):
        step()
    t0 = time.perf_counter()
    for _ in range(n_steps):
        step()
    return (time.perf_counter() - t0) / max(n_steps, 1)


def cpu_pairs(threads):
    """The two CPU-runnable BASELINE configs, timed on the host cores next to the 256x256 sample
    (SURVEY.md 8d): C1 = generator fwd+bwd at 64x64, batch 2, 4 frames (median of 3 after one warm-up);
    C4 = layout + GCN microbench (10 objects, 40 edges, 256x256, D=512; 2 of the 16 frames sampled)."""
    from ag2video_b200.config import make_opt, microbench_graph, synthetic_batch
    from oracle import networks as onet
    from oracle import ops as oops
    torch.manual_seed(0)
    res = {}
    opt = make_opt(64, batch_size=2)
    model = onet.AG2VideoModel(opt).train()
    b = synthetic_batch(B=2, F=4, image_size=64, seed=1234)

    def c1():
        model.zero_grad(set_to_none=True)
        out = model(b['imgs'], b['objs'], b['triplets'], b['actions'], boxes_gt=b['boxes'], use_gt=True)
        surrogate_loss(out, b).backward()
    c1()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); c1(); ts.append(time.perf_counter() - t0)
    ts.sort()
    res['c1'] = {'config': 'BASELINE configs[0]: generator fwd+bwd, 64x64, batch 2, frames_per_action 4',
                 'cpu_s_per_step': ts[1], 'cpu_frames_per_s': 8.0 / ts[1], 'protocol': 'median of 3 after 1 warm-up'}
    # C4: layout for 2 sampled frames of the 16 (each frame is an independent call of the reference) + one GCN layer
    bb = synthetic_batch(B=1, F=2, image_size=8, seed=1, n_objects=10, with_images=False)
    vecs = torch.randn(2, 10, 512, requires_grad=True)

    def c4_layout():
        for f in range(2):
            out = oops.boxes_to_layout(vecs[f], bb['boxes'][0, f, :10], 256)
            out.sum().backward()
    c4_layout()
    t0 = time.perf_counter(); c4_layout(); t_lay = (time.perf_counter() - t0) / 2
    layer = oops.GraphTripleConv(512, 128, 128, 128, 512)
    edges, ind = microbench_graph(B=2, O=10)
    obj = torch.randn(2, 11, 512, requires_grad=True)
    pred = torch.randn(2, 40, 128, requires_grad=True)

    def c4_gcn():
        o, p_ = layer(obj, pred, edges, ind)
        (o.sum() + p_.sum()).backward()
    c4_gcn()
    t0 = time.perf_counter()
    for _ in range(5):
        c4_gcn()
    t_gcn = (time.perf_counter() - t0) / 5
    res['c4'] = {'config': 'BASELINE configs[3]: layout + GCN microbench, 10 objects, 40 edges, 16 frames, 256x256, fwd+bwd',
                 'cpu_layout_ms_per_frame': t_lay * 1e3, 'cpu_gcn_layer_ms': t_gcn * 1e3,
                 'cpu_ms_16_frames': (16 * t_lay + t_gcn) * 1e3,
                 'protocol': 'boxes_to_layout fwd+bwd on 2 of the 16 frames (independent calls), GraphTripleConv layer 0 fwd+bwd x5'}
    res['cores'] = threads
    return res


def gpu_pairs(dev):
    """The same two configs on the B200 through the product path (CUDA events, after warm-up)."""
    from ag2video_b200.config import make_opt, microbench_graph, synthetic_batch
    from ag2video_b200.graph import GraphTripleConv
    from ag2video_b200.layout import boxes_to_layout_batched
    from ag2video_b200.networks import AG2VideoModel
    res = {}
    torch.manual_seed(0)
    model = AG2VideoModel(make_opt(64, batch_size=2), dev).train()
    b = synthetic_batch(B=2, F=4, image_size=64, seed=1234, device=dev)

    def c1():
        model.zero_grad(set_to_none=True)
        out = model(b['imgs'], b['objs'], b['triplets'], b['actions'], boxes_gt=b['boxes'], use_gt=True)
        surrogate_loss(out, b).backward()

    def ev_time(fn, n, warm):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(n):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        ts.sort()
        return ts[len(ts) // 2]
    ms = ev_time(c1, 10, 3)
    res['c1'] = {'gpu_ms_per_step': ms, 'gpu_frames_per_s': 8.0 / (ms / 1e3), 'gpu_protocol': 'eager (no CUDA graph), median of 10 after 3 warm-up'}
    del model
    bb = synthetic_batch(B=16, F=1, image_size=8, seed=1, n_objects=10, with_images=False)
    boxes = bb['boxes'].reshape(16, -1, 4).to(dev)
    valid = torch.ones(16, boxes.shape[1], dtype=torch.bool, device=dev)
    valid[:, -1] = False
    vecs = torch.randn(16, boxes.shape[1], 512, device=dev, requires_grad=True)

    def c4_layout():
        out = boxes_to_layout_batched(vecs, boxes, valid, 256)
        torch.autograd.grad(out, vecs, out)
    t_lay = ev_time(c4_layout, 10, 3)
    layer = GraphTripleConv(512, 128, 128, 128, 512).to(dev)
    edges, ind = microbench_graph(B=2, O=10)
    edges, ind = edges.to(dev), ind.to(dev)
    obj = torch.randn(2, 11, 512, device=dev, requires_grad=True)
    pred = torch.randn(2, 40, 128, device=dev, requires_grad=True)

    def c4_gcn():
        o, p_ = layer(obj, pred, edges, ind)
        torch.autograd.grad((o, p_), [obj, pred] + list(layer.parameters()), (o, p_))
    t_gcn = ev_time(c4_gcn, 10, 3)
    res['c4'] = {'gpu_layout_ms_16_frames': t_lay, 'gpu_gcn_layer_ms': t_gcn, 'gpu_ms_16_frames': t_lay + t_gcn,
                 'gpu_protocol': 'K2 fwd+bwd for all 16 frames in one launch each, K1 layer 0 fwd+bwd; median of 10'}
    torch.cuda.empty_cache()
    return res


def run_reference(args):
    """The reference arm: the oracle port on the host cores (the reference is pure Python on torch and does
    not import under torch >= 2 without shims), EXACTLY --warmup + --steps steps, each a
    bounded sample of the workload.  Rank 0 alone runs it."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    step, desc, frames = cpu_step_fn(args.size, generator_only=args.generator_only, frames=args.frames)
    dt = cpu_run(step, args.warmup, args.steps)
    value = frames / dt
    sample = '%s; %d warm-up + %d timed steps of %.1f s' % (desc, args.warmup, args.steps, dt)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'frames/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': bench_config(args),
        'cpu_baseline': {'value': value, 'unit': 'frames/s', 'cores': threads, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------ GPU arm ---
def run_ours(args):
    from ag2video_b200 import _lib as L
    from ag2video_b200 import dist as agdist
    from ag2video_b200 import spade as sp
    from ag2video_b200.config import make_opt, synthetic_batch
    from ag2video_b200.networks import AG2VideoModel
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)')
    os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')      # keep stdout for the one JSON line
    rank, world, local = agdist.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    # everything (warm-up, capture, replay) runs on one non-default stream: autograd's
    # AccumulateGrad nodes remember the stream of their first backward, and CUDA-graph capture
    # cannot happen on the legacy default stream
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    L.check(L.lib().ag2v_check_device())
    # the library convolutions outside the path (cuDNN): let it pick its algorithms during warm-up
    torch.backends.cudnn.benchmark = not args.no_cudnn_benchmark
    sp.CONV_IMPL = args.conv_impl
    # AG2V_DIAG (comma list; diagnosis of the multi-GPU overhead only, reported in impl_detail, never a bench number):
    #   nosyncbn = per-rank BN statistics, nograd = no gradient all-reduce
    diag = [d for d in os.environ.get('AG2V_DIAG', '').split(',') if d]
    if world > 1 and 'nosyncbn' not in diag:
        sp.set_sync_bn(True)

    cpu_base, pairs = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # bounded CPU work (about 20 s): one step of the 256x256 sample + the two CPU-runnable BASELINE configs
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        step, desc, fr = cpu_step_fn(args.size, generator_only=args.generator_only, frames=args.frames)
        dt = cpu_run(step, 0, 1)
        cpu_base = {'value': fr / dt, 'unit': 'frames/s', 'cores': threads, 'kind': 'port',
                    'sample': '%s; 1 timed step of %.1f s (no warm-up)' % (desc, dt)}
        del step
        pairs = cpu_pairs(threads)
        for k, v in gpu_pairs(dev).items():
            pairs[k].update(v)
        pairs['c1']['same_config'] = pairs['c4']['same_config'] = True
        pairs['c1']['gpu_over_cpu'] = pairs['c1']['gpu_frames_per_s'] / pairs['c1']['cpu_frames_per_s']
        pairs['c4']['gpu_over_cpu'] = pairs['c4']['cpu_ms_16_frames'] / pairs['c4']['gpu_ms_16_frames']
        cpu_base['pairs'] = pairs
    tf32 = None
    if rank == 0:
        if args.skip_peak:
            half = peaks()['bf16_tflops_sustained'] / 2.0
            tf32 = {'tf32_tflops': half, 'tf32_tflops_sustained': half, 'how': 'not measured (--skip-peak): half of the bf16 sustained peak'}
        else:
            tf32 = measure_tf32_peak(dev)

    torch.manual_seed(0)
    opt = make_opt(args.size, batch_size=args.batch, frames_per_action=args.frames)
    model = AG2VideoModel(opt, dev).train()
    trainer = None
    if args.generator_only:
        graph_params = list(model.acts_to_boxes.parameters())
        gen_params = list(model.acts_to_objs.parameters()) + list(model.layout_to_video.parameters())
        opt_graph = torch.optim.Adam(graph_params, lr=opt.learning_rate, betas=(opt.beta1, 0.999), fused=True, capturable=True)
        opt_gen = torch.optim.Adam(gen_params, lr=opt.learning_rate, betas=(opt.beta1, 0.999), fused=True, capturable=True)
        buckets = agdist.GradBuckets(graph_params + gen_params) if world > 1 else None
    else:
        from ag2video_b200.discriminator import MetaDiscriminatorModel
        from ag2video_b200.losses import LossModel
        from ag2video_b200.trainer import Trainer
        discriminator = MetaDiscriminatorModel(opt, dev)
        trainer = Trainer(opt, model, discriminator, LossModel(opt, discriminator), world=1 if 'nograd' in diag else world)

    # a small pool of distinct clips per rank in pinned host memory (disjoint seeds per rank = sharding by clip);
    # keys g_* = the long graph batch of the graph step (no images)
    pool_host = []
    for i in range(4):
        b = synthetic_batch(B=args.batch, F=args.frames, image_size=args.size, seed=1234 + 1000 * rank + i,
                            pad_to=(11, 6))          # CATER maxima: 10 objects + dummy, 6 actions -> static shapes
        if trainer is not None:
            g = synthetic_batch(B=args.batch, F=4 * args.frames, image_size=args.size, seed=4321 + 1000 * rank + i,
                                pad_to=(11, 6), with_images=False)
            b.update({'g_' + k: v for k, v in g.items() if v is not None})
        pool_host.append({k: v.pin_memory() for k, v in b.items()})
    pool_dev = [{k: v.to(dev) for k, v in b.items()} for b in pool_host]
    h2d_bytes = sum(v.numel() * v.element_size() for v in pool_host[0].values())
    loss_host = torch.zeros(1).pin_memory()

    def train_step(batch):
        if trainer is not None:
            clip = {k: v for k, v in batch.items() if not k.startswith('g_')}
            graph = {k[2:]: v for k, v in batch.items() if k.startswith('g_')}
            G, D, GG = trainer.iteration(clip, graph)
            return G['total_loss'].detach() + D['total_img_loss'].detach() + GG['total_loss'].detach()
        out = model(batch['imgs'], batch['objs'], batch['triplets'], batch['actions'], boxes_gt=batch['boxes'], use_gt=True)
        loss = surrogate_loss(out, batch)
        opt_graph.zero_grad(set_to_none=True)
        opt_gen.zero_grad(set_to_none=True)
        loss.backward()
        if buckets is not None:
            buckets.allreduce()
        opt_graph.step()
        opt_gen.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # The whole step (forward, backward, gradient all-reduce, both Adam updates) is captured
    # once into a CUDA graph and replayed: ~3000 kernel launches per step leave the Python /
    # launch path.  Inputs live in static device buffers that each step's batch is copied into.
    static = {k: v.clone() for k, v in pool_dev[0].items()}
    static_loss = torch.zeros(1, device=dev)
    graph, mode = None, 'eager'

    def eager_on_static():
        static_loss.copy_(train_step(static).detach().view(1))

    for i in range(max(args.warmup, 3)):
        for k, v in pool_dev[i % len(pool_dev)].items():
            static[k].copy_(v)
        eager_on_static()
    torch.cuda.synchronize()
    launches0 = L.launch_count()
    eager_on_static()
    launches_per_step = L.launch_count() - launches0
    if not args.no_graph:
        try:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=torch.cuda.current_stream()):
                eager_on_static()
            graph, mode = g, 'cuda_graph'
        except Exception as exc:          # stay eager, but say so in the output
            print('bench.py: CUDA graph capture failed (%s: %s); running eagerly' % (type(exc).__name__, exc), file=sys.stderr)
            torch.cuda.synchronize()
            graph, mode = None, 'eager (graph capture failed)'

    def run_static():
        if graph is not None:
            graph.replay()
        else:
            eager_on_static()

    def timed(n, e2e):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for i in range(n):
            if e2e:
                for k, v in pool_host[i % len(pool_host)].items():      # pinned host -> static device buffers
                    static[k].copy_(v, non_blocking=True)
                run_static()
                loss_host.copy_(static_loss, non_blocking=True)
                torch.cuda.current_stream().synchronize()                 # the user reads the loss every step
            else:
                for k, v in pool_dev[i % len(pool_dev)].items():         # device-resident inputs
                    static[k].copy_(v)
                run_static()
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    for i in range(args.warmup):
        run_static()
    torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_dev = timed(args.steps, e2e=False)
    launches = launches_per_step * args.steps
    ms_e2e = timed(args.steps, e2e=True)
    if sampler:
        sampler.stop_flag = True
        sampler.join(timeout=3)
    # per-launch CUDA events around every kernel of the path: a separate EAGER pass over the
    # same steps (events cannot be recorded inside a replayed graph)
    L.PROFILE = []
    prof_steps = min(args.steps, 3)
    for i in range(prof_steps):
        eager_on_static()
    torch.cuda.synchronize()
    prof, L.PROFILE = L.PROFILE, None

    frames = world * args.batch * args.frames
    value = frames * args.steps / (ms_dev / 1e3)
    e2e = frames * args.steps / (ms_e2e / 1e3)

    pk = peaks()
    roof, roof_all = None, None
    if prof and rank == 0:
        step_ms = ms_dev / args.steps
        tensor_kinds = ('conv3x3', 'wgrad3x3', 'k1r_recur_wgrad')       # work given in FLOPs; everything else in bytes
        tf32_peak = tf32['tf32_tflops_sustained']
        hbm_peak = pk['hbm_gbs']
        agg, by_shape = {}, {}
        for kind, work, s, e, tag in prof:
            ms_ = s.elapsed_time(e)
            for d, key in ((agg, kind), (by_shape, (kind,) + tuple(tag or ()))):
                a = d.setdefault(key, [0.0, 0.0, 0])
                a[0] += work; a[1] += ms_; a[2] += 1
        if os.environ.get('AG2V_BENCH_BREAKDOWN'):
            with open(os.environ['AG2V_BENCH_BREAKDOWN'], 'w') as f:
                f.write('kind tag... launches ms_per_step TFLOP/s-or-GB/s\n')
                for key, (w_, ms_, n_) in sorted(by_shape.items(), key=lambda kv: -kv[1][1]):
                    rate = w_ / (ms_ / 1e3) / (1e12 if key[0] in tensor_kinds else 1e9)
                    f.write('%s %d %.3f %.1f\n' % (' '.join(str(k) for k in key), n_, ms_ / prof_steps, rate))

        def entry(kind, w_, ms_, n_):
            tensor = kind in tensor_kinds
            ach = w_ / (ms_ / 1e3) / (1e12 if tensor else 1e9)
            peak = tf32_peak if tensor else hbm_peak
            return {'bound': 'tensor' if tensor else 'hbm', 'achieved': ach, 'peak': peak, 'unit': 'TFLOP/s' if tensor else 'GB/s',
                    'frac': ach / peak, 'launches_per_step': n_ / prof_steps, 'ms_per_step': ms_ / prof_steps,
                    'avg_launch_us': ms_ * 1e3 / n_, 'share_of_step': (ms_ / prof_steps) / step_ms}
        # every timed kernel family of the path, with its own bound (K1 / K2 / K4 / K7 / element-wise K3: HBM)
        roof_all = {k: entry(k, *v) for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])}
        # The dominant kernel = the kernel with the largest share of the step (all its launches: algorithmic work of
        # every launch / the sum of their durations); `dominant_shape` = its launch shape with the largest share, with
        # the DRAM traffic of one such launch from the round's `ncu --set full` capture.
        fam = max(agg, key=lambda k: agg[k][1])
        roof = entry(fam, *agg[fam])
        key = max((k for k in by_shape if k[0] == fam), key=lambda k: by_shape[k][1])
        w_, ms_, n_ = by_shape[key]
        shape = entry(key[0], w_, ms_, n_)
        batch_imgs = args.batch * (args.frames - 1)
        tr = ncu_traffic()
        tkey = ' '.join(str(k) for k in key + (batch_imgs,))
        shape.update({'kernel': '%s (%d images)' % (' '.join(str(k) for k in key), batch_imgs), 'work_per_launch': w_ / n_,
                      'traffic': tr['launches'].get(tkey)})
        kernel_names = {'conv3x3': 'conv3x3_tc_kernel', 'wgrad3x3': 'wgrad3x3_tc_kernel'}
        roof.update({
            'kernel': '%s, all %d launches of a step (every shape)' % (kernel_names.get(fam, fam), round(agg[fam][2] / prof_steps)),
            'dominant_shape': shape,
            'traffic': shape['traffic'], 'traffic_basis': ('dominant_shape: dram__bytes_read.sum + dram__bytes_write.sum of one launch, '
                                                          'ncu --set full: %s' % tr['source']) if shape['traffic'] else None,
            'work_per_launch': agg[fam][0] / agg[fam][2],
            'peak_basis': ('cuBLAS TF32 8192^3 sustained, measured in this run (%.1f TFLOP/s; burst %.1f); MEASURED_PEAKS.json (%s) '
                           'holds bf16 only: %.1f sustained' % (tf32_peak, tf32['tf32_tflops'], pk['source'], pk['bf16_tflops_sustained'])),
            'frac_of_half_bf16_peak': roof['achieved'] / (pk['bf16_tflops_sustained'] / 2.0) if roof['bound'] == 'tensor' else None,
            'measured_in': 'eager pass of %d steps, CUDA events around every launch' % prof_steps})
    def shutdown():
        # Every rank leaves together and WITHOUT tearing NCCL down: destroying a communicator whose
        # collectives live inside a captured CUDA graph (and the interpreter's own teardown order
        # afterwards) hung the job after the result line had been printed.  os._exit skips both.
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)

    from ag2video_b200 import peer
    peer_status = peer.status()          # 'peer': csrc/k8_peer.cu over NVLink windows; 'group': NCCL all-reduce
    bucket_stats = None
    if trainer is not None and getattr(trainer, 'buckets', None):
        bucket_stats = {k: {'buckets': len(b.buckets), 'mb': round(sum(f.numel() for f in b.flat if f is not None) * 4 / 2**20, 1),
                            'launched_by_hook': b.launched_by_hook, 'launched_at_end': b.launched_at_end}
                        for k, b in trainer.buckets.items()}
    if rank != 0:
        shutdown()
        return
    line = {
        'metric': METRIC, 'value': value, 'unit': 'frames/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_dev / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'tf32 GEMMs / f32 elsewhere', 'data': 'synthetic',
        'config': bench_config(args),
        'impl_detail': {'conv_impl': args.conv_impl, 'step_execution': mode, 'native_so_sha256': so_sha256(),
                        'syncbn_collective': peer_status, 'gradient_buckets': bucket_stats,
                        'gradient_allreduce': ('copy engines over IPC windows (K8b)' if trainer is not None and getattr(trainer, 'buckets', None) and any(b.ce is not None for b in trainer.buckets.values()) else 'nccl') if world > 1 else None, **({'DIAGNOSIS_ONLY': diag} if diag else {})},
        'e2e': {'value': e2e, 'unit': 'frames/s', 'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': 4,
                'ms_per_step': ms_e2e / args.steps},
        'gpu_launches': int(launches),
        'clocks': sampler.summary() if sampler else None,
        'roofline': roof,
        'roofline_kernels': roof_all,
        'tf32_peak_measured': tf32,
        'cpu_baseline': cpu_base,
    }
    print(json.dumps(line), flush=True)
    shutdown()


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()

"""Per-kernel timings with CUDA events (BASELINE config 4 shapes and friends).
Usage on the GPU box:  python tools/microbench.py [k1] [k2] [k3] > gpurun_out/micro.json
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

PEAKS = {'hbm_gbs': 6450.9, 'bf16_tflops': 1647.1}
try:
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')) as f:
        PEAKS.update(json.load(f))
except Exception:
    pass

_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
    _flush.zero_()


def time_fn(fn, iters=20, warmup=5, flush=True):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            flush_l2()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return {'us_median': ts[len(ts) // 2], 'us_min': ts[0]}


def bench_k2(res):
    from ag2video_b200.config import synthetic_batch
    from ag2video_b200.layout import boxes_to_layout_batched
    for name, N, D, H in [('c4_16x512x256', 16, 512, 256), ('c3_gen_8x512x256', 8, 512, 256),
                          ('c3_disc_8x256x256', 8, 256, 256), ('single_1x512x256', 1, 512, 256),
                          ('c2_32x512x128', 32, 512, 128)]:
        b = synthetic_batch(B=N, F=1, image_size=8, seed=1, n_objects=10, with_images=False)
        boxes = b['boxes'].reshape(N, -1, 4).cuda()
        O = boxes.shape[1]
        valid = torch.ones(N, O, dtype=torch.bool, device='cuda')
        valid[:, -1] = False
        vecs = torch.randn(N, O, D, device='cuda', requires_grad=True)
        out = boxes_to_layout_batched(vecs, boxes, valid, H)
        cot = torch.randn_like(out)
        nbytes = 4.0 * N * D * H * H
        with torch.no_grad():
            t = time_fn(lambda: boxes_to_layout_batched(vecs, boxes, valid, H))
        t['GBps'] = nbytes / t['us_median'] / 1e3
        t['frac_hbm'] = t['GBps'] / PEAKS['hbm_gbs']
        res['k2_fwd_' + name] = t
        t = time_fn(lambda: torch.autograd.grad(out, vecs, cot, retain_graph=True))
        t['GBps_dense'] = nbytes / t['us_median'] / 1e3
        t['frac_hbm_dense'] = t['GBps_dense'] / PEAKS['hbm_gbs']
        res['k2_bwd_' + name] = t
        del out, cot


def bench_k1(res):
    from ag2video_b200.config import microbench_graph
    from ag2video_b200.graph import GraphTripleConv
    for name, Din, B, O, E in [('layer0_c4', 512, 2, 11, 40), ('layer1_c4', 128, 2, 11, 40),
                               ('layer0_c3', 512, 2, 11, 16), ('layer0_b8', 512, 8, 11, 16)]:
        m = GraphTripleConv(Din, 128, 128, 128, 512).cuda()
        if E == 40:
            edges, ind = microbench_graph(B=B, O=O - 1)
        else:
            edges = torch.randint(0, O, (B, E, 2))
            ind = torch.ones(B, E, dtype=torch.bool)
        edges, ind = edges.cuda(), ind.cuda()
        obj = torch.randn(B, O, Din, device='cuda', requires_grad=True)
        pred = torch.randn(B, E, 128, device='cuda', requires_grad=True)
        params = sum(p.numel() for p in m.parameters())
        fwd_bytes = 4.0 * (params + B * O * Din + B * E * 128 + 2 * B * E + B * O * 128 + B * E * 128)
        for flush in (True, False):
            tag = 'cold' if flush else 'warmL2'
            with torch.no_grad():
                t = time_fn(lambda: m(obj, pred, edges, ind), flush=flush)
            t['GBps'] = fwd_bytes / t['us_median'] / 1e3
            t['frac_hbm'] = t['GBps'] / PEAKS['hbm_gbs']
            res['k1_fwd_%s_%s' % (name, tag)] = t
            o, p = m(obj, pred, edges, ind)
            c1, c2 = torch.randn_like(o), torch.randn_like(p)
            t = time_fn(lambda: torch.autograd.grad((o, p), [obj, pred] + list(m.parameters()), (c1, c2), retain_graph=True), flush=flush)
            t['GBps'] = 2 * fwd_bytes / t['us_median'] / 1e3
            t['frac_hbm'] = t['GBps'] / PEAKS['hbm_gbs']
            res['k1_bwd_%s_%s' % (name, tag)] = t


def bench_k3(res):
    import ag2video_b200.spade as sp
    for name, C, L, r, Hs, B in [('up3_n0_c128_r256', 128, 512, 256, 256, 2), ('up1_c512_r64', 512, 512, 64, 256, 2),
                                 ('mid_c1024_r16', 1024, 512, 16, 256, 2)]:
        m = sp.SPADE('spadesyncbatch3x3', C, L).cuda().train()
        m.fused_slope = 0.2
        x = torch.randn(B, C, r, r, device='cuda').contiguous(memory_format=torch.channels_last).requires_grad_()
        seg = torch.randn(B, L, Hs, Hs, device='cuda').contiguous(memory_format=torch.channels_last).requires_grad_()
        flops = 2.0 * 9 * B * r * r * (L * 128 + 128 * 2 * C)
        for impl in (1, 0):
            sp.CONV_IMPL = impl
            tag = 'mma' if impl == 1 else 'auto'
            with torch.no_grad():
                t = time_fn(lambda: m(x, seg), iters=10, warmup=3, flush=False)
            t['TFLOPs'] = flops / t['us_median'] / 1e6
            res['k3_fwd_%s_%s' % (name, tag)] = t
            out = m(x, seg)
            cot = torch.randn_like(out)
            t = time_fn(lambda: torch.autograd.grad(out, [x, seg] + list(m.parameters()), cot, retain_graph=True),
                        iters=10, warmup=3, flush=False)
            t['TFLOPs'] = 2 * flops / t['us_median'] / 1e6
            res['k3_bwd_%s_%s' % (name, tag)] = t
            del out, cot
        sp.CONV_IMPL = 0


def main():
    which = sys.argv[1:] or ['k1', 'k2']
    res = {'peaks': PEAKS, 'device': torch.cuda.get_device_name(0)}
    for w in which:
        globals()['bench_' + w](res)
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()

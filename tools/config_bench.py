"""BASELINE.json configs[4]: SPADE generator-only stress - 512x512, 64-channel layout, batch 16 per GPU,
forward + backward through the K3 kernels (CUDA events, median of 5 after 2 warm-ups).
Usage on the GPU box:  python tools/config_bench.py [batch] > gpurun_out/c5.json
Algorithmic FLOPs: per SPADE instance fwd 2*9*B*r^2*(L*128 + 128*2C); backward = 2x."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ag2video_b200.config import make_opt  # noqa: E402
from ag2video_b200.networks import SPADEGenerator  # noqa: E402

dev = torch.device('cuda', 0)
torch.cuda.set_per_process_memory_fraction(0.92, 0)
res = {'config': 'SPADE generator-only stress: 512x512, 64-channel layout', 'device': torch.cuda.get_device_name(0)}
peak = 1400.0
try:
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')) as f:
        peak = json.load(f)['bf16_tflops_sustained']
except Exception:
    pass


def spade_flops(B, size, L):
    nf, total = 64, 0.0
    blocks = [(16, 16, size // 32), (16, 16, size // 16), (16, 16, size // 16), (16, 8, size // 8), (8, 4, size // 4),
              (4, 2, size // 2), (2, 1, size)]
    for fin, fout, r in blocks:
        fin, fout = fin * nf, fout * nf
        mid = min(fin, fout)
        norms = [fin, mid] + ([fin] if fin != fout else [])
        for C in norms:
            total += 2.0 * 9 * B * r * r * (L * 128 + 128 * 2 * C)
    return total


for B in ([int(sys.argv[1])] if len(sys.argv) > 1 else [16, 8, 4]):
    try:
        opt = make_opt(512, batch_size=B, embedding_dim=16)          # semantic_nc = 4 * 16 = 64
        net = SPADEGenerator(opt).to(dev).to(memory_format=torch.channels_last).train()
        seg = torch.randn(B, opt.semantic_nc, 512, 512, device=dev).contiguous(memory_format=torch.channels_last)

        def step():
            net.zero_grad(set_to_none=True)
            out = net(seg)
            out.mean().backward()

        for _ in range(2):
            step()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); step(); e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        ts.sort()
        ms = ts[len(ts) // 2]
        fl = spade_flops(B, 512, opt.semantic_nc)
        res['batch_%d' % B] = {'ms_fwd_bwd': ms, 'images_per_s': B / (ms / 1e3), 'spade_gemm_TFLOP_fwd': fl / 1e12,
                               'spade_gemm_TFLOPs_fwd_bwd': 3 * fl / (ms / 1e3) / 1e12,
                               'frac_of_tf32_peak': 3 * fl / (ms / 1e3) / 1e12 / (peak / 2),
                               'peak_mem_GB': torch.cuda.max_memory_allocated() / 2 ** 30}
        break
    except torch.OutOfMemoryError as exc:
        res['batch_%d' % B] = {'error': 'out of memory: %s' % str(exc)[:120]}
        del exc
        net = seg = None
        torch.cuda.empty_cache()
print(json.dumps(res, indent=1))

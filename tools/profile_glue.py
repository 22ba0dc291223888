"""Which Python call sites of the package launch the library / torch kernels that are NOT this library's own
(copies, adds, fills, reductions, cuDNN ...): device time per (aten op, innermost ag2video_b200 frame) over one
training iteration.  python tools/profile_glue.py > gpurun_out/glue.txt"""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ag2video_b200.config import make_opt, synthetic_batch  # noqa: E402
from ag2video_b200.discriminator import MetaDiscriminatorModel  # noqa: E402
from ag2video_b200.losses import LossModel  # noqa: E402
from ag2video_b200.networks import AG2VideoModel  # noqa: E402
from ag2video_b200.trainer import Trainer  # noqa: E402

dev = torch.device('cuda', 0)
torch.cuda.set_stream(torch.cuda.Stream(device=dev))
opt = make_opt(256, batch_size=2)
model = AG2VideoModel(opt, dev).train()
meta = MetaDiscriminatorModel(opt, dev)
trainer = Trainer(opt, model, meta, LossModel(opt, meta))
b = synthetic_batch(B=2, F=4, image_size=256, seed=1, device=dev, pad_to=(11, 6))
bg = {k: v for k, v in synthetic_batch(B=2, F=16, image_size=256, seed=2, device=dev, pad_to=(11, 6), with_images=False).items()
      if v is not None}
for _ in range(3):
    trainer.iteration(b, bg)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], with_stack=True, record_shapes=True) as prof:
    trainer.iteration(b, bg)
    torch.cuda.synchronize()

by_site = collections.defaultdict(lambda: [0.0, 0])
by_op = collections.defaultdict(lambda: [0.0, 0])
for e in prof.key_averages(group_by_input_shape=True, group_by_stack_n=12):
    if e.device_type != torch.autograd.DeviceType.CPU:
        continue
    t = e.self_device_time_total
    if t <= 0:
        continue
    site = 'autograd / other'
    for fr in e.stack:
        if 'ag2video_b200' in fr or 'bench.py' in fr:
            site = fr.split('ag2video_b200/')[-1].strip()
            break
    shapes = str(e.input_shapes)[:70]
    by_site[(e.key, site, shapes)][0] += t
    by_site[(e.key, site, shapes)][1] += e.count
    by_op[e.key][0] += t
    by_op[e.key][1] += e.count
print('device time by op (us, calls):')
for k, (t, n) in sorted(by_op.items(), key=lambda kv: -kv[1][0])[:40]:
    print('%10.1f %5d  %s' % (t, n, k[:100]))
print()
print('device time by (op, call site, input shapes):')
for (name, site, shapes), (t, n) in sorted(by_site.items(), key=lambda kv: -kv[1][0])[:110]:
    print('%9.1f %4d  %-28s %-62s %s' % (t, n, name[:28], site[:62], shapes))

"""Kernel-time breakdown of one bench step (torch.profiler / CUPTI), top kernels by GPU time.
Usage on the GPU box: python tools/profile_step.py [size] [batch] [iteration|generator] > gpurun_out/step_profile.txt"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ag2video_b200.config import make_opt, synthetic_batch  # noqa: E402
from ag2video_b200.networks import AG2VideoModel  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
what = sys.argv[3] if len(sys.argv) > 3 else 'iteration'          # iteration | generator
world = int(os.environ.get('WORLD_SIZE', '1'))          # under torchrun: the data-parallel step, rank 0 reports
rank = int(os.environ.get('RANK', '0'))
local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    import torch.distributed as dist
    import ag2video_b200.spade as sp
    dist.init_process_group('nccl', device_id=dev)
    sp.set_sync_bn(True)
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
opt = make_opt(size, batch_size=B)
model = AG2VideoModel(opt, dev).train()
b = synthetic_batch(B=B, F=4, image_size=size, seed=1 + 100 * rank, device=dev, pad_to=(11, 6))
if what == 'generator':
    optim = torch.optim.Adam(model.parameters(), lr=1e-4, betas=(0.5, 0.999), fused=True)

    def step():
        out = model(b['imgs'], b['objs'], b['triplets'], b['actions'], boxes_gt=b['boxes'], use_gt=True)
        loss = (out[0] - b['imgs']).abs().mean() + 10 * (out[1] - b['boxes'])[:, 1:].abs().mean()
        optim.zero_grad(set_to_none=True)
        loss.backward()
        optim.step()
else:
    from ag2video_b200.discriminator import MetaDiscriminatorModel
    from ag2video_b200.losses import LossModel
    from ag2video_b200.trainer import Trainer
    meta = MetaDiscriminatorModel(opt, dev)
    trainer = Trainer(opt, model, meta, LossModel(opt, meta), world=world)
    bg = synthetic_batch(B=B, F=16, image_size=size, seed=2, device=dev, pad_to=(11, 6), with_images=False)
    bg = {k: v for k, v in bg.items() if v is not None}

    def step():
        trainer.iteration(b, bg)


for _ in range(2):
    step()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
if world > 1:
    dist.barrier()
    if rank != 0:
        os._exit(0)
from torch.autograd import DeviceType  # noqa: E402
rows = [(e.self_device_time_total, e.count, e.key) for e in prof.key_averages() if e.device_type == DeviceType.CUDA]
rows.sort(reverse=True)
total = sum(r[0] for r in rows)
print('GPU kernels of one step: %.3f ms in %d launches' % (total / 1e3, sum(r[1] for r in rows)))
print('%10s %6s %6s  %s' % ('us', 'share', 'calls', 'kernel'))
for t, n, k in rows[:150]:
    print('%10.1f %5.1f%% %6d  %s' % (t, 100.0 * t / total, n, k[:150]))
print()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=60, max_name_column_width=90))
if world > 1:
    sys.stdout.flush()
    os._exit(0)

"""A short run that launches each hot kernel a few times at BASELINE shapes, for ncu:
  ncu --set full --clock-control none --import-source on -k regex:<pattern> -c N -o gpurun_out/prof python tools/ncu_targets.py <what>
what: spade | layout | gcn | disc | step (one whole train step at 256x256: K4 layout conv, K5 spectral norm, grouped SPADE)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
what = sys.argv[1] if len(sys.argv) > 1 else 'spade'
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2

if what == 'spade':
    import ag2video_b200.spade as sp
    C, L, r, B = (int(a) for a in (sys.argv[3:7] if len(sys.argv) > 6 else (128, 512, 256, 2)))
    m = sp.SPADE('spadesyncbatch3x3', C, L).cuda().train()
    m.fused_slope = 0.2
    x = torch.randn(B, C, r, r, device='cuda').contiguous(memory_format=torch.channels_last).requires_grad_()
    seg = torch.randn(B, L, 256, 256, device='cuda').contiguous(memory_format=torch.channels_last).requires_grad_()
    for _ in range(reps):
        out = m(x, seg)
        out.backward(torch.ones_like(out))
elif what == 'layout':
    from ag2video_b200.config import synthetic_batch
    from ag2video_b200.layout import boxes_to_layout_batched
    N = 8
    b = synthetic_batch(B=N, F=1, image_size=8, seed=1, n_objects=10, with_images=False)
    boxes = b['boxes'].reshape(N, -1, 4).cuda()
    valid = torch.ones(N, boxes.shape[1], dtype=torch.bool, device='cuda')
    valid[:, -1] = False
    vecs = torch.randn(N, boxes.shape[1], 512, device='cuda', requires_grad=True)
    for _ in range(reps):
        out = boxes_to_layout_batched(vecs, boxes, valid, 256)
        out.backward(torch.ones_like(out))
elif what == 'gcn':
    from ag2video_b200.config import microbench_graph
    from ag2video_b200.graph import GraphTripleConv
    m = GraphTripleConv(512, 128, 128, 128, 512).cuda()
    edges, ind = microbench_graph(B=2, O=10)
    obj = torch.randn(2, 11, 512, device='cuda', requires_grad=True)
    pred = torch.randn(2, 40, 128, device='cuda', requires_grad=True)
    for _ in range(reps):
        o, p = m(obj, pred, edges.cuda(), ind.cuda())
        (o.sum() + p.sum()).backward()
elif what == 'disc':
    # the discriminator's conditioning + PatchGAN stems at 256x256 (K1, K7 with pooled tables) and the dense form (K2 strided)
    from ag2video_b200.config import make_opt, synthetic_batch
    from ag2video_b200.discriminator import MultiscaleActionDiscriminator
    from ag2video_b200.networks import AG2VideoModel
    dev = torch.device('cuda', 0)
    for rank1 in (True, False):
        opt = make_opt(256, batch_size=2, rank1_stem=rank1)
        model = AG2VideoModel(opt, dev).train()
        netD = MultiscaleActionDiscriminator(opt).to(dev).train()
        b = synthetic_batch(B=2, F=4, image_size=256, seed=1, device=dev, pad_to=(11, 6))
        with torch.no_grad():
            _, _, ad = model.acts_to_objs(b['objs'], b['triplets'], b['actions'], b['boxes'])
        for _ in range(reps):
            out = netD(b['imgs'][:, 1:], b['objs'], b['boxes'][:, 1:], [a[:, 1:] for a in ad])
            netD.zero_grad(set_to_none=True)
            sum(o[-1].mean() for o in out).backward()
elif what == 'step':
    from ag2video_b200.config import make_opt, synthetic_batch
    from ag2video_b200.networks import AG2VideoModel
    dev = torch.device('cuda', 0)
    model = AG2VideoModel(make_opt(256, batch_size=2), dev).train()
    b = synthetic_batch(B=2, F=4, image_size=256, seed=1, device=dev)
    for _ in range(reps):
        out = model(b['imgs'], b['objs'], b['triplets'], b['actions'], boxes_gt=b['boxes'], use_gt=True)
        loss = (out[0] - b['imgs']).abs().mean() + 10 * (out[1] - b['boxes'])[:, 1:].abs().mean()
        model.zero_grad(set_to_none=True)
        loss.backward()
torch.cuda.synchronize()

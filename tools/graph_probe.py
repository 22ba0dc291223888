"""Which piece of the step breaks CUDA-graph capture?  Captures each op on its own."""
import os
import sys
import traceback

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ag2video_b200.config import make_opt, microbench_graph, synthetic_batch  # noqa: E402


def try_capture(name, fn, mode='global'):
    try:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                fn()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode=mode):
            fn()
        g.replay()
        torch.cuda.synchronize()
        print('OK   %-28s (%s)' % (name, mode))
        return True
    except Exception as e:
        print('FAIL %-28s (%s): %s' % (name, mode, str(e).split('\n')[0][:160]))
        torch.cuda.synchronize()
        return False


def main():
    from ag2video_b200.graph import GraphTripleConv
    from ag2video_b200.layout import boxes_to_layout_batched
    import ag2video_b200.spade as sp
    from ag2video_b200.networks import AG2VideoModel
    dev = 'cuda'
    # K1
    m1 = GraphTripleConv(512, 128, 128, 128, 512).cuda()
    edges, ind = microbench_graph(B=2, O=10)
    edges, ind = edges.cuda(), ind.cuda()
    obj = torch.randn(2, 11, 512, device=dev, requires_grad=True)
    pred = torch.randn(2, 40, 128, device=dev, requires_grad=True)

    def k1_fwd():
        with torch.no_grad():
            m1(obj, pred, edges, ind)

    def k1_fb():
        o, p = m1(obj, pred, edges, ind)
        (o.sum() + p.sum()).backward()
    try_capture('K1 forward', k1_fwd)
    try_capture('K1 forward+backward', k1_fb)
    # K2
    b = synthetic_batch(B=4, F=1, image_size=8, seed=1, n_objects=10, with_images=False)
    boxes = b['boxes'].reshape(4, -1, 4).cuda()
    valid = torch.ones(4, boxes.shape[1], dtype=torch.bool, device=dev)
    vecs = torch.randn(4, boxes.shape[1], 64, device=dev, requires_grad=True)

    def k2_fb():
        boxes_to_layout_batched(vecs, boxes, valid, 64).sum().backward()
    try_capture('K2 forward+backward', k2_fb)
    # K3
    m3 = sp.SPADE('spadesyncbatch3x3', 128, 64).cuda().train()
    x = torch.randn(2, 128, 64, 64, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_()
    seg = torch.randn(2, 64, 64, 64, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_()

    def k3_fwd():
        with torch.no_grad():
            m3(x, seg)

    def k3_fb():
        m3(x, seg).sum().backward()
    try_capture('K3 SPADE forward', k3_fwd)
    try_capture('K3 SPADE forward+backward', k3_fb)
    # the model
    opt = make_opt(64, batch_size=2)
    model = AG2VideoModel(opt, torch.device(dev)).train()
    optim = torch.optim.Adam(model.parameters(), lr=1e-4, betas=(0.5, 0.999), fused=True, capturable=True)
    bb = synthetic_batch(B=2, F=4, image_size=64, seed=3, device=dev, pad_to=(11, 6))

    def g_fwd():
        with torch.no_grad():
            model(bb['imgs'], bb['objs'], bb['triplets'], bb['actions'], boxes_gt=bb['boxes'], use_gt=True)

    def a2l_fwd():
        with torch.no_grad():
            model.acts_to_objs(bb['objs'], bb['triplets'], bb['actions'], bb['boxes'])

    def g_fb():
        out = model(bb['imgs'], bb['objs'], bb['triplets'], bb['actions'], boxes_gt=bb['boxes'], use_gt=True)
        loss = (out[0] - bb['imgs']).abs().mean() + (out[1] - bb['boxes'])[:, 1:].abs().mean()
        optim.zero_grad(set_to_none=True)
        loss.backward()

    def g_step():
        g_fb()
        optim.step()
    try_capture('Acts2Layout forward', a2l_fwd)
    try_capture('generator forward', g_fwd)
    if not try_capture('generator fwd+bwd', g_fb):
        try_capture('generator fwd+bwd', g_fb, 'thread_local')
        try_capture('generator fwd+bwd', g_fb, 'relaxed')
    try_capture('generator full step', g_step)


if __name__ == '__main__':
    main()

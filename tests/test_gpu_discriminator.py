"""Discriminator-side callers and the three losses of one iteration (SURVEY.md section 8, rows
ctx / f4) on the sm_100a path, against the CPU oracle and the reference's golden losses."""
import pytest
import torch

from _util import det_state, det_tensor, golden, max_rel, rel_l2
from ag2video_b200.config import make_opt, synthetic_batch

gpu = pytest.mark.gpu


def _to(b, dev):
    return {k: (v.to(dev) if v is not None else None) for k, v in b.items()}


@gpu
@pytest.mark.parametrize('N,O,D,C,H,W', [(4, 6, 16, 3, 32, 32), (3, 11, 256, 3, 64, 64), (2, 5, 10, 2, 24, 36)])
def test_layout_cat_is_cat_of_layout(N, O, D, C, H, W):
    """layout_cat == cat([img, boxes_to_layout_batched]) bit for bit, forward and backward."""
    from ag2video_b200.layout import boxes_to_layout_batched, layout_cat
    g = torch.Generator().manual_seed(N * 100 + O)
    xy = torch.rand(N, O, 2, generator=g) * 0.7
    wh = torch.rand(N, O, 2, generator=g) * 0.3 + 0.05
    boxes = torch.cat([xy, wh], dim=-1).cuda()
    boxes[0, 0] = 0.0                                        # an all-zero box is dropped
    valid = (torch.rand(N, O, generator=g) > 0.2).cuda()
    img = torch.randn(N, C, H, W, generator=g).cuda()
    vecs = torch.randn(N, O, D, generator=g).cuda()
    cot = torch.randn(N, C + D, H, W, generator=g).cuda()
    i1, v1 = img.clone().requires_grad_(), vecs.clone().requires_grad_()
    got = layout_cat(i1, v1, boxes, valid, H, W)
    (got * cot).sum().backward()
    i2, v2 = img.clone().requires_grad_(), vecs.clone().requires_grad_()
    want = torch.cat([i2, boxes_to_layout_batched(v2, boxes, valid, H, W)], dim=1)
    (want * cot).sum().backward()
    assert torch.equal(got, want)
    assert torch.equal(i1.grad, i2.grad)
    assert torch.equal(v1.grad, v2.grad)


def _models(size, seed_g, seed_d, dev, **over):
    from ag2video_b200.discriminator import MetaDiscriminatorModel
    from ag2video_b200.losses import LossModel
    from ag2video_b200.networks import AG2VideoModel
    opt = make_opt(size, batch_size=2, **over)
    m = AG2VideoModel(opt)
    m.load_state_dict(det_state(m.state_dict(), seed_g), strict=True)
    m = m.to(dev).to(memory_format=torch.channels_last).train()
    meta = MetaDiscriminatorModel(opt, device=dev)
    meta.img_discriminator.load_state_dict(det_state(meta.img_discriminator.state_dict(), seed_d), strict=True)
    return opt, m, meta, LossModel(opt, meta)


@gpu
@pytest.mark.parametrize('rank1', [True, False])
def test_discriminator_forward_backward_matches_oracle(rank1):
    """Our discriminator (K1 graph layers; PatchGAN stems through the rank-1 layout, K7 - or K2 layouts
    written into the cat buffer + dense stem; cuDNN PatchGAN trunk in fp32) against the CPU oracle on
    the same weights: every level of both scales, and parameter gradients."""
    from oracle import losses as oloss
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device('cuda', 0)
    opt, m, meta, lm = _models(64, 61, 71, dev, rank1_stem=rank1)
    assert meta.img_discriminator.rank1_stem == rank1
    ref = oloss.MultiscaleActionDiscriminator(opt)
    ref.load_state_dict(det_state(ref.state_dict(), 71), strict=True)
    ref.train()
    b = synthetic_batch(B=2, F=4, image_size=64, seed=99)
    bc = _to(b, dev)
    with torch.no_grad():
        _, _, actions_data = m.acts_to_objs(bc['objs'], bc['triplets'], bc['actions'], bc['boxes'])
    ad_c = [a[:, 1:] for a in actions_data]
    ad_r = [a[:, 1:].cpu() for a in actions_data]
    netD = meta.img_discriminator
    got = netD(bc['imgs'][:, 1:], bc['objs'], bc['boxes'][:, 1:], ad_c)
    want = ref(b['imgs'][:, 1:], b['objs'], b['boxes'][:, 1:], ad_r)
    errs = {}
    for s in range(len(want)):
        for j in range(len(want[s])):
            errs['out%d.%d' % (s, j)] = max_rel(got[s][j], want[s][j])
    print({k: '%.2e' % v for k, v in errs.items()})
    assert max(errs.values()) <= 1e-4, errs

    # (A) everything the kernels of the path produce - stem weights, fc, graph layers, embeddings - through a
    # random cotangent on the stem outputs (level 0 of both scales): tight.
    cots = [det_tensor('disc.cot.%d' % s, want[s][0].shape, 9) for s in range(len(want))]
    sum((o[0] * c.to(dev)).sum() for o, c in zip(got, cots)).backward(retain_graph=True)
    sum((o[0] * c).sum() for o, c in zip(want, cots)).backward(retain_graph=True)
    gr = {k: p.grad.clone() for k, p in ref.named_parameters() if p.grad is not None}
    errs = {k: max_rel(p.grad, gr[k]) for k, p in netD.named_parameters() if k in gr}
    assert any(k.startswith('gconvs.1') for k in errs) and 'discriminator_1.model0.0.weight' in errs
    print({k: '%.2e' % v for k, v in errs.items()})
    assert max(errs.values()) <= 1e-4, {k: v for k, v in errs.items() if v > 1e-4}

    # (B) through the PatchGAN trunks (library convolutions + instance norm).  With these random weights the
    # trunk's parameter gradients are ill-conditioned: 1e-6 relative noise on a stem output moves them by up
    # to 1e-1 (measured on the CPU oracle itself), so this is a wiring check (signs, missing terms), not a
    # rounding check; the op-level checks above and test_rank1_stem_matches_dense_convolution are the tight ones.
    netD.zero_grad(); ref.zero_grad()
    sum(o[-1].mean() + 0.1 * o[1].abs().mean() for o in got).backward()
    sum(o[-1].mean() + 0.1 * o[1].abs().mean() for o in want).backward()
    gr = {k: p.grad for k, p in ref.named_parameters() if p.grad is not None}
    errs = {}
    for k, p in netD.named_parameters():
        if k in gr:
            errs[k] = rel_l2(p.grad, gr[k])
        else:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
    print({k: '%.2e' % v for k, v in errs.items()})
    assert max(errs.values()) <= 0.2, {k: v for k, v in errs.items() if v > 0.2}


@gpu
def test_iteration_losses_match_reference_golden():
    """One training iteration (train.py:440-493) at BASELINE config 1 through Trainer.iteration:
    the losses against the reference's own values (tests/golden/losses64.pt)."""
    from ag2video_b200.trainer import Trainer
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    c = golden('losses64.pt')
    dev = torch.device('cuda', 0)
    opt, m, meta, lm = _models(64, c['seed_g'], c['seed_d'], dev)
    tr = Trainer(opt, m, meta, lm)
    b = _to(synthetic_batch(B=2, F=4, image_size=64, seed=c['batch_seed']), dev)
    bg = _to(synthetic_batch(B=2, F=16, image_size=64, seed=c['graph_batch_seed'], with_images=False), dev)
    before = {k: v.detach().clone() for k, v in meta.img_discriminator.state_dict().items()}
    G, D, GG = tr.iteration(b, bg)
    torch.cuda.synchronize()
    errs = {}
    for name, got, want in (('G', G, c['G']), ('D', D, c['D']), ('graph', GG, c['graph'])):
        for k, v in want.items():
            errs['%s.%s' % (name, k)] = abs(float(got[k]) - float(v)) / abs(float(v))
    print({k: '%.2e' % v for k, v in errs.items()})
    # the graph loss and the real-image hinge do not pass through the TF32 SPADE stack: tight
    assert errs['graph.total_loss'] <= 1e-4 and errs['D.D_img_real'] <= 1e-3, errs
    # everything computed from the generated frames (18 TF32 SPADE layers deep, see test_gpu_generator)
    assert max(errs.values()) <= 1e-2, errs
    # all three optimisers stepped; the discriminator did not move before its own step's backward
    after = meta.img_discriminator.state_dict()
    moved = [k for k in before if before[k].is_floating_point() and not torch.equal(before[k], after[k])]
    assert any('discriminator_0.model0.0.weight' in k for k in moved) and any('gconvs.0' in k for k in moved)
    for p in tr.graph_params + tr.gen_params:
        assert p.grad is None or torch.isfinite(p.grad).all()


@gpu
@pytest.mark.parametrize('H', [32, 50, 128])
def test_rank1_stem_matches_dense_convolution(H):
    """K7 against the dense library form on the same numbers: conv4x4/s2(cat[img, layout]) at full
    resolution and on the 3x3/s2 average pool (count_include_pad=False) of the concatenation -
    forward, and gradients for the image, the object vectors and the convolution weight."""
    import torch.nn.functional as F
    from ag2video_b200.layout import (boxes_to_layout_batched, layout_sconv, layout_tables, layout_tables_avgpool,
                                      pooled_size)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    N, O, D, Co = 3, 7, 24, 64
    g = torch.Generator().manual_seed(H)
    xy = torch.rand(N, O, 2, generator=g) * 0.8 - 0.05
    wh = torch.rand(N, O, 2, generator=g) * 0.35 + 0.03
    boxes = torch.cat([xy, wh], dim=-1).cuda()
    boxes[0, 1] = 0.0
    boxes[1, 2] = torch.tensor([0.0, 0.0, 1.0, 1.0])          # a full-frame object
    valid = (torch.rand(N, O, generator=g) > 0.15).cuda()
    img0 = torch.randn(N, 3, H, H, generator=g).cuda()
    vecs0 = torch.randn(N, O, D, generator=g).cuda()
    w0 = (torch.randn(Co, 3 + D, 4, 4, generator=g) * 0.05).cuda()
    bias = torch.randn(Co, generator=g).cuda()
    pool = lambda x: F.avg_pool2d(x, kernel_size=3, stride=2, padding=[1, 1], count_include_pad=False)
    for scale in (0, 1):
        img, vecs, w = img0.clone().requires_grad_(), vecs0.clone().requires_grad_(), w0.clone().requires_grad_()
        x = torch.cat([img, boxes_to_layout_batched(vecs, boxes, valid, H, H)], dim=1)
        if scale:
            x = pool(x)
        want = F.conv2d(x, w, bias, stride=2, padding=2)
        cot = torch.randn(want.shape, generator=g).cuda()
        (want * cot).sum().backward()
        img2, vecs2, w2 = img0.clone().requires_grad_(), vecs0.clone().requires_grad_(), w0.clone().requires_grad_()
        tables, Hs, im = layout_tables(boxes, valid, H, H), H, img2
        if scale:
            tables, Hs, im = layout_tables_avgpool(tables, N, O, H, H), pooled_size(H), pool(img2)
        base = F.conv2d(im, w2[:, :3], bias, stride=2, padding=2).contiguous(memory_format=torch.channels_last)
        got = layout_sconv(w2[:, 3:], vecs2, tables, base, Hs, Hs)
        (got * cot).sum().backward()
        errs = dict(out=max_rel(got, want), dimg=max_rel(img2.grad, img.grad), dvecs=max_rel(vecs2.grad, vecs.grad),
                    dw=max_rel(w2.grad, w.grad))
        print(H, scale, {k: '%.2e' % v for k, v in errs.items()})
        assert max(errs.values()) <= 2e-5, (scale, errs)

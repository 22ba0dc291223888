"""Frame groups: a batch that holds G reference calls (group-major) must reproduce G successive
calls — batch-norm statistics, running-stat updates and spectral-norm power iterations per
group — for SPADE, the affine BN + LeakyReLU stage, the sigma-mode spectral norm and the whole
generator (batched frames vs the frame loop of models/spade_models/generator.py:56-94)."""
import copy

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils import spectral_norm

from _util import det_state, max_rel, rel_l2
from ag2video_b200.config import make_opt, synthetic_batch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fp32_library():
    """cuDNN / cuBLAS in fp32: what these tests compare is the grouping logic, not TF32 rounding."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _rand(*shape, seed=0, cl=False):
    t = torch.randn(*shape, generator=torch.Generator().manual_seed(seed)).cuda()
    return t.contiguous(memory_format=torch.channels_last) if cl else t


@pytest.mark.parametrize('training', [True, False])
def test_spade_groups_equal_successive_calls(training):
    import ag2video_b200.spade as sp
    G, B, C, Lc, r, Hs = 3, 2, 64, 32, 16, 32
    m1 = sp.SPADE('spadesyncbatch3x3', C, Lc).cuda()
    m1.load_state_dict(det_state(m1.state_dict(), 5))
    m1.fused_slope = 0.2
    m2 = copy.deepcopy(m1)
    m1.train(training), m2.train(training)
    x = _rand(G * B, C, r, r, seed=1, cl=True).requires_grad_()
    seg = _rand(G * B, Lc, Hs, Hs, seed=2, cl=True).requires_grad_()
    cot = _rand(G * B, C, r, r, seed=3, cl=True)
    out = m1(x, seg, groups=G)
    (out * cot).sum().backward()
    got = [out.detach(), x.grad.clone(), seg.grad.clone()] + [p.grad.clone() for p in m1.parameters()]
    x.grad = seg.grad = None
    outs = [m2(x[g * B:(g + 1) * B], seg[g * B:(g + 1) * B]) for g in range(G)]
    ref = torch.cat(outs, dim=0)
    (ref * cot).sum().backward()
    want = [ref.detach(), x.grad, seg.grad] + [p.grad for p in m2.parameters()]
    for i, (a, b) in enumerate(zip(got, want)):
        assert max_rel(a, b) <= 1e-5, (i, max_rel(a, b))
    for name in ('running_mean', 'running_var'):
        a, b = getattr(m1.param_free_norm, name), getattr(m2.param_free_norm, name)
        assert max_rel(a, b) <= 1e-6, name


@pytest.mark.parametrize('G,C,slope', [(3, 32, 0.2), (1, 512, 0.2), (2, 64, 1.0)])
@pytest.mark.parametrize('training', [True, False])
def test_bn_act_matches_torch(G, C, slope, training):
    from ag2video_b200.networks import _BN2d
    B, H = 2, 12
    bn = _BN2d(C).cuda()
    bn.load_state_dict(det_state(bn.state_dict(), 2))
    with torch.no_grad():
        bn.running_var.abs_().add_(0.5)
    ref = copy.deepcopy(bn)
    bn.train(training), ref.train(training)
    x = (_rand(G * B, C, H, H, seed=4, cl=True) * 2 + 0.3).requires_grad_()
    cot = _rand(G * B, C, H, H, seed=5, cl=True)
    y = bn(x, groups=G, slope=slope)
    (y * cot).sum().backward()
    got = [y.detach(), x.grad.clone(), bn.weight.grad.clone(), bn.bias.grad.clone()]
    x.grad = None
    parts = []
    for g in range(G):
        z = F.batch_norm(x[g * B:(g + 1) * B], ref.running_mean, ref.running_var, ref.weight, ref.bias, training, 0.1, 1e-5)
        parts.append(F.leaky_relu(z, slope) if slope != 1.0 else z)
    yr = torch.cat(parts, dim=0)
    (yr * cot).sum().backward()
    want = [yr.detach(), x.grad, ref.weight.grad, ref.bias.grad]
    for name, a, b in zip(('y', 'dx', 'dweight', 'dbias'), got, want):
        assert max_rel(a, b) <= 2e-5, (name, max_rel(a, b))
    assert max_rel(bn.running_mean, ref.running_mean) <= 1e-6
    assert max_rel(bn.running_var, ref.running_var) <= 1e-6


@pytest.mark.parametrize('training', [True, False])
def test_specnorm_sigma_mode_equals_successive_hook_calls(training):
    """conv_scaled in sigma mode on a group-major batch == G successive calls of torch's hook."""
    from ag2video_b200.specnorm import SpectralNormGroup, conv_scaled
    G, B = 3, 2
    shapes = [(64, 128, 3, True), (35, 16, 3, True), (128, 64, 1, True), (1027, 32, 3, True)]
    ref = []
    for i, (cin, cout, k, bias) in enumerate(shapes):
        torch.manual_seed(i)
        ref.append(spectral_norm(nn.Conv2d(cin, cout, k, padding=k // 2, bias=bias)).cuda().to(memory_format=torch.channels_last))
    ours = copy.deepcopy(ref)
    holder = nn.ModuleList(ours)
    group = SpectralNormGroup(holder)
    holder.train(training)
    for m in ref:
        m.train(training)
    xs = [_rand(G * B, s[0], 8, 8, seed=10 + i, cl=True) for i, s in enumerate(shapes)]
    cots = [_rand(G * B, s[1], 8, 8, seed=20 + i, cl=True) for i, s in enumerate(shapes)]
    total = 0.0
    want_out = []
    for m, x, c in zip(ref, xs, cots):
        ys = torch.cat([m(x[g * B:(g + 1) * B]) for g in range(G)], dim=0)     # G hook calls = G power iterations
        want_out.append(ys.detach())
        total = total + (ys * c).sum()
    total.backward()
    group.refresh_sigma(G, B)
    total = 0.0
    got_out = []
    for m, x, c in zip(ours, xs, cots):
        y = conv_scaled(m, x)
        got_out.append(y.detach())
        total = total + (y * c).sum()
    group.end_sigma()
    total.backward()
    for s, a, b, mo, mr in zip(shapes, got_out, want_out, ours, ref):
        assert max_rel(a, b) <= 3e-5, (s, 'out', max_rel(a, b))
        assert max_rel(mo.weight_u, mr.weight_u) <= 2e-5, (s, 'u')
        assert max_rel(mo.weight_v, mr.weight_v) <= 2e-5, (s, 'v')
        assert max_rel(mo.weight_orig.grad, mr.weight_orig.grad) <= 5e-5, (s, 'dW', max_rel(mo.weight_orig.grad, mr.weight_orig.grad))
        if mo.bias is not None:
            assert max_rel(mo.bias.grad, mr.bias.grad) <= 2e-5, (s, 'db')


def test_generator_batched_frames_equal_frame_loop():
    """Layout2VidGenerator: T-1 frames as one group-major batch vs the reference's frame loop."""
    import ag2video_b200.spade as sp
    from ag2video_b200.networks import AG2VideoModel
    old = sp.CONV_IMPL
    sp.CONV_IMPL = 3                      # fp32-class GEMMs: what is compared is the grouping logic
    try:
        res = []
        for batched in (False, True):
            m = AG2VideoModel(make_opt(64, batch_size=2))
            m.load_state_dict(det_state(m.state_dict(), 3), strict=True)
            m = m.cuda().to(memory_format=torch.channels_last).train()
            m.layout_to_video.batch_frames = batched
            b = synthetic_batch(B=2, F=4, image_size=64, seed=11, device='cuda')
            imgs, boxes, flows, conf = m(b['imgs'], b['objs'], b['triplets'], b['actions'], boxes_gt=b['boxes'], use_gt=True)[:4]
            loss = (imgs - b['imgs']).abs().mean() + flows.abs().mean() * 0.01
            loss.backward()
            res.append((imgs.detach(), flows.detach(), conf.detach(), {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None},
                        {k: v.clone() for k, v in m.state_dict().items() if 'running' in k or k.endswith(('_u', '_v'))}))
        (i0, f0, c0, g0, s0), (i1, f1, c1, g1, s1) = res
        print('batched vs loop: imgs %.2e flows %.2e' % (max_rel(i1, i0), max_rel(f1, f0)))
        # the random-weight SPADE stack amplifies fp32 rounding ~1000x (eager GPU vs CPU: 3.5e-3)
        assert max_rel(i1, i0) <= 2e-3 and max_rel(f1, f0) <= 1e-4
        assert (c0 != c1).float().mean() <= 1e-3
        assert g0.keys() == g1.keys()
        worst = max((rel_l2(g1[k], g0[k]), k) for k in g0 if g0[k].abs().max() > 1e-7)
        print('batched vs loop: worst gradient rel-L2 %.2e (%s)' % worst)
        assert worst[0] <= 0.1, worst      # LeakyReLU / ReLU gate flips bound this (see test_gpu_generator)
        for k in s0:                      # running statistics and power-iteration vectors end up identical
            assert max_rel(s1[k], s0[k]) <= 1e-4, (k, max_rel(s1[k], s0[k]))
    finally:
        sp.CONV_IMPL = old


@pytest.mark.parametrize('impl,tol', [(3, 1e-3), (0, 2e-2)])
@pytest.mark.parametrize('fin,fout', [(64, 64), (128, 64)])
def test_resnet_block_sigma_mode_equals_successive_calls(fin, fout, impl, tol):
    """SPADEResnetBlock on a group-major batch in sigma mode (3x3 convolutions on the tcgen05
    kernel with 1/sigma, bias and residual in the epilogue, d sigma out of the SPADE backward)
    vs G successive calls through the spectral-norm hooks and cuDNN."""
    import ag2video_b200.spade as sp
    G, B, r, Hs = 3, 2, 16, 32
    opt = make_opt(64)
    old = sp.CONV_IMPL
    sp.CONV_IMPL = impl
    try:
        blocks = []
        for _ in range(2):
            blk = sp.SPADEResnetBlock(fin, fout, opt)
            blk.load_state_dict(det_state(blk.state_dict(), 9))
            blocks.append(blk.cuda().to(memory_format=torch.channels_last).train())
        b1, b2 = blocks
        x = _rand(G * B, fin, r, r, seed=1, cl=True).requires_grad_()
        seg = _rand(G * B, opt.semantic_nc, Hs, Hs, seed=2, cl=True).requires_grad_()
        cot = _rand(G * B, fout, r, r, seed=3, cl=True)
        b1._sn.refresh_sigma(G, B)
        out = b1(x, seg, groups=G)
        b1._sn.end_sigma()
        (out * cot).sum().backward()
        got = {'out': out.detach(), 'dx': x.grad.clone(), 'dseg': seg.grad.clone()}
        got.update({k: p.grad.clone() for k, p in b1.named_parameters()})
        x.grad = seg.grad = None
        ref = torch.cat([b2(x[g * B:(g + 1) * B], seg[g * B:(g + 1) * B]) for g in range(G)], dim=0)
        (ref * cot).sum().backward()
        want = {'out': ref.detach(), 'dx': x.grad, 'dseg': seg.grad}
        want.update({k: p.grad for k, p in b2.named_parameters()})
        assert got.keys() == want.keys()
        for k in want:
            if k == 'conv_0.bias':        # a bias in front of a batch norm: its true gradient is zero, both are noise
                continue
            err = max_rel(got[k], want[k]) if k == 'out' else rel_l2(got[k], want[k])   # gradients: LeakyReLU gate flips
            assert err <= tol, (k, err)
        for (k, a), (_, b) in zip(b1.state_dict().items(), b2.state_dict().items()):
            if k.endswith(('_u', '_v', 'running_mean', 'running_var')):
                assert max_rel(a, b) <= (1e-4 if impl == 3 else 2e-3), (k, max_rel(a, b))
    finally:
        sp.CONV_IMPL = old


@pytest.mark.parametrize('training', [True, False])
def test_bn_act_folds_an_input_scale_into_eps(training):
    """bn(in_scale_g * z) computed as BN(z; eps / s^2): outputs, dz, d in_scale, running stats."""
    from ag2video_b200.networks import _BN2d
    G, B, C, H, slope = 3, 2, 32, 10, 0.2
    bn = _BN2d(C).cuda()
    bn.load_state_dict(det_state(bn.state_dict(), 2))
    with torch.no_grad():
        bn.running_var.abs_().add_(0.5)
    ref = copy.deepcopy(bn)
    bn.train(training), ref.train(training)
    # small activations so that eps matters and d in_scale is not pure rounding noise
    z = (_rand(G * B, C, H, H, seed=4, cl=True) * 0.02).requires_grad_()
    scale = torch.tensor([0.7, 1.3, 2.1], device='cuda', requires_grad=True)
    cot = _rand(G * B, C, H, H, seed=5, cl=True)
    y = bn(z, groups=G, slope=slope, in_scale=scale)
    (y * cot).sum().backward()
    got = [y.detach(), z.grad.clone(), scale.grad.clone(), bn.weight.grad.clone(), bn.bias.grad.clone()]
    z.grad = scale.grad = None
    parts = []
    for g in range(G):
        zz = z[g * B:(g + 1) * B] * scale[g]
        parts.append(F.leaky_relu(F.batch_norm(zz, ref.running_mean, ref.running_var, ref.weight, ref.bias, training, 0.1, 1e-5), slope))
    yr = torch.cat(parts, dim=0)
    (yr * cot).sum().backward()
    want = [yr.detach(), z.grad, scale.grad, ref.weight.grad, ref.bias.grad]
    for name, a, b in zip(('y', 'dz', 'dscale', 'dweight', 'dbias'), got, want):
        assert max_rel(a, b) <= (2e-3 if name == 'dscale' and training else 5e-5), (name, max_rel(a, b), a.flatten()[:4], b.flatten()[:4])
    assert max_rel(bn.running_mean, ref.running_mean) <= 1e-5
    assert max_rel(bn.running_var, ref.running_var) <= 1e-5


def test_packed_weights_follow_a_fused_optimizer():
    """torch.optim.Adam(fused=True) updates parameters without bumping Tensor._version; the packed
    GEMM-layout weight copies must still be rebuilt for the next forward."""
    import ag2video_b200.spade as sp
    opt = make_opt(64)
    G, B = 2, 1
    blk = sp.SPADEResnetBlock(64, 64, opt)
    blk.load_state_dict(det_state(blk.state_dict(), 9))
    blk = blk.cuda().to(memory_format=torch.channels_last).train()
    x = _rand(G * B, 64, 16, 16, seed=1, cl=True)
    seg = _rand(G * B, opt.semantic_nc, 32, 32, seed=2, cl=True)
    optim = torch.optim.Adam(blk.parameters(), lr=1e-2, fused=True)

    def run(block):
        block._sn.refresh_sigma(G, B)
        out = block(x, seg, groups=G)
        block._sn.end_sigma()
        return out
    out0 = run(blk)
    out0.square().mean().backward()
    optim.step()
    state = {k: v.clone() for k, v in blk.state_dict().items()}
    out1 = run(blk).detach()
    fresh = sp.SPADEResnetBlock(64, 64, opt)
    fresh.load_state_dict(state)
    fresh = fresh.cuda().to(memory_format=torch.channels_last).train()
    out2 = run(fresh).detach()
    assert max_rel(out1, out0.detach()) > 1e-2, 'the optimizer step did not change the output'
    assert max_rel(out1, out2) <= 1e-5, max_rel(out1, out2)


@pytest.mark.parametrize('training', [True, False])
@pytest.mark.parametrize('impl', [1, 0])
def test_spade_reads_its_input_through_the_upsampling(training, impl):
    """SPADE(x, seg, upsample=True) == SPADE(F.interpolate(x, scale_factor=2), seg), forward, backward
    (gradient of the low-resolution x = sum over its four children) and running statistics."""
    import ag2video_b200.spade as sp
    G, B, C, Lc, h, Hs = 2, 2, 64, 32, 8, 32
    old = sp.CONV_IMPL
    sp.CONV_IMPL = impl
    try:
        m1 = sp.SPADE('spadesyncbatch3x3', C, Lc).cuda()
        m1.load_state_dict(det_state(m1.state_dict(), 5))
        m1.fused_slope = 0.2
        m2 = copy.deepcopy(m1)
        m1.train(training), m2.train(training)
        x = _rand(G * B, C, h, h, seed=1, cl=True).requires_grad_()
        seg = _rand(G * B, Lc, Hs, Hs, seed=2, cl=True).requires_grad_()
        cot = _rand(G * B, C, 2 * h, 2 * h, seed=3, cl=True)
        out = m1(x, seg, groups=G, upsample=True)
        (out * cot).sum().backward()
        got = [out.detach(), x.grad.clone(), seg.grad.clone()] + [p.grad.clone() for p in m1.parameters()]
        x.grad = seg.grad = None
        ref = m2(F.interpolate(x, scale_factor=2, mode='nearest'), seg, groups=G)
        (ref * cot).sum().backward()
        want = [ref.detach(), x.grad, seg.grad] + [p.grad for p in m2.parameters()]
        for i, (a, b) in enumerate(zip(got, want)):
            assert max_rel(a, b) <= 2e-5, (i, max_rel(a, b))
        for name in ('running_mean', 'running_var'):
            a, b = getattr(m1.param_free_norm, name), getattr(m2.param_free_norm, name)
            assert max_rel(a, b) <= 1e-6, (name, max_rel(a, b))
    finally:
        sp.CONV_IMPL = old

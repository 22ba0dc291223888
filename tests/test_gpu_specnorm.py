"""K5 (csrc/k5_specnorm.cu) against torch.nn.utils.spectral_norm on the same GPU: the weight,
the updated u / v buffers and the gradient of weight_orig, for every storage order and shape
class the generator holds (3x3 / 1x1, channels_last / contiguous, odd Cin, nn.Linear)."""
import copy

import pytest
import torch
import torch.nn as nn
from torch.nn.utils import spectral_norm

from _util import max_rel

pytestmark = pytest.mark.gpu

CASES = [  # (kind, cin, cout, k, channels_last)
    ('conv', 64, 128, 3, True), ('conv', 1027, 32, 3, True), ('conv', 1027, 512, 3, False),
    ('conv', 256, 128, 1, True), ('conv', 1024, 1024, 3, True), ('conv', 35, 17, 3, True),
    ('conv', 8, 8, 7, False), ('linear', 300, 70, 0, False),
]


def _make(kind, cin, cout, k, cl, seed):
    torch.manual_seed(seed)
    m = nn.Conv2d(cin, cout, k, padding=k // 2) if kind == 'conv' else nn.Linear(cin, cout)
    m = spectral_norm(m).cuda()
    if cl:
        m = m.to(memory_format=torch.channels_last)
    return m


def _run(mods, x_list, g_list, training):
    outs, total = [], 0.0
    for m, x, g in zip(mods, x_list, g_list):
        m.train(training)
        m.zero_grad()
        total = total + (m(x) * g).sum()
        outs.append((m.weight.detach().clone(), m.weight_u.clone(), m.weight_v.clone()))
    total.backward()                 # one backward: the batched op is a single autograd node
    return outs, [m.weight_orig.grad.clone() for m in mods]


@pytest.mark.parametrize('training', [True, False])
def test_specnorm_matches_torch(training):
    from ag2video_b200.specnorm import SpectralNormGroup
    ref = [_make(*c, seed=i) for i, c in enumerate(CASES)]
    ours = copy.deepcopy(ref)
    holder = nn.ModuleList(ours)
    group = SpectralNormGroup(holder)
    assert len(group.entries) == len(CASES)
    torch.manual_seed(99)
    xs, gs = [], []
    for (kind, cin, cout, k, cl), m in zip(CASES, ref):
        x = torch.randn(2, cin, 8, 8, device='cuda') if kind == 'conv' else torch.randn(5, cin, device='cuda')
        xs.append(x)
        with torch.no_grad():
            gs.append(torch.randn_like(m(x)))
    for m_ref, m_our in zip(ref, ours):           # the probing call above advanced u/v of ref only
        m_our.load_state_dict(m_ref.state_dict())
    for rounds in range(2):                        # two calls: the buffers carry over
        o_ref, g_ref = _run(ref, xs, gs, training)
        holder.train(training)
        group.refresh()
        o_our, g_our = _run(ours, xs, gs, training)
        for c, a, b, ga, gb in zip(CASES, o_ref, o_our, g_ref, g_our):
            assert b[0].stride() == a[0].stride(), c
            for name, ta, tb in (('weight', a[0], b[0]), ('u', a[1], b[1]), ('v', a[2], b[2]), ('grad', ga, gb)):
                err = max_rel(tb, ta)
                assert err <= 2e-5, (c, rounds, name, err)


def test_specnorm_standalone_and_deterministic():
    """A module outside any refresh() computes its own weight; two runs give identical bits."""
    from ag2video_b200.specnorm import SpectralNormGroup
    res = []
    for _ in range(2):
        m = _make('conv', 128, 64, 3, True, seed=3)
        SpectralNormGroup(m)
        x = torch.randn(2, 128, 8, 8, device='cuda', generator=torch.Generator('cuda').manual_seed(1))
        y = m(x)
        y.square().sum().backward()
        res.append((m.weight.detach().clone(), m.weight_orig.grad.clone(), m.weight_u.clone()))
    for a, b in zip(*res):
        assert torch.equal(a, b)


def test_sigma_mode_per_weight_gradient_nodes_equal_the_batched_node():
    """Sigma mode (convolutions on weight_orig, outputs scaled by 1/sigma per frame group): the gradient of every
    weight through its scales is the same whether ONE batched autograd node produces it for all weights or every
    weight has its own node (specnorm.PER_WEIGHT_SIGMA_GRAD; one launch per weight, so gradients complete layer by
    layer).  Same u, v, sigma and the same rank-1 sums: 1e-6."""
    from ag2video_b200 import specnorm
    torch.manual_seed(3)
    root = nn.Sequential(spectral_norm(nn.Conv2d(16, 24, 3, padding=1)), spectral_norm(nn.Conv2d(24, 8, 3, padding=1, bias=False)),
                         spectral_norm(nn.Conv2d(8, 8, 1))).cuda()
    x = torch.randn(6, 16, 10, 10, device='cuda')                       # 3 frame groups of 2 images
    cot = torch.randn(6, 8, 10, 10, device='cuda')
    results = []
    for per_weight in (False, True):
        net = copy.deepcopy(root).train()
        group = specnorm.SpectralNormGroup(net)
        old = specnorm.PER_WEIGHT_SIGMA_GRAD
        specnorm.PER_WEIGHT_SIGMA_GRAD = per_weight
        try:
            group.refresh_sigma(3, 2)
            h = specnorm.conv_scaled(net[0], x)                          # per-image scale
            z, scale_g = specnorm.conv_unscaled(net[1], torch.relu(h))   # per-group scale handed to the consumer
            h = z * scale_g.repeat_interleave(2).view(-1, 1, 1, 1)
            y = specnorm.conv_scaled(net[2], torch.relu(h))
            group.end_sigma()
            (y * cot).sum().backward()
        finally:
            specnorm.PER_WEIGHT_SIGMA_GRAD = old
        results.append((y.detach(), [m.weight_orig.grad.clone() for m in net], [m.weight_u.clone() for m in net]))
    (y0, g0, u0), (y1, g1, u1) = results
    assert torch.equal(y0, y1) and all(torch.equal(a, b) for a, b in zip(u0, u1))
    worst = max(max_rel(b, a) for a, b in zip(g0, g1))
    print('per-weight vs batched sigma gradient: %.2e' % worst)
    assert worst <= 1e-6

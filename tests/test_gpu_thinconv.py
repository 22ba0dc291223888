"""K6 (csrc/k6_thinconv.cu) against torch: act_out(conv(leaky_relu(x))) with 2-3 output channels,
forward, input / weight / bias gradients, both weight storage orders, ragged image sizes."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from _util import max_rel

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fp32_library():
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32 = old


@pytest.mark.parametrize('ci,co,slope,act', [(64, 3, 0.2, 'tanh'), (32, 2, 1.0, None), (64, 3, 1.0, None)])
@pytest.mark.parametrize('cl', [True, False])
@pytest.mark.parametrize('B,H,W', [(2, 16, 16), (3, 13, 21), (1, 64, 64)])
def test_thin_conv_matches_torch(ci, co, slope, act, cl, B, H, W):
    from ag2video_b200.thinconv import thin_conv3x3
    g = torch.Generator().manual_seed(ci + co + H)
    conv = nn.Conv2d(ci, co, 3, padding=1).cuda()
    if cl:
        conv = conv.to(memory_format=torch.channels_last)
    x = torch.randn(B, ci, H, W, generator=g).cuda().contiguous(memory_format=torch.channels_last).requires_grad_()
    cot = torch.randn(B, co, H, W, generator=g).cuda()
    y = thin_conv3x3(conv, x, slope_in=slope, act_out=act)
    (y * cot).sum().backward()
    got = [y.detach(), x.grad.clone(), conv.weight.grad.clone(), conv.bias.grad.clone()]
    x.grad = None
    conv.zero_grad()
    z = conv(F.leaky_relu(x, slope) if slope != 1.0 else x)
    ref = torch.tanh(z) if act == 'tanh' else z
    (ref * cot).sum().backward()
    want = [ref.detach(), x.grad, conv.weight.grad, conv.bias.grad]
    for name, a, b in zip(('y', 'dx', 'dw', 'db'), got, want):
        assert max_rel(a, b) <= 2e-5, (name, max_rel(a, b))


def test_thin_conv_other_shapes_raise_unless_library_is_allowed():
    """No silent library path: a shape without a kernel raises; the explicit opt-in warns and uses cuDNN."""
    from ag2video_b200.thinconv import thin_conv3x3
    conv = nn.Conv2d(16, 5, 3, padding=1).cuda()
    x = torch.randn(1, 16, 8, 8, device='cuda')
    with pytest.raises(NotImplementedError):
        thin_conv3x3(conv, x, 0.2, 'tanh')
    with pytest.warns(RuntimeWarning):
        y = thin_conv3x3(conv, x, 0.2, 'tanh', allow_library=True)
    assert max_rel(y, torch.tanh(conv(F.leaky_relu(x, 0.2)))) <= 1e-6

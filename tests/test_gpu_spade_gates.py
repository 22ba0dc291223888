"""Gate-aware gradient parity for SPADE (K3) with RANDOM weights at the real widths.

ReLU (inside the modulation MLP) and the fused LeakyReLU make SPADE piecewise linear: an element whose
pre-activation lies within TF32 rounding of zero may take the other branch than the fp32 reference, and
its gradient then differs by O(1).  That is a property of comparing two roundings, not an error of the
backward kernels.  This test therefore differentiates the fp32 torch reference ON THE BRANCH THE KERNEL
TOOK: the reference forward is evaluated with the gates (actv > 0, out > 0) read back from our forward,
which makes it a smooth (multilinear) function around the operating point.  Every gradient - x, segmap,
all six parameter tensors - must then agree to 1e-3 (max|err| / max|ref| per tensor, the north_star bar),
and the share of gates that differ from the reference's own gates is printed and bounded.
"""
import pytest
import torch
import torch.nn.functional as F

from _util import det_tensor, load_det, max_rel

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.fixture(autouse=True)
def _exact_library_convs():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _forced_reference(x, seg, p, gate_a, gate_out, slope, eps=1e-5):
    """SPADE.forward of normalization.py:96-110 (+ the LeakyReLU of architecture.py:68 when slope != 1)
    in fp32 torch ops with the two gates given instead of computed."""
    xhat = F.batch_norm(x, None, None, None, None, True, 0.1, eps)
    seg_r = F.interpolate(seg, size=x.shape[2:], mode='nearest')
    z = F.conv2d(seg_r, p['mlp_shared.0.weight'], p['mlp_shared.0.bias'], padding=1)
    a = z * gate_a
    gamma = F.conv2d(a, p['mlp_gamma.weight'], p['mlp_gamma.bias'], padding=1)
    beta = F.conv2d(a, p['mlp_beta.weight'], p['mlp_beta.bias'], padding=1)
    pre = xhat * (1 + gamma) + beta
    return pre * gate_out, z, pre


@pytest.mark.parametrize('C,L,r,Hs,B,slope', [
    (64, 512, 256, 256, 1, 0.2),       # up_3.norm_1 at BASELINE config 3 (one image)
    (128, 512, 128, 256, 2, 0.2),      # up_2.norm_1, segmap read with stride 2
    (1024, 512, 8, 256, 2, 1.0),       # head_0.norm_s-like: no activation fused, stride 32
    (256, 64, 64, 64, 3, 0.2),         # config 5's 64-channel layout
])
def test_spade_gradients_on_taken_branch_tol1e3(C, L, r, Hs, B, slope):
    import ag2video_b200.spade as sp
    name = 'gates.%d.%d.%d' % (C, L, r)
    m = sp.SPADE('spadesyncbatch3x3', C, L)
    load_det(m, 91)
    m.fused_slope = slope
    m = m.cuda().train()
    x0 = det_tensor(name + '.x', (B, C, r, r), 3).mul(1.3).add(0.2).cuda()
    seg0 = det_tensor(name + '.seg', (B, L, Hs, Hs), 3).cuda()
    cot = det_tensor(name + '.cot', (B, C, r, r), 3).cuda()

    x, seg = x0.clone().requires_grad_(), seg0.clone().requires_grad_()
    out = m(x, seg)
    saved = out.grad_fn.saved_tensors            # (x, seg, actv, gamma, out, mean, rstd, ...), actv NHWC
    actv = saved[2]
    assert actv.shape == (B, r, r, sp.NHIDDEN)
    gate_a = (actv > 0).permute(0, 3, 1, 2).float()
    gate_out = torch.where(out.detach() > 0, 1.0, slope) if slope != 1.0 else torch.ones_like(out)
    (out * cot).sum().backward()
    got = dict(x=x.grad, seg=seg.grad, **{k: p.grad for k, p in m.named_parameters()})

    p = {k: v.detach().clone().requires_grad_() for k, v in m.named_parameters()}
    xr, segr = x0.clone().requires_grad_(), seg0.clone().requires_grad_()
    ref, z, pre = _forced_reference(xr, segr, p, gate_a, gate_out, slope)
    (ref * cot).sum().backward()
    want = dict(x=xr.grad, seg=segr.grad, **{k: v.grad for k, v in p.items()})

    flip_a = float(((z.detach() > 0).float() != gate_a).float().mean())
    flip_o = float(((pre.detach() > 0) != (out.detach() > 0)).float().mean()) if slope != 1.0 else 0.0
    errs = {k: max_rel(got[k], want[k]) for k in want}
    print('SPADE C=%d L=%d r=%d B=%d: out %.2e | gates differing from the fp32 reference: actv %.3f%%, out %.3f%% | '
          % (C, L, r, B, max_rel(out, ref), 100 * flip_a, 100 * flip_o) + ' '.join('%s %.1e' % kv for kv in errs.items()))
    assert max_rel(out, ref) <= TOL
    assert flip_a <= 5e-3 and flip_o <= 5e-3          # TF32 rounding moves only gates that sit at zero
    assert max(errs.values()) <= TOL, {k: v for k, v in errs.items() if v > TOL}

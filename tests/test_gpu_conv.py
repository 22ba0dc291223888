"""The two implicit-GEMM 3x3 convolution kernels (mma.sync and tcgen05) against
torch's fp32 convolution, every epilogue, including the strided input view that
implements the nearest down-sample and tiles that hang over the image edge."""
import pytest
import torch
import torch.nn.functional as F

from _util import max_rel

pytestmark = pytest.mark.gpu
TOL = 1e-3          # TF32 operands, fp32 accumulation


@pytest.fixture(autouse=True)
def _fp32_reference():
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32 = old


def _run(impl, B, r, rw, Cin, Nout, stride=1, epi='bias_relu', seed=0):
    import ag2video_b200.spade as sp
    g = torch.Generator().manual_seed(seed)
    Hs, Ws = r * stride, rw * stride
    full = torch.randn(B, Cin, Hs, Ws, generator=g).cuda().contiguous(memory_format=torch.channels_last)
    w = (torch.randn(Nout, Cin, 3, 3, generator=g) / (3.0 * Cin ** 0.5)).cuda()
    bias = torch.randn(Nout, generator=g).cuda()
    view = full[:, :, ::stride, ::stride]
    ref = F.conv2d(view.contiguous(), w, bias, padding=1)
    wpk, bpk = sp._pack(w, None, bias, None, False)
    strides = (Hs * Ws * Cin, stride * Ws * Cin, stride * Cin)
    out = torch.full((B, r, rw, Nout), float('nan'), device='cuda')
    ostr = (r * rw * Nout, rw * Nout, Nout)
    old = sp.CONV_IMPL
    sp.CONV_IMPL = impl
    try:
        if epi == 'bias_relu':
            sp._conv(full, strides, B, r, rw, Cin, wpk, bpk, Nout, out, ostr, sp.EPI_BIAS_RELU)
            want = F.relu(ref)
        elif epi == 'bias':
            sp._conv(full, strides, B, r, rw, Cin, wpk, bpk, Nout, out, ostr, sp.EPI_BIAS)
            want = ref
        elif epi == 'gate':
            gate = torch.randn(B, r, rw, Nout, generator=g).cuda()
            sp._conv(full, strides, B, r, rw, Cin, wpk, None, Nout, out, ostr, sp.EPI_GATE, gate=gate)
            want = (ref - bias.view(1, -1, 1, 1)) * (gate.permute(0, 3, 1, 2) > 0)
        elif epi == 'accum':
            init = torch.randn(B, r, rw, Nout, generator=g).cuda()
            out.copy_(init)
            sp._conv(full, strides, B, r, rw, Cin, wpk, None, Nout, out, ostr, sp.EPI_ACCUM)
            want = (ref - bias.view(1, -1, 1, 1)) + init.permute(0, 3, 1, 2)
    finally:
        sp.CONV_IMPL = old
    got = out.permute(0, 3, 1, 2)
    assert torch.isfinite(got).all()
    return max_rel(got, want)


SHAPES = [
    # B, r, rw, Cin, Nout, stride
    (2, 16, 16, 64, 128, 1),        # 8 rows x 16 cols tiles
    (2, 8, 8, 128, 256, 4),         # whole batch in one tile, strided view
    (1, 32, 32, 32, 128, 2),
    (1, 4, 128, 64, 128, 1),        # full-width row tiles
    (3, 8, 8, 32, 160, 1),          # batch tile hangs over (3 images, 2 per tile); Nout not a multiple of 128
    (1, 16, 256, 32, 64, 1),
]


@pytest.mark.parametrize('impl', [1, 2])
@pytest.mark.parametrize('shape', SHAPES)
def test_conv_bias_relu(shape, impl):
    assert _run(impl, *shape, epi='bias_relu') <= TOL


@pytest.mark.parametrize('impl', [1, 2])
@pytest.mark.parametrize('epi', ['bias', 'gate', 'accum'])
def test_conv_epilogues(epi, impl):
    assert _run(impl, 2, 16, 16, 64, 128, 1, epi=epi) <= TOL
    assert _run(impl, 2, 8, 8, 128, 256, 2, epi=epi) <= TOL


def test_tc_and_mma_agree_closely():
    """Operands pre-rounded to TF32 (tcgen05 truncates, mma.sync rounds in registers):
    identical products in both kernels, differences are accumulation order only."""
    import ag2video_b200.spade as sp
    g = torch.Generator().manual_seed(5)
    B, r, Cin, Nout = 2, 32, 128, 128
    x = torch.randn(B, Cin, r, r, generator=g).cuda().contiguous(memory_format=torch.channels_last)
    w = (torch.randn(Nout, Cin, 3, 3, generator=g) / 30).cuda()
    wpk, _ = sp._pack(w, None, None, None, False)
    x = sp._seg_operand(x)          # rounded to TF32 like every GEMM operand of the path
    outs = []
    old = sp.CONV_IMPL
    try:
        for impl in (1, 2):
            sp.CONV_IMPL = impl
            out = torch.empty(B, r, r, Nout, device='cuda')
            sp._conv(x, (r * r * Cin, r * Cin, Cin), B, r, r, Cin, wpk, None, Nout, out, (r * r * Nout, r * Nout, Nout), sp.EPI_BIAS)
            outs.append(out)
    finally:
        sp.CONV_IMPL = old
    assert max_rel(outs[0], outs[1]) <= 1e-5

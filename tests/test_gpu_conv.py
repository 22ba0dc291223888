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
    (2, 128, 128, 32, 128, 1),      # 256 M tiles: persistent CTAs take 2 work items (both TMEM accumulator stages)
    (1, 128, 256, 64, 256, 1),      # 256-column work items, 3+ per CTA
    (2, 64, 64, 32, 512, 2),        # 64 M tiles x 2 column tiles, strided view
]


@pytest.mark.parametrize('impl', [1, 2])
@pytest.mark.parametrize('shape', SHAPES)
def test_conv_bias_relu(shape, impl):
    assert _run(impl, *shape, epi='bias_relu') <= TOL


@pytest.mark.parametrize('tile', ['1x128', '2x128', '1x256', '2x256'])
@pytest.mark.parametrize('epi', ['bias_relu', 'gate', 'accum'])
def test_tc_tile_shapes(epi, tile, monkeypatch):
    """Every CTA tile shape of the tcgen05 kernel (1 or 2 M tiles x 128 or 256 columns),
    including an odd number of M tiles and Nout that does not fill the last column tile."""
    monkeypatch.setenv('AG2V_TC_TILE', tile)
    assert _run(2, 3, 8, 8, 64, 288, 1, epi=epi) <= TOL          # 2 M tiles (3 images, 2 per tile), Nout = 256 + 32
    assert _run(2, 1, 48, 128, 32, 128, 1, epi=epi) <= TOL        # 48 row tiles
    assert _run(2, 2, 16, 16, 128, 512, 2, epi=epi) <= TOL        # strided view, 4 M tiles


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


@pytest.mark.parametrize('impl', [1, 2])
@pytest.mark.parametrize('B,r,rw,Cin,Nout,stride,two', [
    (2, 16, 16, 64, 128, 1, False),      # 2 x 16 pixel chunks
    (2, 8, 8, 128, 256, 2, True),        # gamma|beta pair, strided input view, batch-spanning chunk
    (1, 64, 64, 32, 32, 1, False),       # many chunks, split-K
    (2, 32, 32, 512, 128, 4, False),     # the shared conv at real widths (label_nc 512), stride 4
])
def test_wgrad_kernels(B, r, rw, Cin, Nout, stride, two, impl):
    """dW of a 3x3 conv from both weight-gradient kernels vs torch autograd (fp32)."""
    import ag2video_b200.spade as sp
    g = torch.Generator().manual_seed(3)
    Hs, Ws = r * stride, rw * stride
    full = sp._seg_operand(torch.randn(B, Cin, Hs, Ws, generator=g).cuda())
    dy = sp._seg_operand(torch.randn(B, Nout, r, rw, generator=g).cuda())        # NHWC, TF32-rounded like the path
    view = full[:, :, ::stride, ::stride].contiguous()
    w = torch.zeros(Nout, Cin, 3, 3, device='cuda', requires_grad=True)
    F.conv2d(view, w, padding=1).backward(dy)
    dy_rows = dy.permute(0, 2, 3, 1).reshape(B * r * rw, Nout)
    if two:      # columns of dy in the gamma|beta interleaved order, two OIHW outputs
        C = Nout // 2
        cols = torch.tensor([16 * (c // 8) + 8 * isb + c % 8 for isb in (0, 1) for c in range(C)], device='cuda')
        dy_pk = torch.empty_like(dy_rows)
        dy_pk[:, cols] = dy_rows
        like_a = torch.empty(C, Cin, 3, 3, device='cuda')
        old = sp.CONV_IMPL
        sp.CONV_IMPL = impl
        try:
            da, db = sp._wgrad(dy_pk.contiguous(), Nout, full, (Hs * Ws * Cin, stride * Ws * Cin, stride * Cin), Cin, B, r, rw,
                               True, like_a, like_a)
        finally:
            sp.CONV_IMPL = old
        got = torch.cat([da, db], 0)
    else:
        like_a = torch.empty(Nout, Cin, 3, 3, device='cuda')
        old = sp.CONV_IMPL
        sp.CONV_IMPL = impl
        try:
            got, _ = sp._wgrad(dy_rows.contiguous(), Nout, full, (Hs * Ws * Cin, stride * Ws * Cin, stride * Cin), Cin, B, r, rw,
                               False, like_a, None)
        finally:
            sp.CONV_IMPL = old
    assert max_rel(got, w.grad) <= TOL


def test_plain_conv3x3_fc_shape_forward_and_gradients_tol1e3():
    """The generator's `fc` (spade_generator.py:16,56: 512 -> 1024 at 8x8, 6 images) on the tcgen05 implicit-GEMM
    kernel: output, input gradient, weight gradient and bias gradient against torch's fp32 convolution."""
    import ag2video_b200.spade as sp
    g = torch.Generator().manual_seed(3)
    conv = torch.nn.Conv2d(512, 1024, 3, padding=1).cuda().to(memory_format=torch.channels_last)
    x0 = torch.randn(6, 512, 8, 8, generator=g).cuda().contiguous(memory_format=torch.channels_last)
    x0 = x0.view(torch.int32).bitwise_and(-8192).view(torch.float32)         # TF32-representable operand, as SharedSeg hands it over
    cot = torch.randn(6, 1024, 8, 8, generator=g).cuda()
    x = x0.clone().requires_grad_()
    assert sp.plain_conv3x3_usable(conv, x)
    y = sp.plain_conv3x3(conv, x)
    (y * cot).sum().backward()
    got = [y.detach(), x.grad.clone(), conv.weight.grad.clone(), conv.bias.grad.clone()]
    conv.zero_grad()
    x2 = x0.clone().requires_grad_()
    ref = F.conv2d(x2, conv.weight, conv.bias, padding=1)
    (ref * cot).sum().backward()
    want = [ref.detach(), x2.grad, conv.weight.grad, conv.bias.grad]
    errs = {n: max_rel(a, b) for n, a, b in zip(('y', 'dx', 'dw', 'db'), got, want)}
    print('plain_conv3x3 512->1024 @8x8: ' + ' '.join('%s %.1e' % kv for kv in errs.items()))
    assert max(errs.values()) <= TOL, errs


@pytest.mark.parametrize('B,r,Cin,Nout,groups,with_res', [(6, 256, 128, 64, 3, True), (6, 128, 128, 256, 3, False),
                                                         (6, 64, 256, 256, 3, True), (2, 256, 128, 64, 1, False),
                                                         (4, 32, 512, 512, 2, True)])
def test_conv_epilogue_bn_statistics_tol1e5(B, r, Cin, Nout, groups, with_res):
    """BN statistics fused into the tcgen05 convolution epilogue (warp-shuffle column sums per M tile + fixed-order
    reduction): per-group sum / sum of squares of the OUTPUT against torch on the same output (1e-5), the output itself
    bit-identical to the plain launch, and bitwise reproducible."""
    import ctypes
    import ag2video_b200.spade as sp
    from ag2video_b200 import _lib as L
    lib = L.lib()
    mt, tpg = ctypes.c_int(0), ctypes.c_int(0)
    if not lib.ag2v_conv3x3_stats_info(B, r, r, Cin, Nout, groups, ctypes.byref(mt), ctypes.byref(tpg)):
        pytest.skip('this shape runs split-K: statistics stay in their own kernel')
    g = torch.Generator().manual_seed(B * r + Cin)
    x = torch.randn(B, Cin, r, r, generator=g).cuda().contiguous(memory_format=torch.channels_last)
    w = (torch.randn(Nout, Cin, 3, 3, generator=g) / (3.0 * Cin ** 0.5)).cuda()
    bias = torch.randn(Nout, generator=g).cuda()
    scale = (torch.rand(groups, generator=g) + 0.5).cuda()
    res = torch.randn(B, Nout, r, r, generator=g).cuda().contiguous(memory_format=torch.channels_last) if with_res else None
    wpk, _ = sp._pack(w, None, None, None, False)
    Pg = B * r * r // groups
    outs = []
    for _ in range(2):
        out = torch.full((B, Nout, r, r), float('nan'), device='cuda').contiguous(memory_format=torch.channels_last)
        part = torch.empty(mt.value * 2 * Nout, device='cuda')
        sums = torch.empty(groups * 2 * Nout, device='cuda', dtype=torch.float64)
        L.check(lib.ag2v_conv3x3_bias_stats(L.ptr(x), B, r, r, Cin, L.ptr(wpk), L.ptr(bias), Nout, L.ptr(out), 0,
                                            Pg if groups > 1 else 0, L.ptr(scale), L.ptr(res), groups, L.ptr(part), L.ptr(sums), L.stream()))
        outs.append((out, sums))
    (out, sums), (out2, sums2) = outs
    assert torch.equal(out, out2) and torch.equal(sums, sums2)
    plain = torch.empty_like(out, memory_format=torch.channels_last)
    sp._conv(x, (r * r * Cin, r * Cin, Cin), B, r, r, Cin, wpk, bias, Nout, plain, (r * r * Nout, r * Nout, Nout), sp.EPI_BIAS,
             group_pixels=Pg if groups > 1 else 0, scale=scale, res=res)
    assert torch.equal(out, plain)
    og = out.double().view(groups, B // groups, Nout, r, r)
    want = torch.stack([og.sum(dim=(1, 3, 4)), (og * og).sum(dim=(1, 3, 4))], dim=1).reshape(-1)
    err = max_rel(sums, want)
    print('fused BN statistics B=%d r=%d %d->%d groups=%d: %.1e' % (B, r, Cin, Nout, groups, err))
    assert err <= 1e-5

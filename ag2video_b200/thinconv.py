"""3x3 convolutions with 2-3 output channels at full resolution (csrc/k6_thinconv.cu): the
generator's ``conv_img`` (spade_models/networks/generator.py: ``tanh(conv_img(leaky_relu(x, 0.2)))``)
and the flow head ``conv_flow`` (flows_generator.py).  ``thin_conv3x3(conv, x, slope_in, act_out)``
evaluates ``act_out(conv(leaky_relu(x, slope_in)))`` for a plain ``nn.Conv2d`` in one streaming
kernel per direction; shapes the kernels are not instantiated for raise (no silent library path)."""
import torch
import torch.nn.functional as F

from . import _lib as L

c_p, c_i, c_f, c_sz = L.c_p, L.c_i, L.c_f, L.c_sz
L.register('ag2v_thin_conv3x3_supported', c_i, [c_i, c_i])
L.register('ag2v_thin_conv3x3_workspace_floats', c_sz, [c_i] * 5)
L.register('ag2v_thin_conv3x3_fwd', c_i, [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_i, c_p, c_p])
L.register('ag2v_thin_conv3x3_bwd', c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_i, c_p, c_p, c_p, c_p, c_p])

_ACT = {None: 0, 'tanh': 1}
CL = torch.channels_last


def _w_layout(w):
    """(tensor to hand to the kernel, channels_last flag)."""
    if w.is_contiguous():
        return w, 0
    if w.is_contiguous(memory_format=CL):
        return w, 1
    return w.contiguous(), 0


class _ThinConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, slope_in, act_out):
        L.need_cuda(x, weight)
        x = x.float().contiguous(memory_format=CL)
        B, CI, H, W = x.shape
        CO = weight.shape[0]
        w, w_cl = _w_layout(weight.float())
        y = torch.empty(B, CO, H, W, device=x.device, dtype=torch.float32, memory_format=CL)
        L.check(L.lib().ag2v_thin_conv3x3_fwd(L.ptr(x), L.ptr(w), L.ptr(bias), B, H, W, CI, CO, w_cl, float(slope_in), act_out,
                                              L.ptr(y), L.stream()))
        ctx.save_for_backward(x, weight, y)
        ctx.meta = (slope_in, act_out, bias is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y = ctx.saved_tensors
        slope_in, act_out, has_bias = ctx.meta
        B, CI, H, W = x.shape
        CO = weight.shape[0]
        lib = L.lib()
        dy = dy.float().contiguous(memory_format=CL)
        w, w_cl = _w_layout(weight.float())
        ws = torch.empty(lib.ag2v_thin_conv3x3_workspace_floats(B, H, W, CI, CO), device=x.device, dtype=torch.float32)
        dx = torch.empty_like(x, memory_format=CL) if ctx.needs_input_grad[0] else None
        dw = torch.empty_like(w)                      # same storage order as the weight the kernel read
        db = torch.empty(CO, device=x.device, dtype=torch.float32)
        L.check(lib.ag2v_thin_conv3x3_bwd(L.ptr(x), L.ptr(w), L.ptr(dy), L.ptr(y), B, H, W, CI, CO, w_cl, float(slope_in), act_out,
                                          L.ptr(dx), L.ptr(ws), L.ptr(dw), L.ptr(db), L.stream()))
        return dx, dw, (db if has_bias else None), None, None


def thin_conv3x3(conv, x, slope_in=1.0, act_out=None, allow_library=False):
    """``act_out(conv(leaky_relu(x, slope_in)))`` for a 3x3, stride-1, padding-1 ``nn.Conv2d``.
    The kernels are instantiated for (Cin, Cout) = (64, 3) and (32, 2) - ``conv_img`` at ngf = 64 and
    ``conv_flow`` at nff = 32, the reference defaults.  Any other shape RAISES: this package has no silent
    library path.  ``allow_library=True`` is the explicit opt-in to evaluate such a shape with the module
    itself (cuDNN); it warns once per shape."""
    w = conv.weight
    usable = (x.is_cuda and tuple(w.shape[2:]) == (3, 3) and conv.stride == (1, 1) and conv.padding == (1, 1)
              and conv.dilation == (1, 1) and conv.groups == 1
              and L.lib().ag2v_thin_conv3x3_supported(w.shape[1], w.shape[0]))
    if not usable:
        what = 'thin_conv3x3: no sm_100a kernel for a %s convolution %d -> %d on a %s tensor' % (
            'x'.join(str(k) for k in w.shape[2:]), w.shape[1], w.shape[0], x.device.type)
        if not allow_library:
            raise NotImplementedError(what + ' (instantiated: 64 -> 3 and 32 -> 2, 3x3 / stride 1 / padding 1); '
                                      'pass allow_library=True to evaluate it with cuDNN instead')
        key = (tuple(w.shape), x.device.type)
        if key not in _WARNED:
            _WARNED.add(key)
            import warnings
            warnings.warn(what + ': evaluated by the library (cuDNN) because allow_library=True', RuntimeWarning)
        z = conv(F.leaky_relu(x, slope_in) if slope_in != 1.0 else x)
        return torch.tanh(z) if act_out == 'tanh' else z
    return _ThinConvFn.apply(x, w, conv.bias, float(slope_in), _ACT[act_out])


_WARNED = set()

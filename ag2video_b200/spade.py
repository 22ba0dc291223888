"""SPADE and SPADEResnetBlock on sm_100a (drop-in for
models/spade_models/networks/normalization.py:66-110 and architecture.py:21-68).

Same constructors, ``forward`` signatures and state-dict keys as the reference
(``param_free_norm.{running_mean,running_var,num_batches_tracked}``,
``mlp_shared.0.*``, ``mlp_gamma.*``, ``mlp_beta.*``; ``conv_0/1/s`` keep torch's
spectral-norm hooks).  The modulation path runs on this library's kernels:

  stats kernel -> [SyncBN all-reduce] -> implicit-GEMM conv (seg -> actv, ReLU)
  -> implicit-GEMM conv (actv -> gamma|beta) whose epilogue applies
  ``(x-mean)*rstd*(1+gamma)+beta`` (+ LeakyReLU inside SPADEResnetBlock).

Activations are handled as NHWC (``torch.channels_last``); the nearest
down-sample of the segmap is a strided view, never materialised.
"""
import re

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils import spectral_norm

from . import _lib as L
from ._lib import c_f, c_i, c_p, c_sz
from .specnorm import SpectralNormGroup, conv_scaled

c_ll = L.ctypes.c_longlong
c_d = L.ctypes.c_double

L.register('ag2v_chan_partial_floats', c_sz, [c_ll, c_i, c_i])
L.register('ag2v_bn_stats', c_i, [c_p, c_ll, c_i, c_i, c_p, c_p, c_p])
L.register('ag2v_bn_finalize', c_i, [c_p, c_d, c_d, c_i, c_i, c_f, c_f, c_p, c_p, c_p, c_p, c_p, c_p])
L.register('ag2v_bn_act_fwd', c_i, [c_p] * 5 + [c_ll, c_i, c_i, c_f, c_p, c_p])
L.register('ag2v_bn_eval_stats', c_i, [c_p, c_p, c_i, c_i, c_f, c_p, c_p, c_p, c_p])
L.register('ag2v_spade_bwd_pre', c_i, [c_p] * 6 + [c_ll, c_i, c_i, c_i, c_f, c_i, c_i, c_i, c_i] + [c_p] * 4 + [c_p])
L.register('ag2v_spade_bwd_dx', c_i, [c_p] * 5 + [c_d, c_i, c_ll, c_i, c_i, c_i, c_i, c_p, c_p])
L.register('ag2v_pack_w3x3', c_i, [c_p] * 4 + [c_i, c_i, c_i, c_i, c_p, c_p, c_p])
L.register('ag2v_unpack_dw3x3', c_i, [c_p, c_i, c_i, c_i, c_i, c_p, c_p, c_p])
L.register('ag2v_double_to_float', c_i, [c_p, c_i, c_i, c_ll, c_p, c_p])
L.register('ag2v_round_tf32', c_i, [c_p, c_p, c_ll, c_p])
L.register('ag2v_pack_w3x3_cl', c_i, [c_p] * 4 + [c_i, c_i, c_i, c_i, c_p, c_p, c_p])
L.register('ag2v_unpack_dw3x3_cl', c_i, [c_p, c_i, c_i, c_i, c_i, c_p, c_p, c_p])
L.register('ag2v_scaled_grad_pre', c_i, [c_p, c_p, c_ll, c_i, c_i, c_p, c_p, c_p, c_p])
L.register('ag2v_wgrad3x3_nsplit', c_i, [c_i] * 6)
L.register('ag2v_wgrad3x3', c_i, [c_p, c_i, c_p, c_ll, c_ll, c_ll, c_i, c_i, c_i, c_i, c_p, c_i, c_p])
L.register('ag2v_conv3x3', c_i, [c_p, c_ll, c_ll, c_ll, c_i, c_i, c_i, c_i, c_p, c_p, c_i, c_p, c_ll, c_ll, c_ll,
                                 c_i, c_i, c_p, c_p, c_p, c_p, c_f, c_i, c_ll, c_i, c_p, c_p, c_p, c_p, c_sz, c_i, c_p])
L.register('ag2v_conv3x3_splitk_floats', c_sz, [c_i] * 5)
L.register('ag2v_conv3x3_tc_supported', c_i, [c_i] * 6)
L.register('ag2v_conv3x3_stats_info', c_i, [c_i] * 6 + [c_p, c_p])
L.register('ag2v_conv3x3_bias_stats', c_i, [c_p, c_i, c_i, c_i, c_i, c_p, c_p, c_i, c_p, c_i, c_ll, c_p, c_p, c_i, c_p, c_p, c_p])

EPI_BIAS, EPI_BIAS_RELU, EPI_SPADE, EPI_GATE, EPI_ACCUM = 0, 1, 2, 3, 4
NHIDDEN = 128                      # "Yes, hardcoded" (normalization.py:84)
FUSE_STATS = True                  # BN statistics of a block convolution's output come out of its epilogue (tests flip this)
CONV_IMPL = 0                      # 0 auto, 1 mma.sync, 2 tcgen05, 3 mma.sync 3xTF32 (validation; tests flip this)


def _precise():
    return CONV_IMPL == 3

_sync_group = {'group': None, 'enabled': False}


def set_sync_bn(enabled, group=None):
    """SyncBN semantics across data-parallel ranks (reference:
    sync_batchnorm/batchnorm.py:74-83): all-reduce the per-channel sums."""
    _sync_group['enabled'] = bool(enabled)
    _sync_group['group'] = group


def _world():
    import torch.distributed as dist
    if _sync_group['enabled'] and dist.is_available() and dist.is_initialized():
        return dist, dist.get_world_size(_sync_group['group'])
    return None, 1


def _allreduce_sums(sums, dist, overlap=False):
    """Sum the per-channel double sums over the data-parallel ranks (sync_batchnorm/batchnorm.py:74-83, :105-145), in
    place.  One kernel of this library over NVLink peer memory (csrc/k8_peer.cu) where the ranks share a node, the
    group's own all-reduce otherwise.  overlap=True returns an object whose .wait() orders the current stream after the
    exchange, so that work which does not need the sums can be enqueued in between."""
    from . import peer
    group = _sync_group['group']
    ex = peer.get(group)
    if ex is not None and sums.numel() <= ex.cap:
        if overlap:
            return ex.allreduce_async(sums)
        ex.allreduce(sums)
        return None
    if overlap:
        return dist.all_reduce(sums, group=group, async_op=True)
    dist.all_reduce(sums, group=group)
    return None


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last)


def _seg_operand(seg):
    """NHWC copy of the segmap rounded to TF32 (it is the A operand of the shared
    convolution and of its weight gradient; tcgen05's TF32 path truncates)."""
    src = _cl(seg.detach().float())
    if _precise():
        return src
    if src.numel() % 4:
        raise NotImplementedError('segmap element count must be a multiple of 4')
    dst = torch.empty_like(src, memory_format=torch.channels_last)
    L.check(L.lib().ag2v_round_tf32(L.ptr(src), L.ptr(dst), src.numel(), L.stream()))
    return dst


_Timed = L.timed                   # per-launch CUDA-event timing, active while _lib.PROFILE is a list


def _conv(inp, in_strides, B, Hh, Ww, Cin, wpk, bias, Nout, out, out_strides, epi, round_out=0,
          x=None, mean=None, rstd=None, gamma_out=None, slope=1.0, C=0, gate=None, group_pixels=0, scale=None, res=None,
          x_up=0):
    lib = L.lib()
    nws = lib.ag2v_conv3x3_splitk_floats(B, Hh, Ww, Cin, Nout) if CONV_IMPL in (0, 2) else 0
    ws = torch.empty(nws, device=out.device, dtype=torch.float32) if nws else None
    with _Timed('conv3x3', 2.0 * 9 * B * Hh * Ww * Cin * Nout, ('epi%d' % epi, Hh, Cin, Nout)):
        L.check(lib.ag2v_conv3x3(L.ptr(inp), in_strides[0], in_strides[1], in_strides[2], B, Hh, Ww, Cin,
                                 L.ptr(wpk), L.ptr(bias), Nout, L.ptr(out), out_strides[0], out_strides[1],
                                 out_strides[2], epi, round_out, L.ptr(x), L.ptr(mean), L.ptr(rstd),
                                 L.ptr(gamma_out), float(slope), C, group_pixels, int(x_up), L.ptr(scale), L.ptr(res), L.ptr(gate),
                                 L.ptr(ws), nws, CONV_IMPL, L.stream()))


def _is_cl(w):
    """True when a 4-D weight is stored channels_last ([Co][kh][kw][Ci]) and not also contiguous."""
    return w.dim() == 4 and not w.is_contiguous() and w.is_contiguous(memory_format=torch.channels_last) and w.shape[1] % 4 == 0


def _pack(wa, wb, ba, bb, dgrad):
    """GEMM-layout copy of one (or two interleaved) 3x3 weight(s), straight from the storage order
    the module keeps them in (OIHW or channels_last)."""
    Co, Ci = wa.shape[0], wa.shape[1]
    Ntot = 2 * Co if wb is not None else Co
    dst = torch.empty(9 * Ntot * Ci, device=wa.device, dtype=torch.float32)
    bias = torch.empty(Ntot, device=wa.device, dtype=torch.float32) if not dgrad else None
    if _is_cl(wa) and (wb is None or _is_cl(wb)):
        L.check(L.lib().ag2v_pack_w3x3_cl(L.ptr(wa), L.ptr(wb), L.ptr(ba), L.ptr(bb), Co, Ci, int(dgrad), int(not _precise()),
                                          L.ptr(dst), L.ptr(bias), L.stream()))
        return dst, bias
    wa = wa.contiguous()
    wb = wb.contiguous() if wb is not None else None
    L.check(L.lib().ag2v_pack_w3x3(L.ptr(wa), L.ptr(wb), L.ptr(ba), L.ptr(bb), Co, Ci, int(dgrad), int(not _precise()),
                                   L.ptr(dst), L.ptr(bias), L.stream()))
    return dst, bias


def _wgrad(dy, Nout, x, x_strides, Cin, B, Hh, Ww, two, like_a, like_b):
    lib = L.lib()
    nsplit = lib.ag2v_wgrad3x3_nsplit(B, Hh, Ww, Nout, Cin, CONV_IMPL)
    part = torch.empty(nsplit * 9 * Nout * Cin, device=dy.device, dtype=torch.float32)
    with _Timed('wgrad3x3', 2.0 * 9 * B * Hh * Ww * Cin * Nout, ('wgrad', Hh, Cin, Nout)):
        L.check(lib.ag2v_wgrad3x3(L.ptr(dy), Nout, L.ptr(x), x_strides[0], x_strides[1], x_strides[2], Cin, B, Hh, Ww,
                                  L.ptr(part), CONV_IMPL, L.stream()))
    Co = Nout // 2 if two else Nout
    if _is_cl(like_a) and (not two or _is_cl(like_b)):        # gradients in the parameters' own storage order
        dwa = torch.empty_like(like_a, memory_format=torch.channels_last)
        dwb = torch.empty_like(like_b, memory_format=torch.channels_last) if two else None
        L.check(lib.ag2v_unpack_dw3x3_cl(L.ptr(part), nsplit, Co, Cin, int(two), L.ptr(dwa), L.ptr(dwb), L.stream()))
        return dwa, dwb
    dwa = torch.empty_like(like_a, memory_format=torch.contiguous_format)
    dwb = torch.empty_like(like_b, memory_format=torch.contiguous_format) if two else None
    L.check(lib.ag2v_unpack_dw3x3(L.ptr(part), nsplit, Co, Cin, int(two), L.ptr(dwa), L.ptr(dwb), L.stream()))
    return dwa, dwb


class SharedSeg:
    """One segmap shared by many SPADE layers of a generator call (18 in the
    reference's SPADEGenerator): converted to NHWC once, and all layers add their
    input gradient into ONE buffer instead of materialising 18 full-resolution
    gradients.  Create with ``SharedSeg.wrap(seg)`` and pass it as ``segmap``."""

    def __init__(self):
        self.seg = None        # NHWC data, detached
        self.token = None      # autograd handle tying the consumers to the producer
        self.grad = None

    @staticmethod
    def wrap(seg):
        L.need_cuda(seg)
        h = SharedSeg()
        seg_cl, token = _SharedSegFn.apply(seg, h)
        h.seg = seg_cl.detach()
        h.token = token
        return h

    def grad_buffer(self):
        if self.grad is None:
            self.grad = torch.zeros_like(self.seg, memory_format=torch.channels_last)
        return self.grad

    def grad_buffer_for_full_cover(self):
        """(buffer, fresh): for a contribution that covers EVERY element of the segmap.  The first such contribution
        of a backward pass (the full-resolution SPADE layers run their backward first) can be written instead of
        added, so the buffer is neither zero-filled nor read: fresh = True and the caller must overwrite all of it."""
        if self.grad is None:
            self.grad = torch.empty_like(self.seg, memory_format=torch.channels_last)
            return self.grad, True
        return self.grad, False

    def nearest(self, h, w):
        """``F.interpolate(seg, size=(h, w))`` (nearest, integer ratio) read from the shared copy;
        its gradient goes straight into the shared buffer instead of a full-resolution tensor of
        zeros (the generator's ``fc`` input, spade_models/networks/generator.py)."""
        H, W = self.seg.shape[2:]
        if H % h or W % w:
            raise NotImplementedError('nearest resize %dx%d -> %dx%d is not an integer ratio' % (H, W, h, w))
        return _SegNearestFn.apply(self.token, self, H // h, W // w)


class _SharedSegFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, seg, handle):
        ctx.handle = handle
        seg_cl = _seg_operand(seg)
        token = torch.zeros(1, device=seg.device)
        ctx.mark_non_differentiable(seg_cl)
        return seg_cl, token

    @staticmethod
    def backward(ctx, _dseg, _dtoken):
        g = ctx.handle.grad
        ctx.handle.grad = None
        return g, None


class _SegNearestFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, token, handle, sy, sx):
        ctx.handle, ctx.step = handle, (sy, sx)
        return handle.seg[:, :, ::sy, ::sx].contiguous(memory_format=torch.channels_last)

    @staticmethod
    def backward(ctx, g):
        sy, sx = ctx.step
        ctx.handle.grad_buffer()[:, :, ::sy, ::sx] += g
        return torch.zeros(1, device=g.device), None, None, None


class _SpadeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, seg, token, w_sh, b_sh, w_g, b_g, w_b, b_b, mod, shared, slope, groups=1, next_scale=None,
                round_out=False, upsample=False, pre_sums=None):
        L.need_cuda(x, seg, w_sh)
        lib = L.lib()
        dev = x.device
        x = _cl(x.float())
        seg = seg if shared is not None else _seg_operand(seg)
        B, C, r, rw = x.shape
        if upsample:                               # x is the low-resolution source of a nearest 2x up-sampling
            r, rw = 2 * r, 2 * rw
        _, Lc, Hs, Ws = seg.shape
        if Hs % r or Ws % rw:
            raise NotImplementedError('SPADE: segmap %dx%d is not an integer multiple of the activation %dx%d' % (Hs, Ws, r, rw))
        if C % 8 or Lc % 4:
            raise NotImplementedError('SPADE kernels need norm_nc %% 8 == 0 and label_nc %% 4 == 0 (got %d, %d)' % (C, Lc))
        sy, sx = Hs // r, Ws // rw
        seg_strides = (Hs * Ws * Lc, sy * Ws * Lc, sx * Lc)
        P = B * r * rw
        training = mod.training
        bn = mod.param_free_norm
        if B % groups:
            raise ValueError('SPADE: batch %d is not a multiple of groups=%d' % (B, groups))
        G = groups                                 # statistics per group of B/G images (one reference call each)
        Pg = P // G
        mean = torch.empty(G * C, device=dev, dtype=torch.float32)
        rstd = torch.empty(G * C, device=dev, dtype=torch.float32)
        count = float(Pg)
        if training:
            Ps = Pg // 4 if upsample else Pg       # statistics of up(x) are those of x (every element appears 4 times)
            dist, world = _world()
            if pre_sums is not None and pre_sums.numel() == G * 2 * C:
                # the convolution that produced x left its per-group sums behind (fused into its epilogue); the
                # all-reduce below works in place and two layers may normalise the same x, so take a private copy
                sums = pre_sums.clone() if world > 1 else pre_sums
            else:
                part = torch.empty(G * lib.ag2v_chan_partial_floats(Ps, C, 2), device=dev, dtype=torch.float32)
                sums = torch.empty(G * 2 * C, device=dev, dtype=torch.float64)
                with _Timed('k3_bn_stats', 4.0 * G * Ps * C, (r, C)):
                    L.check(lib.ag2v_bn_stats(L.ptr(x), Ps, C, G, L.ptr(part), L.ptr(sums), L.stream()))
            pending = None
            if world > 1:
                # SyncBN: the sums travel while the shared convolution (which does not need them) runs
                pending = _allreduce_sums(sums, dist, overlap=True)
            count = float(Pg * world)
        _cache_of(mod).forward_begin()
        pk = mod._packed(w_sh, b_sh, w_g, b_g, w_b, b_b)
        actv = torch.empty(B, r, rw, NHIDDEN, device=dev, dtype=torch.float32)
        _conv(seg, seg_strides, B, r, rw, Lc, pk['w1'], pk['b1'], NHIDDEN, actv, (r * rw * NHIDDEN, rw * NHIDDEN, NHIDDEN),
              EPI_BIAS_RELU, round_out=1)
        if training:
            if pending is not None:
                pending.wait()                     # stream-level wait, the host does not block
            L.check(lib.ag2v_bn_finalize(L.ptr(sums), float(Ps * world), count, C, G, bn.eps, bn.momentum, None,
                                         L.ptr(bn.running_mean), L.ptr(bn.running_var), L.ptr(mean), L.ptr(rstd), L.stream()))
        else:                                      # same running estimates for every group
            L.check(lib.ag2v_bn_eval_stats(L.ptr(bn.running_mean), L.ptr(bn.running_var), C, G, bn.eps, None, L.ptr(mean),
                                           L.ptr(rstd), L.stream()))
        out = torch.empty(B, C, r, rw, device=dev, dtype=torch.float32, memory_format=torch.channels_last)
        need_grad = any(ctx.needs_input_grad)
        gamma = torch.empty(B, r, rw, C, device=dev, dtype=torch.float32) if need_grad else None
        _conv(actv, (r * rw * NHIDDEN, rw * NHIDDEN, NHIDDEN), B, r, rw, NHIDDEN, pk['w2'], pk['b2'], 2 * C, out,
              (r * rw * C, rw * C, C), EPI_SPADE, round_out=int(round_out and not _precise()), x=x, mean=mean, rstd=rstd,
              gamma_out=gamma, slope=slope, C=C, group_pixels=Pg if G > 1 else 0, x_up=int(upsample))
        if need_grad:
            ctx.save_for_backward(x, seg, actv, gamma, out, mean, rstd, w_sh, w_g, w_b, next_scale)
        ctx.meta = (B, C, r, rw, Lc, Hs, Ws, seg_strides, P, count, training, slope, G, bool(upsample))
        ctx.mod, ctx.shared = mod, shared
        return out

    @staticmethod
    def backward(ctx, dout):
        x, seg, actv, gamma, out, mean, rstd, w_sh, w_g, w_b, next_scale = ctx.saved_tensors
        B, C, r, rw, Lc, Hs, Ws, seg_strides, P, count, training, slope, G, upsample = ctx.meta
        Pg = P // G
        uh, uw = (r, rw) if upsample else (0, 0)
        mod, shared = ctx.mod, ctx.shared
        _cache_of(mod).backward_seen()
        lib = L.lib()
        dev = x.device
        dout = _cl(dout.float())
        act = 0 if slope == 1.0 else 1
        dgb = torch.empty(P, 2 * C, device=dev, dtype=torch.float32)
        dx = torch.empty(B, C, r, rw, device=dev, dtype=torch.float32, memory_format=torch.channels_last)
        part = torch.empty(G * lib.ag2v_chan_partial_floats(Pg, C, 5), device=dev, dtype=torch.float32)
        sums = torch.empty(G * 5 * C, device=dev, dtype=torch.float64)
        with _Timed('k3_spade_bwd_pre', 4.0 * P * C * 7, (r, C)):       # reads dout, out, x, gamma; writes dgamma|dbeta, dxhat
            L.check(lib.ag2v_spade_bwd_pre(L.ptr(dout), L.ptr(out), L.ptr(x), L.ptr(gamma), L.ptr(mean), L.ptr(rstd), Pg, C, G,
                                           act, float(slope), int(not _precise()), 0, uh, uw, L.ptr(dgb), L.ptr(dx), L.ptr(part),
                                           L.ptr(sums), L.stream()))
        db = torch.empty(2 * C, device=dev, dtype=torch.float32)     # [sum g | sum g*xhat] = [d bias_beta | d bias_gamma]
        L.check(lib.ag2v_double_to_float(L.ptr(sums), 2 * C, G, 5 * C, L.ptr(db), L.stream()))
        dscale = None
        if next_scale is not None and ctx.needs_input_grad[13]:
            # y = conv(out) * scale_g: d scale_g = <dy_g, conv(out)_g> = <dout_g, out_g> / scale_g  (adjoint of the conv)
            dscale = (sums.view(G, 5, C)[:, 4].sum(dim=1) / next_scale.double()).float()
        pending = None
        if training:
            dist, world = _world()
            if world > 1:              # db (local sums) is already extracted; the BN backward needs global sums:
                # they travel while the gamma / beta weight and input gradients (which do not need them) run
                pending = _allreduce_sums(sums, dist, overlap=True)
        a_strides = (r * rw * NHIDDEN, rw * NHIDDEN, NHIDDEN)
        # gamma / beta convolutions: weight gradient, then input gradient gated by the ReLU of actv
        dw_g, dw_b = _wgrad(dgb, 2 * C, actv, a_strides, NHIDDEN, B, r, rw, True, w_g, w_b)
        pkt = mod._packed_t(w_sh, w_g, w_b)
        dactv = torch.empty(B, r, rw, NHIDDEN, device=dev, dtype=torch.float32)
        _conv(dgb, (r * rw * 2 * C, rw * 2 * C, 2 * C), B, r, rw, 2 * C, pkt['w2t'], None, NHIDDEN, dactv, a_strides,
              EPI_GATE, round_out=1, gate=actv)
        if pending is not None:
            pending.wait()
        dx_low = torch.empty_like(x, memory_format=torch.channels_last) if upsample else None
        with _Timed('k3_spade_bwd_dx', 4.0 * P * C * (2.5 if upsample else 3), (r, C)):
            L.check(lib.ag2v_spade_bwd_dx(L.ptr(x), L.ptr(dx), L.ptr(mean), L.ptr(rstd), L.ptr(sums), float(count),
                                          int(training), Pg, C, G, uh, uw, L.ptr(dx_low), L.stream()))
        if upsample:
            dx = dx_low
        # shared convolution: bias / weight gradients, then the gradient w.r.t. the (strided) segmap
        part1 = torch.empty(lib.ag2v_chan_partial_floats(P, NHIDDEN, 2), device=dev, dtype=torch.float32)
        sums1 = torch.empty(2 * NHIDDEN, device=dev, dtype=torch.float64)
        L.check(lib.ag2v_bn_stats(L.ptr(dactv), P, NHIDDEN, 1, L.ptr(part1), L.ptr(sums1), L.stream()))
        db_sh = torch.empty(NHIDDEN, device=dev, dtype=torch.float32)
        L.check(lib.ag2v_double_to_float(L.ptr(sums1), NHIDDEN, 1, 0, L.ptr(db_sh), L.stream()))
        dw_sh, _ = _wgrad(dactv, NHIDDEN, seg, seg_strides, Lc, B, r, rw, False, w_sh, None)
        dseg = None
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            full = (Hs == r and Ws == rw)                  # this layer's segmap view is the whole segmap
            fresh = False
            if shared is not None:
                buf, fresh = shared.grad_buffer_for_full_cover() if full else (shared.grad_buffer(), False)
            else:
                buf = torch.empty(B, Lc, Hs, Ws, device=dev, dtype=torch.float32, memory_format=torch.channels_last)
                fresh = full
                if not full:
                    buf.zero_()
                dseg = buf
            # the first full-resolution contribution writes (no zero fill, no read of the buffer), the others add
            _conv(dactv, a_strides, B, r, rw, NHIDDEN, pkt['w1t'], None, Lc, buf, seg_strides, EPI_BIAS if fresh else EPI_ACCUM)
        dtoken = torch.zeros(1, device=dev) if (shared is not None and ctx.needs_input_grad[2]) else None
        return dx, dseg, dtoken, dw_sh, db_sh, dw_g, db[C:], dw_b, db[:C], None, None, None, None, dscale, None, None, None


_PACK_EPOCH = [0]


def invalidate_packs():
    """Declare every packed (GEMM-layout, TF32-rounded) weight copy stale.  ``Trainer`` calls this from a
    post-step hook on each of its optimisers; call it yourself after any in-place weight update that
    neither bumps ``Tensor._version`` (fused optimisers do not) nor follows a backward through the module."""
    _PACK_EPOCH[0] += 1


class _PackCache:
    """Packed (GEMM-layout, TF32-rounded) copies of a module's weights.  A pack is reused only while ALL
    of these are unchanged: the tensors' (data_ptr, _version); the global epoch advanced by
    ``invalidate_packs()`` (optimiser post-step hooks); and a per-module epoch that advances at the first
    forward after a backward (fused optimisers update parameters without bumping _version, so for callers
    that step an optimiser of their own without the hook "a backward has happened" is the remaining signal
    that the weights may have changed).  Within one step every pack is built at most once."""

    def __init__(self):
        self.epoch, self.dirty, self.store = 0, False, {}

    def forward_begin(self):
        if self.dirty:
            self.epoch += 1
            self.dirty = False

    def backward_seen(self):
        self.dirty = True

    def get(self, name, tensors, build):
        key = (self.epoch, _PACK_EPOCH[0], _precise()) + tuple((t.data_ptr(), t._version) for t in tensors if t is not None)
        hit = self.store.get(name)
        if hit is None or hit[0] != key:
            hit = (key, build())
            self.store[name] = hit
        return hit[1]


def _cache_of(module):
    cache = module.__dict__.get('_ag2v_packs')
    if cache is None:
        cache = module.__dict__['_ag2v_packs'] = _PackCache()
    return cache


def _packed_cl(conv, w, dgrad):
    """[9][Co][Ci] (or the transposed + flipped dgrad form) of a channels_last 3x3 weight, rounded to TF32."""
    def build():
        Co, Ci = w.shape[0], w.shape[1]
        dst = torch.empty(9 * Co * Ci, device=w.device, dtype=torch.float32)
        L.check(L.lib().ag2v_pack_w3x3_cl(L.ptr(w), None, None, None, Co, Ci, int(dgrad), int(not _precise()), L.ptr(dst), None,
                                          L.stream()))
        return dst
    return _cache_of(conv).get('dgrad' if dgrad else 'fwd', (w,), build)


class _SnConvFn(torch.autograd.Function):
    """y = conv3x3(x, weight) * scale[group] + bias (+ res): a spectrally normalised 3x3 convolution
    of SPADEResnetBlock (architecture.py:34-41,56-62) evaluated on weight_orig, 1/sigma applied in
    the GEMM epilogue together with the bias and the residual sum.  ``scale`` gets its gradient
    from the SPADE layer that produced ``x`` (SPADE.forward(next_scale=...)), not from here."""

    @staticmethod
    def forward(ctx, x, weight, bias, scale, res, conv, groups, want_stats=False):
        L.need_cuda(x, weight)
        dev = x.device
        B, Cin, r, rw = x.shape
        Nout = weight.shape[0]
        if not x.is_contiguous(memory_format=torch.channels_last) or x.dtype != torch.float32:
            raise RuntimeError('sn_conv3x3: x must be a float32 channels_last tensor')
        if res is not None:
            res = _cl(res.float())
        out = torch.empty(B, Nout, r, rw, device=dev, dtype=torch.float32, memory_format=torch.channels_last)
        Pg = B * r * rw // groups
        _cache_of(conv).forward_begin()
        wpk = _packed_cl(conv, weight, False)
        sums = None
        lib = L.lib()
        mt, tpg = L.ctypes.c_int(0), L.ctypes.c_int(0)
        if (want_stats and FUSE_STATS and CONV_IMPL in (0, 2)
                and lib.ag2v_conv3x3_stats_info(B, r, rw, Cin, Nout, groups, L.ctypes.byref(mt), L.ctypes.byref(tpg))):
            # the batch norm that consumes this output gets its sums from the epilogue: no extra pass over the tensor
            part = torch.empty(mt.value * 2 * Nout, device=dev, dtype=torch.float32)
            sums = torch.empty(groups * 2 * Nout, device=dev, dtype=torch.float64)
            with _Timed('conv3x3', 2.0 * 9 * B * r * rw * Cin * Nout, ('epi0+stats', r, Cin, Nout)):
                L.check(lib.ag2v_conv3x3_bias_stats(L.ptr(x), B, r, rw, Cin, L.ptr(wpk), L.ptr(bias), Nout, L.ptr(out), 0,
                                                    Pg if groups > 1 else 0, L.ptr(scale), L.ptr(res), groups, L.ptr(part),
                                                    L.ptr(sums), L.stream()))
        else:
            _conv(x, (r * rw * Cin, rw * Cin, Cin), B, r, rw, Cin, wpk, bias, Nout, out, (r * rw * Nout, rw * Nout, Nout), EPI_BIAS,
                  group_pixels=Pg if groups > 1 else 0, scale=scale, res=res)
        ctx.save_for_backward(x, weight, scale)
        ctx.meta = (B, Cin, r, rw, Nout, groups, bias is not None, res is not None)
        ctx.conv = conv
        if sums is None:
            return out, None
        ctx.mark_non_differentiable(sums)
        return out, sums

    @staticmethod
    def backward(ctx, dy, _dsums=None):
        x, weight, scale = ctx.saved_tensors
        B, Cin, r, rw, Nout, G, has_bias, has_res = ctx.meta
        _cache_of(ctx.conv).backward_seen()
        lib = L.lib()
        dev = x.device
        dy = _cl(dy.float())
        Pg = B * r * rw // G
        dys = torch.empty_like(dy, memory_format=torch.channels_last)
        part = torch.empty(G * lib.ag2v_chan_partial_floats(Pg, Nout, 1), device=dev, dtype=torch.float32)
        sums = torch.empty(G * Nout, device=dev, dtype=torch.float64)
        L.check(lib.ag2v_scaled_grad_pre(L.ptr(dy), L.ptr(scale), Pg, Nout, G, L.ptr(dys), L.ptr(part), L.ptr(sums), L.stream()))
        dbias = None
        if has_bias:
            dbias = torch.empty(Nout, device=dev, dtype=torch.float32)
            L.check(lib.ag2v_double_to_float(L.ptr(sums), Nout, G, Nout, L.ptr(dbias), L.stream()))
        y_strides = (r * rw * Nout, rw * Nout, Nout)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(B, Cin, r, rw, device=dev, dtype=torch.float32, memory_format=torch.channels_last)
            _conv(dys, y_strides, B, r, rw, Nout, _packed_cl(ctx.conv, weight, True), None, Cin, dx,
                  (r * rw * Cin, rw * Cin, Cin), EPI_BIAS)
        nsplit = lib.ag2v_wgrad3x3_nsplit(B, r, rw, Nout, Cin, CONV_IMPL)
        wpart = torch.empty(nsplit * 9 * Nout * Cin, device=dev, dtype=torch.float32)
        with _Timed('wgrad3x3', 2.0 * 9 * B * r * rw * Cin * Nout, ('wgrad', r, Cin, Nout)):
            L.check(lib.ag2v_wgrad3x3(L.ptr(dys), Nout, L.ptr(x), r * rw * Cin, rw * Cin, Cin, Cin, B, r, rw, L.ptr(wpart),
                                      CONV_IMPL, L.stream()))
        dw = torch.empty_like(weight, memory_format=torch.channels_last)
        L.check(lib.ag2v_unpack_dw3x3_cl(L.ptr(wpart), nsplit, Nout, Cin, 0, L.ptr(dw), None, L.stream()))
        return dx, dw, dbias, None, (dy if has_res else None), None, None, None


def sn_conv3x3(conv, x, groups=1, res=None, stats=True):
    """``conv(x) (+ res)`` for a spectrally normalised 3x3 nn.Conv2d in sigma mode on the tcgen05
    implicit-GEMM kernel; ``x`` must come from ``SPADE.forward(..., next_scale=scale, round_out=True)``."""
    entry = conv.__dict__['_ag2v_sn_entry']
    y, sums = _SnConvFn.apply(x, getattr(conv, entry.name + '_orig'), conv.bias, entry.scale_g, res, conv, int(groups),
                              bool(stats and conv.training))
    if sums is not None:
        y._ag2v_stats = (sums, int(groups))       # per-group [sum | sum of squares] of y, read by the SPADE that normalises y
    return y


def plain_conv3x3(conv, x):
    """``conv(x)`` for a plain (not spectrally normalised) 3x3 / stride 1 / padding 1 ``nn.Conv2d`` with channels_last
    weights on the tcgen05 implicit-GEMM kernel, forward, input and weight gradient - the generator's ``fc``
    (spade_generator.py:16,56).  ``x``: float32 channels_last, already rounded to TF32 (e.g. ``SharedSeg.nearest``)."""
    return _SnConvFn.apply(x, conv.weight, conv.bias, None, None, conv, 1, False)[0]


def plain_conv3x3_usable(conv, x):
    w = conv.weight
    return (tuple(w.shape[2:]) == (3, 3) and conv.stride == (1, 1) and conv.padding == (1, 1) and conv.dilation == (1, 1)
            and conv.groups == 1 and w.shape[1] % 4 == 0 and w.shape[0] % 4 == 0 and not w.is_contiguous()
            and w.is_contiguous(memory_format=torch.channels_last) and x.is_cuda and x.dtype == torch.float32
            and x.is_contiguous(memory_format=torch.channels_last) and not _precise())


def sn_conv3x3_usable(conv, x):
    entry = conv.__dict__.get('_ag2v_sn_entry')
    w = getattr(conv, entry.name + '_orig', None) if entry is not None else None
    return (entry is not None and entry.scale_g is not None and w is not None and tuple(w.shape[2:]) == (3, 3)
            and conv.stride == (1, 1) and conv.padding == (1, 1) and conv.dilation == (1, 1) and conv.groups == 1
            and w.shape[1] % 4 == 0 and w.shape[0] % 4 == 0
            and w.is_contiguous(memory_format=torch.channels_last) and x.is_cuda)


class _BnActFn(torch.autograd.Function):
    """act(batch_norm(in_scale_g * x) * weight + bias) on an NHWC batch of ``groups`` reference calls
    (group-major), statistics and running-stat updates per group: the conv -> SyncBN ->
    LeakyReLU(0.2) stages around the SPADE generator (normalization.py:16-50).  ``in_scale``
    ([groups], optional) is the 1/sigma of the spectrally normalised convolution that produced x
    on weight_orig; it is folded into eps (BN(s x; eps) == BN(x; eps / s^2)), never multiplied in."""

    @staticmethod
    def forward(ctx, x, weight, bias, in_scale, running_mean, running_var, training, momentum, eps, slope, groups, sync=True):
        L.need_cuda(x, weight, bias)
        lib = L.lib()
        dev = x.device
        x = _cl(x.float())
        B, C, H, W = x.shape
        if C % 4 or B % groups:
            raise NotImplementedError('bn_act: channels %% 4 == 0 and batch %% groups == 0 required (got C=%d, B=%d, groups=%d)'
                                      % (C, B, groups))
        P = B * H * W
        G = groups if (training or in_scale is not None) else 1
        Pg = P // G
        sc = L.f32c(in_scale.detach()) if in_scale is not None else None
        mean = torch.empty(G * C, device=dev, dtype=torch.float32)
        rstd = torch.empty(G * C, device=dev, dtype=torch.float32)
        count = float(Pg)
        if training:
            part = torch.empty(G * lib.ag2v_chan_partial_floats(Pg, C, 2), device=dev, dtype=torch.float32)
            sums = torch.empty(G * 2 * C, device=dev, dtype=torch.float64)
            L.check(lib.ag2v_bn_stats(L.ptr(x), Pg, C, G, L.ptr(part), L.ptr(sums), L.stream()))
            dist, world = _world() if sync else (None, 1)
            if world > 1:
                _allreduce_sums(sums, dist)
                count = float(Pg * world)
            L.check(lib.ag2v_bn_finalize(L.ptr(sums), count, 0.0, C, G, eps, momentum, L.ptr(sc), L.ptr(running_mean),
                                         L.ptr(running_var), L.ptr(mean), L.ptr(rstd), L.stream()))
        else:
            L.check(lib.ag2v_bn_eval_stats(L.ptr(running_mean), L.ptr(running_var), C, G, eps, L.ptr(sc), L.ptr(mean),
                                           L.ptr(rstd), L.stream()))
        w, b = L.f32c(weight), L.f32c(bias)
        y = torch.empty_like(x, memory_format=torch.channels_last)
        L.check(lib.ag2v_bn_act_fwd(L.ptr(x), L.ptr(mean), L.ptr(rstd), L.ptr(w), L.ptr(b), Pg, C, G, float(slope), L.ptr(y),
                                    L.stream()))
        ctx.save_for_backward(x, y, mean, rstd, w, sc)
        ctx.meta = (C, Pg, G, count, training, slope, eps, bool(sync))
        return y

    @staticmethod
    def backward(ctx, dout):
        x, y, mean, rstd, w, sc = ctx.saved_tensors
        C, Pg, G, count, training, slope, eps, sync = ctx.meta
        lib = L.lib()
        dev = x.device
        dout = _cl(dout.float())
        dx = torch.empty_like(x, memory_format=torch.channels_last)
        part = torch.empty(G * lib.ag2v_chan_partial_floats(Pg, C, 5), device=dev, dtype=torch.float32)
        sums = torch.empty(G * 5 * C, device=dev, dtype=torch.float64)
        L.check(lib.ag2v_spade_bwd_pre(L.ptr(dout), L.ptr(y), L.ptr(x), L.ptr(w), L.ptr(mean), L.ptr(rstd), Pg, C, G,
                                       0 if slope == 1.0 else 1, float(slope), 0, 1, 0, 0, None, L.ptr(dx), L.ptr(part),
                                       L.ptr(sums), L.stream()))
        db = torch.empty(2 * C, device=dev, dtype=torch.float32)           # [d bias | d weight]
        L.check(lib.ag2v_double_to_float(L.ptr(sums), 2 * C, G, 5 * C, L.ptr(db), L.stream()))
        dscale = None
        if sc is not None and ctx.needs_input_grad[3]:
            # sums[g] = [sum g, sum g xhat, sum dxhat, sum dxhat xhat, .] with dxhat = g * weight
            s = sums.view(G, 5, C)
            r, m, s_ = rstd.view(G, C).double(), mean.view(G, C).double(), sc.double().view(G, 1)
            if training:     # y depends on s only through rstd = (var + eps / s^2)^-1/2:  dy/ds = xhat w rstd^2 eps / s^3
                dscale = (eps * (r * r * s[:, 3]).sum(dim=1) / s_.view(G) ** 3).float()
            else:            # y = (s x - rm) r0 w + b with r0 = rstd / s:  dy/ds = x r0 w,  x = xhat / rstd + mean
                dscale = ((r / s_) * (s[:, 3] / r + m * s[:, 2])).sum(dim=1).float()
        if training and sync:
            dist, world = _world()
            if world > 1:
                _allreduce_sums(sums, dist)
        L.check(lib.ag2v_spade_bwd_dx(L.ptr(x), L.ptr(dx), L.ptr(mean), L.ptr(rstd), L.ptr(sums), float(count),
                                      int(training), Pg, C, G, 0, 0, None, L.stream()))
        return dx, db[C:], db[:C], dscale, None, None, None, None, None, None, None, None


def bn_act(x, weight, bias, running_mean, running_var, training, momentum=0.1, eps=1e-5, slope=1.0, groups=1,
           in_scale=None, sync=True):
    return _BnActFn.apply(x, weight, bias, in_scale, running_mean, running_var, bool(training), float(momentum), float(eps),
                          float(slope), int(groups), bool(sync))


_IN_CONST = {}


def instance_norm_act(x, eps=1e-5, slope=1.0):
    """``leaky_relu(instance_norm(x), slope)`` (nn.InstanceNorm2d(affine=False, track_running_stats=False) followed by
    nn.LeakyReLU: the PatchGAN trunk, discriminator.py:373-380) on the grouped batch-norm kernels: one statistics group
    per image, never synchronised across ranks (the statistics are per sample by definition), NHWC in and out."""
    N, C = x.shape[:2]
    key = (C, str(x.device))
    if key not in _IN_CONST:
        _IN_CONST[key] = (torch.ones(C, device=x.device), torch.zeros(C, device=x.device))
    one, zero = _IN_CONST[key]
    scratch_m, scratch_v = torch.zeros(C, device=x.device), torch.ones(C, device=x.device)      # running stats are not tracked
    return bn_act(x, one, zero, scratch_m, scratch_v, True, 0.0, eps, slope=slope, groups=N, sync=False)


class _ParamFreeNorm(nn.Module):
    """Buffers of SynchronizedBatchNorm2d(affine=False) (normalization.py:76-77) —
    storage only; the statistics are computed by the SPADE kernels."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1):
        super().__init__()
        self.num_features, self.eps, self.momentum = num_features, eps, momentum
        self.register_buffer('running_mean', torch.zeros(num_features))
        self.register_buffer('running_var', torch.ones(num_features))
        self.register_buffer('num_batches_tracked', torch.tensor(0, dtype=torch.long))


class SPADE(nn.Module):
    """SPADE(config_text, norm_nc, label_nc).forward(x, segmap)  (normalization.py:66-110).
    ``segmap`` may also be a ``SharedSeg`` handle.  ``fused_slope`` (used by
    SPADEResnetBlock) folds ``leaky_relu(., slope)`` into the epilogue."""

    def __init__(self, config_text, norm_nc, label_nc):
        super().__init__()
        assert config_text.startswith('spade')
        parsed = re.search(r'spade(\D+)(\d)x\d', config_text)
        kind, ks = str(parsed.group(1)), int(parsed.group(2))
        if kind not in ('syncbatch', 'batch'):
            if kind == 'instance':
                raise NotImplementedError('spadeinstance is not on the accelerated path')
            raise ValueError('%s is not a recognized param-free norm type in SPADE' % kind)
        if ks != 3:
            raise NotImplementedError('the implicit-GEMM kernels are 3x3 (the reference default)')
        self.param_free_norm = _ParamFreeNorm(norm_nc)
        self.mlp_shared = nn.Sequential(nn.Conv2d(label_nc, NHIDDEN, kernel_size=3, padding=1), nn.ReLU())
        self.mlp_gamma = nn.Conv2d(NHIDDEN, norm_nc, kernel_size=3, padding=1)
        self.mlp_beta = nn.Conv2d(NHIDDEN, norm_nc, kernel_size=3, padding=1)
        self.fused_slope = 1.0

    def _packed(self, w_sh, b_sh, w_g, b_g, w_b, b_b):
        def build():
            w1, b1 = _pack(w_sh, None, b_sh, None, False)
            w2, b2 = _pack(w_g, w_b, b_g, b_b, False)
            return dict(w1=w1, b1=b1, w2=w2, b2=b2)
        return _cache_of(self).get('fwd', (w_sh, b_sh, w_g, b_g, w_b, b_b), build)

    def _packed_t(self, w_sh, w_g, w_b):
        def build():
            w1t, _ = _pack(w_sh, None, None, None, True)
            w2t, _ = _pack(w_g, w_b, None, None, True)
            return dict(w1t=w1t, w2t=w2t)
        return _cache_of(self).get('dgrad', (w_sh, w_g, w_b), build)

    def forward(self, x, segmap, groups=1, next_scale=None, round_out=False, upsample=False):
        """``groups`` > 1: the batch holds that many reference calls (group-major); batch statistics
        and running-stat updates are per group, as if the layer were called once per group.
        ``next_scale`` [groups]: the 1/sigma vector of the scaled convolution (``sn_conv3x3``) that
        consumes this output; its gradient is produced here, from sums the backward makes anyway.
        ``round_out``: store the output rounded to TF32 (operand of a tcgen05 convolution).
        ``upsample``: the layer's input is ``F.interpolate(x, scale_factor=2)`` (nearest), evaluated
        without materialising it; the output has twice the resolution of ``x``."""
        shared = segmap if isinstance(segmap, SharedSeg) else None
        seg = shared.seg if shared is not None else segmap
        token = shared.token if shared is not None else None
        pre = getattr(x, '_ag2v_stats', None)          # left by sn_conv3x3 on its output: (sums, groups)
        pre_sums = pre[0] if (pre is not None and pre[1] == int(groups) and self.training) else None
        return _SpadeFn.apply(x, seg, token, self.mlp_shared[0].weight, self.mlp_shared[0].bias,
                              self.mlp_gamma.weight, self.mlp_gamma.bias, self.mlp_beta.weight, self.mlp_beta.bias,
                              self, shared, float(self.fused_slope), int(groups), next_scale, bool(round_out), bool(upsample),
                              pre_sums)


class SPADEResnetBlock(nn.Module):
    """SPADEResnetBlock(fin, fout, opt).forward(x, seg)  (architecture.py:21-68).
    The LeakyReLU(0.2) after norm_0 / norm_1 is fused into the SPADE epilogue;
    conv_0 / conv_1 / conv_s stay torch modules with their spectral-norm hooks."""

    def __init__(self, fin, fout, opt):
        super().__init__()
        self.learned_shortcut = (fin != fout)
        fmiddle = min(fin, fout)
        self.conv_0 = nn.Conv2d(fin, fmiddle, kernel_size=3, padding=1)
        self.conv_1 = nn.Conv2d(fmiddle, fout, kernel_size=3, padding=1)
        if self.learned_shortcut:
            self.conv_s = nn.Conv2d(fin, fout, kernel_size=1, bias=False)
        if 'spectral' in opt.norm_G:
            self.conv_0 = spectral_norm(self.conv_0)
            self.conv_1 = spectral_norm(self.conv_1)
            if self.learned_shortcut:
                self.conv_s = spectral_norm(self.conv_s)
        cfg = opt.norm_G.replace('spectral', '')
        self.norm_0 = SPADE(cfg, fin, opt.semantic_nc)
        self.norm_1 = SPADE(cfg, fmiddle, opt.semantic_nc)
        self.norm_0.fused_slope = 0.2
        self.norm_1.fused_slope = 0.2
        if self.learned_shortcut:
            self.norm_s = SPADE(cfg, fin, opt.semantic_nc)
        self.__dict__['_sn'] = SpectralNormGroup(self)        # spectral norm through csrc/k5_specnorm.cu

    def forward(self, x, seg, groups=1, upsample=False):
        """``groups`` > 1: the batch holds that many reference calls (see SPADE.forward); the caller
        has put the spectral norms in sigma mode (SpectralNormGroup.refresh_sigma).
        ``upsample``: evaluate the block on ``F.interpolate(x, scale_factor=2)`` (the ``up(x)`` in
        front of every block of the generator); with a learned shortcut x only feeds SPADE layers,
        which read it through the up-sampling, so the 4x larger tensor is never written."""
        if groups == 1:
            self._sn.refresh_stale()                          # no-op when the generator prepared the weights
        if not isinstance(seg, SharedSeg):
            seg = SharedSeg.wrap(seg)
        if upsample and not self.learned_shortcut:             # x itself is the residual: materialise it
            x, upsample = F.interpolate(x, scale_factor=2, mode='nearest'), False
        up = dict(upsample=True) if upsample else {}
        x_s = conv_scaled(self.conv_s, self.norm_s(x, seg, groups, **up)) if self.learned_shortcut else x
        if sn_conv3x3_usable(self.conv_0, x) and sn_conv3x3_usable(self.conv_1, x):
            # sigma mode: the 3x3 convolutions run on the tcgen05 kernel with 1/sigma, bias and the
            # residual sum in the epilogue; d(1/sigma) comes out of the SPADE backward
            sc0 = self.conv_0.__dict__['_ag2v_sn_entry'].scale_g
            sc1 = self.conv_1.__dict__['_ag2v_sn_entry'].scale_g
            dx = sn_conv3x3(self.conv_0, self.norm_0(x, seg, groups, next_scale=sc0, round_out=True, **up), groups)
            return sn_conv3x3(self.conv_1, self.norm_1(dx, seg, groups, next_scale=sc1, round_out=True), groups, res=x_s)
        dx = conv_scaled(self.conv_0, self.norm_0(x, seg, groups, **up))
        dx = conv_scaled(self.conv_1, self.norm_1(dx, seg, groups))
        return x_s + dx

    def actvn(self, x):
        return F.leaky_relu(x, 2e-1)

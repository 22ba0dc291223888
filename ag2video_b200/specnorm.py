"""Batched spectral normalisation (csrc/k5_specnorm.cu).

The reference wraps every convolution of SPADEResnetBlock (architecture.py:34-41), of the flow
network and of conv_dim_in (normalization.py:16-50) in ``torch.nn.utils.spectral_norm``: a
forward pre-hook that recomputes ``module.weight`` from ``weight_orig / weight_u / weight_v``
on every call, one power iteration per call in training mode.  ``SpectralNormGroup`` keeps
those modules exactly as torch built them (same parameters, buffers, state-dict keys and
per-call semantics) and replaces the arithmetic of the hooks by ONE kernel launch for all
weights of a call (one more for the backward).

    group = SpectralNormGroup(root_module)   # takes over the hooks of every SN module below root
    group.refresh()                          # at the start of each root forward

A module whose weight was not refreshed by the group since its last use (stand-alone use of a
block) refreshes itself through the same kernel, so the per-call contract holds either way.
"""
import ctypes
import os

import torch
from torch.nn.utils.spectral_norm import SpectralNorm

from . import _lib as L

_pp = ctypes.POINTER(ctypes.c_void_p)
_ip = ctypes.POINTER(ctypes.c_int)
_sp = ctypes.POINTER(ctypes.c_size_t)
L.register('ag2v_spectral_norm_sizes', L.c_i, [L.c_i, _ip, _ip, _ip, _sp, _sp, _sp])
L.register('ag2v_spectral_norm_fwd', L.c_i, [L.c_i, _pp, _pp, _pp, _pp, _ip, _ip, _ip, _ip, L.c_p, L.c_sz, L.c_p, L.c_sz,
                                             L.c_i, L.c_f, L.c_p])
L.register('ag2v_spectral_norm_bwd', L.c_i, [L.c_i, _pp, _pp, _pp, _ip, _ip, _ip, _ip, L.c_p, L.c_sz, L.c_p, L.c_sz, L.c_p])

L.register('ag2v_spectral_norm_sigma_fwd', L.c_i, [L.c_i, _pp, _pp, _pp, _ip, _ip, _ip, _ip, L.c_i, L.c_p, L.c_sz, L.c_p, L.c_sz,
                                                   L.c_i, L.c_f, L.c_p])
L.register('ag2v_spectral_norm_sigma_bwd', L.c_i, [L.c_i, L.c_p, _pp, _ip, _ip, _ip, _ip, L.c_i, L.c_p, L.c_sz, L.c_p])

L.register('ag2v_spectral_norm_scale_bwd_one', L.c_i, [L.c_i, L.c_i, L.c_p, L.c_p, L.c_i, L.c_p, _ip, _ip, _ip, _ip, L.c_i,
                                                       L.c_p, L.c_sz, L.c_p])

MAX_PER_LAUNCH = 48
# Sigma mode, training: optionally every weight gets its own autograd node for the 1/sigma scales (one small launch per
# weight in the backward) instead of one batched node for all weights.  A batched node runs last, and autograd hands a
# leaf its gradient only when ALL of its consumers have run - so with it every spectrally normalised weight (two thirds
# of the generator's gradient bytes) completes at the very end of the backward pass and its gradient bucket cannot be
# all-reduced earlier.  Measured on 2 B200s the step time is the same either way (52.6 ms: NCCL's kernels do not get SMs
# next to the persistent convolutions), so the cheaper batched node stays the default;
# AG2V_SN_PER_WEIGHT_GRAD=1 selects the per-weight nodes.
PER_WEIGHT_SIGMA_GRAD = os.environ.get('AG2V_SN_PER_WEIGHT_GRAD', '0') == '1'


def _ptr_array(tensors):
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def _int_array(values):
    return (ctypes.c_int * len(values))(*values)


def _geometry(ws):
    """(co, cin, taps, channels_last) per weight; only dense NCHW / channels_last storage."""
    co, cin, taps, cl = [], [], [], []
    for w in ws:
        if w.dtype != torch.float32:
            raise RuntimeError('spectral norm: float32 weights only')
        shape = tuple(w.shape) + (1,) * (4 - w.dim())
        if w.is_contiguous():
            cl.append(0)
        elif w.dim() == 4 and w.is_contiguous(memory_format=torch.channels_last):
            cl.append(1)
        else:
            raise RuntimeError('spectral norm: weight must be contiguous or channels_last, strides %s' % (w.stride(),))
        co.append(shape[0])
        if w.dim() == 4:
            cin.append(shape[1])
            taps.append(shape[2] * shape[3])
        else:                                   # [out, in] matrices (nn.Linear)
            cin.append(int(w.numel() // shape[0]))
            taps.append(1)
    return co, cin, taps, cl


def _sizes(co, cin, taps):
    save, fwd, bwd = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t()
    L.check(L.lib().ag2v_spectral_norm_sizes(len(co), _int_array(co), _int_array(cin), _int_array(taps),
                                             ctypes.byref(save), ctypes.byref(fwd), ctypes.byref(bwd)))
    return save.value, fwd.value, bwd.value


class _SpectralNormFn(torch.autograd.Function):
    """(*weight_orig) -> (*weight); u / v buffers are updated in place when ``training``."""

    @staticmethod
    def forward(ctx, us, vs, training, eps, *ws):
        L.need_cuda(*ws)
        co, cin, taps, cl = _geometry(ws)
        save_n, fwd_n, _ = _sizes(co, cin, taps)
        dev = ws[0].device
        outs = [torch.empty_like(w) for w in ws]                  # keeps the storage order of w
        save = torch.empty(save_n, device=dev, dtype=torch.float32)
        scratch = torch.empty(fwd_n, device=dev, dtype=torch.float32)
        L.check(L.lib().ag2v_spectral_norm_fwd(len(ws), _ptr_array(ws), _ptr_array(outs), _ptr_array(us), _ptr_array(vs),
                                               _int_array(co), _int_array(cin), _int_array(taps), _int_array(cl),
                                               L.ptr(save), save_n, L.ptr(scratch), fwd_n, int(training), float(eps),
                                               L.stream()))
        ctx.save_for_backward(save, *ws)
        ctx.geom = (co, cin, taps, cl)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gs):
        save, *ws = ctx.saved_tensors
        co, cin, taps, cl = ctx.geom
        _, _, bwd_n = _sizes(co, cin, taps)
        fmt = lambda g, w, c: (torch.zeros_like(w) if g is None else
                               g.float().contiguous(memory_format=torch.channels_last) if c else g.float().contiguous())
        gs = [fmt(g, w, c) for g, w, c in zip(gs, ws, cl)]
        dws = [torch.empty_like(w) for w in ws]
        scratch = torch.empty(bwd_n, device=save.device, dtype=torch.float32)
        L.check(L.lib().ag2v_spectral_norm_bwd(len(ws), _ptr_array(ws), _ptr_array(gs), _ptr_array(dws), _int_array(co),
                                               _int_array(cin), _int_array(taps), _int_array(cl), L.ptr(save),
                                               save.numel(), L.ptr(scratch), bwd_n, L.stream()))
        return (None, None, None, None) + tuple(dws)


class _SpectralSigmaFn(torch.autograd.Function):
    """(*weight_orig) -> sigma [iters, n]: ``iters`` successive power iterations (one per frame group
    of a batched call).  The convolutions then run on weight_orig and scale group g's output by
    1 / sigma[g]; autograd brings back d sigma, whose weight gradient is sum_g dsigma_g u_g v_g^T."""

    @staticmethod
    def forward(ctx, us, vs, training, eps, iters, *ws):
        L.need_cuda(*ws)
        co, cin, taps, cl = _geometry(ws)
        save_n, fwd_n, _ = _sizes(co, cin, taps)
        dev = ws[0].device
        k = iters if training else 1               # eval mode: u / v are fixed, every group sees the same sigma
        save = torch.empty(k * save_n, device=dev, dtype=torch.float32)
        scratch = torch.empty(fwd_n, device=dev, dtype=torch.float32)
        L.check(L.lib().ag2v_spectral_norm_sigma_fwd(len(ws), _ptr_array(ws), _ptr_array(us), _ptr_array(vs), _int_array(co),
                                                     _int_array(cin), _int_array(taps), _int_array(cl), k, L.ptr(save),
                                                     save.numel(), L.ptr(scratch), fwd_n, int(training), float(eps),
                                                     L.stream()))
        n = len(ws)
        n4 = (n + 3) & ~3
        sigma = save[:k * n4].view(k, n4)[:, :n]
        sigma = sigma.expand(iters, n).contiguous() if k != iters else sigma.contiguous()
        ctx.save_for_backward(save, *ws)
        ctx.geom, ctx.k, ctx.iters = (co, cin, taps, cl), k, iters
        return sigma

    @staticmethod
    def backward(ctx, dsigma):
        save, *ws = ctx.saved_tensors
        co, cin, taps, cl = ctx.geom
        if ctx.k != ctx.iters:
            dsigma = dsigma.sum(dim=0, keepdim=True)
        dsigma = dsigma.float().contiguous()
        dws = [torch.empty_like(w) for w in ws]
        L.check(L.lib().ag2v_spectral_norm_sigma_bwd(len(ws), L.ptr(dsigma), _ptr_array(dws), _int_array(co), _int_array(cin),
                                                     _int_array(taps), _int_array(cl), ctx.k, L.ptr(save), save.numel(),
                                                     L.stream()))
        return (None, None, None, None, None) + tuple(dws)


class _ScaleOfWeight(torch.autograd.Function):
    """(weight_orig, 1/sigma per image [N], 1/sigma per group [G]) -> (scale [N,1,1,1], scale_g [G]) for ONE weight of a
    batched sigma call: the values were computed by the batched forward kernel, this node only ties them to the weight.
    Backward: dW = sum_g coef_g u_g v_g^T with coef_g = -(d scale_g[g] + sum of d scale over the images of g) / sigma_g^2."""

    @staticmethod
    def forward(ctx, w, inv_img, inv_g, save, geom, index, images_per_group):
        ctx.save_for_backward(save)
        ctx.info = (geom, index, images_per_group, inv_g.numel(), w)
        return inv_img.detach().view(-1, 1, 1, 1), inv_g.detach()

    @staticmethod
    def backward(ctx, d_img, d_g):
        (save,) = ctx.saved_tensors
        (co, cin, taps, cl), index, ipg, iters, w = ctx.info
        d_img = None if d_img is None else d_img.float().contiguous()
        d_g = None if d_g is None else d_g.float().contiguous()
        if d_img is None and d_g is None:
            return (None,) * 7
        dw = torch.empty_like(w)
        L.check(L.lib().ag2v_spectral_norm_scale_bwd_one(len(co), index, L.ptr(d_g), L.ptr(d_img), ipg, L.ptr(dw), _int_array(co),
                                                         _int_array(cin), _int_array(taps), _int_array(cl), iters, L.ptr(save),
                                                         save.numel(), L.stream()))
        return (dw,) + (None,) * 6


def spectral_sigmas(weights, us, vs, iters, training=True, eps=1e-12):
    """sigma [iters, n] for lists of weight_orig / weight_u / weight_v."""
    outs = []
    for i in range(0, len(weights), MAX_PER_LAUNCH):
        j = i + MAX_PER_LAUNCH
        outs.append(_SpectralSigmaFn.apply(list(us[i:j]), list(vs[i:j]), training, eps, iters, *weights[i:j]))
    return outs[0] if len(outs) == 1 else torch.cat(outs, dim=1)


def spectral_normalize(weights, us, vs, training=True, eps=1e-12):
    """Normalised weights for lists of weight_orig / weight_u / weight_v (any number)."""
    outs = []
    for i in range(0, len(weights), MAX_PER_LAUNCH):
        j = i + MAX_PER_LAUNCH
        outs += _SpectralNormFn.apply(list(us[i:j]), list(vs[i:j]), training, eps, *weights[i:j])
    return outs


class _Entry:
    __slots__ = ('module', 'name', 'eps', 'fresh', 'scale', 'scale_g')

    def __init__(self, module, name, eps):
        self.module, self.name, self.eps, self.fresh = module, name, eps, False
        self.scale = None          # sigma mode: [N,1,1,1] per-image 1/sigma of the current batched call
        self.scale_g = None        # sigma mode: [groups] 1/sigma per frame group (contiguous)

    def tensors(self):
        m, n = self.module, self.name
        return getattr(m, n + '_orig'), getattr(m, n + '_u'), getattr(m, n + '_v')


def _compute(entries):
    """One launch per (training, eps) combination for the given entries."""
    groups = {}
    for e in entries:
        groups.setdefault((e.module.training, e.eps), []).append(e)
    for (training, eps), es in groups.items():
        trip = [e.tensors() for e in es]
        outs = spectral_normalize([t[0] for t in trip], [t[1] for t in trip], [t[2] for t in trip], training, eps)
        for e, w in zip(es, outs):
            setattr(e.module, e.name, w)
            e.fresh = True
            e.scale = e.scale_g = None


def _make_hook(entry):
    def hook(module, _inputs):
        if not entry.fresh:                # nobody prepared this call's weight: do it alone
            _compute([entry])
        entry.fresh = False                # consumed by this call
    return hook


class SpectralNormGroup:
    """Takes over torch's spectral-norm forward pre-hooks of every module under ``root``
    (modules already taken over by another group are shared, not duplicated)."""

    def __init__(self, root):
        self.entries = []
        for module in root.modules():
            entry = module.__dict__.get('_ag2v_sn_entry')
            if entry is not None:
                self.entries.append(entry)
                continue
            for key, hook in list(module._forward_pre_hooks.items()):
                if not isinstance(hook, SpectralNorm):
                    continue
                if hook.dim != 0 or hook.n_power_iterations != 1:
                    raise NotImplementedError('spectral norm with dim=%d, n_power_iterations=%d'
                                              % (hook.dim, hook.n_power_iterations))
                del module._forward_pre_hooks[key]
                entry = _Entry(module, hook.name, hook.eps)
                module.register_forward_pre_hook(_make_hook(entry))
                module.__dict__['_ag2v_sn_entry'] = entry
                self.entries.append(entry)

    def refresh(self):
        """Compute the weight of every module for the call that is about to happen."""
        _compute(self.entries)

    def refresh_sigma(self, groups, images_per_group):
        """Sigma mode for a call that batches ``groups`` reference calls (group-major batch of
        groups * images_per_group images): one power iteration per group, in order; afterwards
        ``conv_scaled`` evaluates each convolution on weight_orig with a per-image 1/sigma."""
        modes = {(e.module.training, e.eps) for e in self.entries}
        if len(modes) != 1:
            raise RuntimeError('spectral norm sigma mode: modules disagree on training mode / eps')
        (training, eps), = modes
        trip = [e.tensors() for e in self.entries]
        if PER_WEIGHT_SIGMA_GRAD and training and torch.is_grad_enabled():
            return self._refresh_sigma_per_weight(trip, groups, images_per_group, eps)
        sigma = spectral_sigmas([t[0] for t in trip], [t[1] for t in trip], [t[2] for t in trip], groups, training, eps)
        inv = sigma.reciprocal()                                                     # [G, n]
        inv_img = inv.repeat_interleave(images_per_group, dim=0)                     # [N, n]
        inv_t = inv.t().contiguous()                                                 # [n, G]: rows are contiguous
        for i, e in enumerate(self.entries):
            e.scale = inv_img[:, i].view(-1, 1, 1, 1)
            e.scale_g = inv_t[i]
            e.fresh = False
        return sigma

    def _refresh_sigma_per_weight(self, trip, groups, images_per_group, eps):
        """The batched forward kernel without an autograd node of its own; one _ScaleOfWeight node per weight."""
        ws, us, vs = [t[0] for t in trip], [t[1] for t in trip], [t[2] for t in trip]
        sigmas, chunks = [], []
        with torch.no_grad():
            for i in range(0, len(ws), MAX_PER_LAUNCH):
                j = i + MAX_PER_LAUNCH
                L.need_cuda(*ws[i:j])
                geom = _geometry(ws[i:j])
                save_n, fwd_n, _ = _sizes(*geom[:3])
                dev = ws[i].device
                save = torch.empty(groups * save_n, device=dev, dtype=torch.float32)
                scratch = torch.empty(fwd_n, device=dev, dtype=torch.float32)
                L.check(L.lib().ag2v_spectral_norm_sigma_fwd(len(ws[i:j]), _ptr_array(ws[i:j]), _ptr_array(us[i:j]), _ptr_array(vs[i:j]),
                                                             _int_array(geom[0]), _int_array(geom[1]), _int_array(geom[2]),
                                                             _int_array(geom[3]), groups, L.ptr(save), save.numel(), L.ptr(scratch),
                                                             fwd_n, 1, float(eps), L.stream()))
                n = len(ws[i:j])
                n4 = (n + 3) & ~3
                sigmas.append(save[:groups * n4].view(groups, n4)[:, :n])
                chunks.append((save, geom))
            sigma = sigmas[0].contiguous() if len(sigmas) == 1 else torch.cat(sigmas, dim=1)          # [G, n]
            inv_t = sigma.reciprocal().t().contiguous()                                                # [n, G]
            inv_img_t = inv_t.repeat_interleave(images_per_group, dim=1).contiguous()                  # [n, N]
        for i, e in enumerate(self.entries):
            save, geom = chunks[i // MAX_PER_LAUNCH]
            if ws[i].requires_grad:
                e.scale, e.scale_g = _ScaleOfWeight.apply(ws[i], inv_img_t[i], inv_t[i], save, geom, i % MAX_PER_LAUNCH,
                                                          images_per_group)
            else:
                e.scale, e.scale_g = inv_img_t[i].view(-1, 1, 1, 1), inv_t[i]
            e.fresh = False
        return sigma

    def refresh_stale(self):
        """Same, for the modules whose weight has not been prepared yet (nested use)."""
        _compute([e for e in self.entries if not e.fresh and e.scale is None])

    def end_sigma(self):
        """Leave sigma mode (the autograd graph of the finished call keeps what it needs)."""
        for e in self.entries:
            e.scale = e.scale_g = None


def conv_scaled(conv, x):
    """``conv(x)`` for a spectrally normalised nn.Conv2d.  In sigma mode (after
    ``refresh_sigma``) the convolution runs on weight_orig and image n of the output is scaled by
    1/sigma of its group; otherwise this is a plain module call (hooks or ``refresh``)."""
    entry = conv.__dict__.get('_ag2v_sn_entry')
    if entry is None or entry.scale is None:
        return conv(x)
    z = torch.nn.functional.conv2d(x, getattr(conv, entry.name + '_orig'), None, conv.stride, conv.padding,
                                   conv.dilation, conv.groups)
    if conv.bias is None:
        return z * entry.scale
    return torch.addcmul(conv.bias.view(1, -1, 1, 1), z, entry.scale)


def conv_unscaled(conv, x):
    """(z, scale): in sigma mode z = conv on weight_orig (no bias allowed) and scale = the pending
    per-group 1/sigma [groups] for the consumer to fold in (a batch norm absorbs it into eps);
    otherwise (conv(x), None)."""
    entry = conv.__dict__.get('_ag2v_sn_entry')
    if entry is None or entry.scale_g is None or conv.bias is not None:
        return conv_scaled(conv, x), None
    z = torch.nn.functional.conv2d(x, getattr(conv, entry.name + '_orig'), None, conv.stride, conv.padding,
                                   conv.dilation, conv.groups)
    return z, entry.scale_g

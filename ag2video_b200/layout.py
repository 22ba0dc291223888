"""Layout composition on sm_100a (drop-in for models/layout.py).

``boxes_to_layout`` keeps the signature of layout.py:28-63.  The batched form
``boxes_to_layout_batched`` composes every (clip, frame) of a batch in one
launch and takes the callers' object masks (models/utils.py:95-102) as a device
tensor instead of boolean-index compaction, so there is no host sync.
"""
import ctypes

import torch

from . import _lib as L

_LIN = {}


def _linspace(n, device):
    """The reference builds ``torch.linspace(0, 1, steps=n)`` on the CPU and moves it
    to the boxes' device (layout.py:116-117), so the CPU values are the contract
    on every device; torch's CUDA linspace rounds differently."""
    key = (n, str(device))
    t = _LIN.get(key)
    if t is None:
        t = torch.linspace(0, 1, steps=n).to(device)
        _LIN[key] = t
    return t


class _BoxesToLayoutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vecs, boxes, valid, H, W, avg):
        L.need_cuda(vecs, boxes, valid)
        vecs, boxes = L.f32c(vecs), L.f32c(boxes)
        N, O, D = vecs.shape
        lib = L.lib()
        dev = vecs.device
        if valid is not None:
            valid = valid.contiguous()
            valid = valid.view(torch.uint8) if valid.dtype == torch.bool else (valid != 0).view(torch.uint8)
        lin_x, lin_y = _linspace(W, dev), _linspace(H, dev)
        ws = torch.empty(max(lib.ag2v_boxes_to_layout_workspace_bytes(N, O, H, W), 16), device=dev, dtype=torch.uint8)
        out = torch.empty(N, D, H, W, device=dev, dtype=torch.float32)
        with L.timed('k2_layout_fwd', 4.0 * N * D * H * W, (N, D, H)):
            L.check(lib.ag2v_boxes_to_layout_fwd(L.ptr(vecs), L.ptr(boxes), L.ptr(valid), L.ptr(lin_x), L.ptr(lin_y),
                                                 N, O, D, H, W, int(avg), L.ptr(ws), L.ptr(out), L.stream()))
        ctx.save_for_backward(ws, vecs, boxes)
        ctx.dims = (N, O, D, H, W, int(avg))
        return out

    @staticmethod
    def backward(ctx, dout):
        ws, vecs, boxes = ctx.saved_tensors
        N, O, D, H, W, avg = ctx.dims
        dout = L.f32c(dout)
        dvecs = dboxes = None
        if ctx.needs_input_grad[0]:
            dvecs = torch.zeros(N, O, D, device=dout.device, dtype=torch.float32)
            with L.timed('k2_layout_bwd', 4.0 * N * D * H * W, (N, D, H)):
                L.check(L.lib().ag2v_boxes_to_layout_bwd(L.ptr(dout), None, None, None, None, N, O, D, H, W, avg, 0,
                                                         L.ptr(ws), L.ptr(dvecs), L.stream()))
        # boxes are data on the training path (meta_models.py:53 passes ground truth or detached predictions);
        # a caller that does differentiate them gets what autograd gives the reference through grid_sample
        if ctx.needs_input_grad[1]:
            dev = dout.device
            dboxes = torch.zeros(N, O, 4, device=dev, dtype=torch.float32)
            L.check(L.lib().ag2v_boxes_to_layout_dboxes(L.ptr(dout), D * H * W, L.ptr(vecs), L.ptr(boxes),
                                                        L.ptr(_linspace(W, dev)), L.ptr(_linspace(H, dev)), N, O, D, H, W,
                                                        L.ptr(ws), L.ptr(dboxes), L.stream()))
        return dvecs, dboxes, None, None, None, None


def _pooling_flag(pooling):
    if pooling == 'sum':
        return False
    if pooling == 'avg':
        return True
    raise ValueError('Invalid pooling "%s"' % pooling)


def boxes_to_layout_batched(vecs, boxes, valid, H, W=None, pooling='sum'):
    """vecs [N,O,D], boxes [N,O,4] xywh, valid [N,O] bool/uint8 or None -> [N,D,H,W].
    All-zero boxes are dropped inside the kernel (layout.py:40-42)."""
    W = H if W is None else W
    return _BoxesToLayoutFn.apply(vecs, boxes, valid, int(H), int(W), _pooling_flag(pooling))


def boxes_to_layout(vecs, boxes, H, W=None, pooling='sum'):
    """vecs [O,D], boxes [O,4] xywh in [0,1] -> [1,D,H,W]  (layout.py:28-63)."""
    avg = _pooling_flag(pooling)
    W = H if W is None else W
    return _BoxesToLayoutFn.apply(vecs.unsqueeze(0), boxes.unsqueeze(0), None, int(H), int(W), avg)


L.register('ag2v_boxes_to_layout_dboxes', L.c_i, [L.c_p, ctypes.c_longlong] + [L.c_p] * 4 + [L.c_i] * 5 + [L.c_p] * 2 + [L.c_p])
L.register('ag2v_boxes_to_layout_fwd_strided', L.c_i, [L.c_p] * 5 + [L.c_i] * 6 + [L.c_p] * 2 + [ctypes.c_longlong, L.c_p])
L.register('ag2v_boxes_to_layout_bwd_strided', L.c_i, [L.c_p, ctypes.c_longlong] + [L.c_p] * 4 + [L.c_i] * 7 + [L.c_p] * 2 + [L.c_p])


class _LayoutCatFn(torch.autograd.Function):
    """x = cat([img, boxes_to_layout(vecs, boxes)], dim=1) in ONE pass: the layout kernel writes
    channels [C, C+D) of the NCHW result directly (batch-strided output), the image is copied
    into channels [0, C).  Backward: dimg = dx[:, :C]; dvecs read from dx[:, C:] in place."""

    @staticmethod
    def forward(ctx, img, vecs, boxes, valid, H, W):
        L.need_cuda(img, vecs, boxes, valid)
        vecs, boxes = L.f32c(vecs), L.f32c(boxes)
        N, O, D = vecs.shape
        C = img.shape[1]
        if tuple(img.shape) != (N, C, H, W):
            raise RuntimeError('layout_cat: img %s does not match N=%d H=%d W=%d' % (tuple(img.shape), N, H, W))
        lib = L.lib()
        dev = vecs.device
        if valid is not None:
            valid = valid.contiguous()
            valid = valid.view(torch.uint8) if valid.dtype == torch.bool else (valid != 0).view(torch.uint8)
        ws = torch.empty(max(lib.ag2v_boxes_to_layout_workspace_bytes(N, O, H, W), 16), device=dev, dtype=torch.uint8)
        x = torch.empty(N, C + D, H, W, device=dev, dtype=torch.float32)
        x[:, :C].copy_(img)
        L.check(lib.ag2v_boxes_to_layout_fwd_strided(L.ptr(vecs), L.ptr(boxes), L.ptr(valid), L.ptr(_linspace(W, dev)),
                                                     L.ptr(_linspace(H, dev)), N, O, D, H, W, 0, L.ptr(ws),
                                                     ctypes.c_void_p(x.data_ptr() + 4 * C * H * W), (C + D) * H * W, L.stream()))
        ctx.save_for_backward(ws)
        ctx.dims = (N, O, D, H, W, C)
        return x

    @staticmethod
    def backward(ctx, dx):
        (ws,) = ctx.saved_tensors
        N, O, D, H, W, C = ctx.dims
        dx = L.f32c(dx)
        dimg = dx[:, :C] if ctx.needs_input_grad[0] else None
        dvecs = None
        if ctx.needs_input_grad[1]:
            dvecs = torch.zeros(N, O, D, device=dx.device, dtype=torch.float32)
            L.check(L.lib().ag2v_boxes_to_layout_bwd_strided(ctypes.c_void_p(dx.data_ptr() + 4 * C * H * W), (C + D) * H * W,
                                                             None, None, None, None, N, O, D, H, W, 0, 0, L.ptr(ws),
                                                             L.ptr(dvecs), L.stream()))
        return dimg, dvecs, None, None, None, None


def layout_cat(img, vecs, boxes, valid, H, W=None):
    """cat([img, boxes_to_layout_batched(vecs, boxes, valid, H, W)], dim=1) without building
    the layout or the copy of it: img [N,C,H,W], vecs [N,O,D], boxes [N,O,4] -> [N,C+D,H,W]
    (the discriminator input of discriminator.py:317-342)."""
    W = H if W is None else W
    return _LayoutCatFn.apply(img, vecs, boxes, valid, int(H), int(W))


L.register('ag2v_masks_to_layout_fwd', L.c_i, [L.c_p] * 5 + [L.c_i] * 6 + [L.c_p] * 3 + [L.c_p])
L.register('ag2v_masks_to_layout_bwd', L.c_i, [L.c_p] * 2 + [L.c_i] * 4 + [L.c_p] + [L.c_p])
L.register('ag2v_masks_to_layout_bwd_inputs', L.c_i, [L.c_p] * 6 + [L.c_i] * 5 + [L.c_p] * 3 + [L.c_p])


class _MasksToLayoutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vecs, boxes, masks, H, W, test_mode):
        L.need_cuda(vecs, boxes, masks)
        vecs, boxes, masks = L.f32c(vecs), L.f32c(boxes), L.f32c(masks)
        O, D = vecs.shape
        M = masks.shape[1]
        dev = vecs.device
        S = torch.empty(max(O, 1), H, W, device=dev, dtype=torch.float32)
        order = torch.empty(max(O, 1), device=dev, dtype=torch.int32)
        out = torch.empty(1, D, H, W, device=dev, dtype=torch.float32)
        L.check(L.lib().ag2v_masks_to_layout_fwd(L.ptr(vecs), L.ptr(boxes), L.ptr(masks), L.ptr(_linspace(W, dev)),
                                                 L.ptr(_linspace(H, dev)), O, D, M, H, W, int(test_mode), L.ptr(S),
                                                 L.ptr(order), L.ptr(out), L.stream()))
        ctx.save_for_backward(S, vecs, boxes, masks)
        ctx.dims = (O, D, M, H, W, bool(test_mode))
        return out

    @staticmethod
    def backward(ctx, dout):
        S, vecs, boxes, masks = ctx.saved_tensors
        O, D, M, H, W, test_mode = ctx.dims
        if test_mode:
            raise RuntimeError('masks_to_layout(test_mode=True) is an inference-only compositing path')
        dev = dout.device
        dout = L.f32c(dout)
        dvecs = dboxes = dmasks = None
        if ctx.needs_input_grad[0]:
            dvecs = torch.zeros(O, D, device=dev, dtype=torch.float32)
            L.check(L.lib().ag2v_masks_to_layout_bwd(L.ptr(dout), L.ptr(S), O, D, H, W, L.ptr(dvecs), L.stream()))
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            # what autograd gives the reference through F.grid_sample (layout.py:87-91): mask taps and grid gradient
            if ctx.needs_input_grad[1]:
                dboxes = torch.zeros(O, 4, device=dev, dtype=torch.float32)
            if ctx.needs_input_grad[2]:
                dmasks = torch.zeros(O, M, M, device=dev, dtype=torch.float32)
            G = torch.empty(max(O, 1), H, W, device=dev, dtype=torch.float32)
            L.check(L.lib().ag2v_masks_to_layout_bwd_inputs(L.ptr(dout), L.ptr(vecs), L.ptr(boxes), L.ptr(masks),
                                                            L.ptr(_linspace(W, dev)), L.ptr(_linspace(H, dev)), O, D, M, H, W,
                                                            L.ptr(G), L.ptr(dmasks), L.ptr(dboxes), L.stream()))
        return dvecs, dboxes, dmasks, None, None, None


def masks_to_layout(vecs, boxes, masks, H, W=None, pooling='sum', test_mode=False):
    """vecs [O,D], boxes [O,4] xywh, masks [O,M,M] -> [1,D,H,W]  (layout.py:66-95).
    ``test_mode`` composites objects in ascending order of sampled mass, first writer
    wins where the clean mask exceeds 0.5 (layout.py:185-197)."""
    O, D = vecs.size()
    M = masks.size(1)
    assert masks.size() == (O, M, M)
    W = H if W is None else W
    out = _MasksToLayoutFn.apply(vecs, boxes, masks.float(), int(H), int(W), bool(test_mode))
    if pooling != 'sum':
        raise ValueError('Invalid pooling "%s"' % pooling)
    return out


# ---- K4: layout fused into its consumer convolution -------------
L.register('ag2v_boxes_to_layout_tables', L.c_i, [L.c_p] * 4 + [L.c_i] * 4 + [L.c_p] + [L.c_p])
L.register('ag2v_layout_conv_bwd_workspace_floats', L.c_sz, [L.c_i] * 4)
L.register('ag2v_layout_conv_fwd', L.c_i, [L.c_p, L.c_p] + [L.c_i] * 5 + [L.c_p] + [L.c_p])
L.register('ag2v_layout_conv_bwd', L.c_i, [L.c_p, L.c_p] + [L.c_i] * 5 + [L.c_p, L.c_p] + [L.c_p])


def layout_tables(boxes, valid, H, W):
    """Separable weight tables of K2 for boxes [N,S,4] / valid [N,S] (opaque workspace tensor)."""
    L.need_cuda(boxes, valid)
    boxes = L.f32c(boxes)
    N, S = boxes.shape[:2]
    dev = boxes.device
    if valid is not None:
        valid = valid.contiguous()
        valid = valid.view(torch.uint8) if valid.dtype == torch.bool else (valid != 0).view(torch.uint8)
    ws = torch.empty(max(L.lib().ag2v_boxes_to_layout_workspace_bytes(N, S, H, W), 16), device=dev, dtype=torch.uint8)
    L.check(L.lib().ag2v_boxes_to_layout_tables(L.ptr(boxes), L.ptr(valid), L.ptr(_linspace(W, dev)), L.ptr(_linspace(H, dev)),
                                                N, S, H, W, L.ptr(ws), L.stream()))
    return ws


class _LayoutConvFn(torch.autograd.Function):
    """base [N,Co,H,W] (channels_last, modified IN PLACE) += sum_s sum_k U[n,s,k,:] * m_s(p+k)."""

    @staticmethod
    def forward(ctx, base, U, tables, S):
        L.need_cuda(base, U, tables)
        N, Co, H, W = base.shape
        if not base.is_contiguous(memory_format=torch.channels_last) or base.dtype != torch.float32:
            raise RuntimeError('layout_conv: base must be a float32 channels_last tensor')
        U = L.f32c(U)
        with L.timed('k4_layout_conv_fwd', 8.0 * N * H * W * Co, (N, Co, H)):      # output tile read-modify-write
            L.check(L.lib().ag2v_layout_conv_fwd(L.ptr(U), L.ptr(tables), N, S, Co, H, W, L.ptr(base), L.stream()))
        ctx.mark_dirty(base)
        ctx.save_for_backward(tables)
        ctx.dims = (N, S, Co, H, W)
        return base

    @staticmethod
    def backward(ctx, dout):
        (tables,) = ctx.saved_tensors
        N, S, Co, H, W = ctx.dims
        dout = dout.float().contiguous(memory_format=torch.channels_last)
        lib = L.lib()
        part = torch.empty(lib.ag2v_layout_conv_bwd_workspace_floats(N, S, Co, H), device=dout.device, dtype=torch.float32)
        dU = torch.empty(N, S, 9, Co, device=dout.device, dtype=torch.float32)
        with L.timed('k4_layout_conv_bwd', 4.0 * N * H * W * Co, (N, Co, H)):
            L.check(lib.ag2v_layout_conv_bwd(L.ptr(dout), L.ptr(tables), N, S, Co, H, W, L.ptr(part), L.ptr(dU), L.stream()))
        return dout, dU, None, None


def layout_conv3x3(weight, vec_slots, tables, base):
    """conv3x3(layout(vec_slots), weight) added in place to ``base`` without ever building the layout.
    weight [Co, n_slots*D, 3, 3] (the layout channels of the consumer conv), vec_slots: list of
    [N,O,D] object vectors, one per frame slot (their boxes went into ``tables``, slot-major),
    base [N,Co,H,W] channels_last (e.g. the convolution of the remaining image channels)."""
    Co = weight.shape[0]
    D = vec_slots[0].shape[-1]
    N, O = vec_slots[0].shape[:2]
    us = []
    for j, v in enumerate(vec_slots):
        wj = weight[:, j * D:(j + 1) * D].permute(2, 3, 0, 1).reshape(9 * Co, D)          # [(k, co), ci]
        us.append((v.reshape(N * O, D) @ wj.t()).view(N, O, 9, Co))
    U = torch.cat(us, dim=1)                                                              # [N, slots*O, 9, Co]
    return _LayoutConvFn.apply(base, U, tables, U.shape[1])


# ---- K7: the PatchGAN stem through the rank-1 layout ---------------
L.register('ag2v_layout_tables_avgpool', L.c_i, [L.c_p] + [L.c_i] * 4 + [L.c_p] + [L.c_p])
L.register('ag2v_layout_sconv_bwd_workspace_floats', L.c_sz, [L.c_i] * 5)
L.register('ag2v_layout_sconv_fwd', L.c_i, [L.c_p, L.c_p] + [L.c_i] * 8 + [L.c_p] + [L.c_p])
L.register('ag2v_layout_sconv_bwd', L.c_i, [L.c_p, L.c_p] + [L.c_i] * 8 + [L.c_p, L.c_p] + [L.c_p])


def pooled_size(n):
    """Output size of avg_pool2d(kernel 3, stride 2, padding 1)."""
    return (n - 1) // 2 + 1


def layout_tables_avgpool(tables, N, S, H, W):
    """Tables of the 3x3 / stride-2 / pad-1 average pool (count_include_pad=False) of the masks in
    ``tables`` (boxes [N,S,4] at H x W): the pooled masks stay separable, so every consumer of the
    tables works unchanged at the coarser scale (discriminator.py:271)."""
    H2, W2 = pooled_size(H), pooled_size(W)
    out = torch.empty(max(L.lib().ag2v_boxes_to_layout_workspace_bytes(N, S, H2, W2), 16), device=tables.device, dtype=torch.uint8)
    L.check(L.lib().ag2v_layout_tables_avgpool(L.ptr(tables), N, S, H, W, L.ptr(out), L.stream()))
    return out


class _LayoutSConvFn(torch.autograd.Function):
    """base [N,Co,Ho,Wo] (channels_last, modified IN PLACE) += sum_s sum_k U[n,s,k,:] * m_s(stride*p + k - pad)."""

    @staticmethod
    def forward(ctx, base, U, tables, H, W, kernel, stride, pad):
        L.need_cuda(base, U, tables)
        N, Co, Ho, Wo = base.shape
        S = U.shape[1]
        if not base.is_contiguous(memory_format=torch.channels_last) or base.dtype != torch.float32:
            raise RuntimeError('layout_sconv: base must be a float32 channels_last tensor')
        if (Ho, Wo) != ((H + 2 * pad - kernel) // stride + 1, (W + 2 * pad - kernel) // stride + 1):
            raise RuntimeError('layout_sconv: base %s does not match input %dx%d' % (tuple(base.shape), H, W))
        U = L.f32c(U)
        with L.timed('k7_layout_sconv_fwd', 8.0 * N * Ho * Wo * Co, (N, Co, Ho)):
            L.check(L.lib().ag2v_layout_sconv_fwd(L.ptr(U), L.ptr(tables), N, S, Co, H, W, kernel, stride, pad, L.ptr(base), L.stream()))
        ctx.mark_dirty(base)
        ctx.save_for_backward(tables)
        ctx.dims = (N, S, Co, H, W, Ho, kernel, stride, pad)
        return base

    @staticmethod
    def backward(ctx, dout):
        (tables,) = ctx.saved_tensors
        N, S, Co, H, W, Ho, kernel, stride, pad = ctx.dims
        dout = dout.float().contiguous(memory_format=torch.channels_last)
        dU = None
        if ctx.needs_input_grad[1]:
            lib = L.lib()
            KK = kernel * kernel
            part = torch.empty(lib.ag2v_layout_sconv_bwd_workspace_floats(N, S, Co, Ho, KK), device=dout.device, dtype=torch.float32)
            dU = torch.empty(N, S, KK, Co, device=dout.device, dtype=torch.float32)
            with L.timed('k7_layout_sconv_bwd', 4.0 * N * Ho * Ho * Co, (N, Co, Ho)):
                L.check(lib.ag2v_layout_sconv_bwd(L.ptr(dout), L.ptr(tables), N, S, Co, H, W, kernel, stride, pad, L.ptr(part),
                                                  L.ptr(dU), L.stream()))
        return dout, dU, None, None, None, None, None, None


def layout_sconv(weight, vecs, tables, base, H, W, stride=2, pad=2):
    """conv2d(layout(vecs), weight, stride, pad) added in place to ``base`` without building the layout:
    weight [Co, D, k, k] (the layout channels of the consumer convolution), vecs [N,O,D], tables for the
    boxes [N,O,4] at the convolution's input resolution H x W, base [N,Co,Ho,Wo] channels_last."""
    Co, D, k, _ = weight.shape
    N, O = vecs.shape[:2]
    wk = weight.permute(2, 3, 0, 1).reshape(k * k * Co, D)                     # [(ky, kx, co), ci]
    U = (vecs.reshape(N * O, D) @ wk.t()).view(N, O, k * k, Co)
    return _LayoutSConvFn.apply(base, U, tables, int(H), int(W), int(k), int(stride), int(pad))

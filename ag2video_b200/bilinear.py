"""Differentiable bilinear crops on sm_100a (drop-in for models/bilinear.py).

``crop_bbox_batch`` / ``crop_bbox`` keep the reference signatures
(bilinear.py:29-57, 102-131).  The reference replicates every frame once per
object before sampling (bilinear.py:86); here each crop indexes its source frame.
"""
import torch

from . import _lib as L
from .networks import real_object_mask

L.register('ag2v_crop_bbox_fwd', L.c_i, [L.c_p] * 7 + [L.c_i] * 6 + [L.c_p] + [L.c_p])
L.register('ag2v_crop_bbox_bwd', L.c_i, [L.c_p] * 7 + [L.c_i] * 6 + [L.c_p] + [L.c_p])
L.register('ag2v_crop_bbox_jj_fwd', L.c_i, [L.c_p] * 7 + [L.c_i] * 6 + [L.c_p] + [L.c_p])
L.register('ag2v_crop_bbox_jj_bwd', L.c_i, [L.c_p] * 7 + [L.c_i] * 6 + [L.c_p] + [L.c_p])

_LIN = {}


def _lerp_weights(steps, device):
    """torch.linspace(1, 0, steps) / (0, 1, steps) computed on the CPU like
    tensor_linspace does (bilinear.py:212-214), cached per device."""
    key = (steps, str(device))
    if key not in _LIN:
        _LIN[key] = (torch.linspace(1, 0, steps=steps).to(device), torch.linspace(0, 1, steps=steps).to(device))
    return _LIN[key]


class _CropFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, frame, boxes, HH, WW, jj=False):
        L.need_cuda(feats, frame, boxes)
        feats, boxes = L.f32c(feats), L.f32c(boxes)
        frame = frame.to(torch.int32).contiguous()
        NF, C, H, W = feats.shape
        n = boxes.shape[0]
        ws_x, we_x = _lerp_weights(WW, feats.device)
        ws_y, we_y = _lerp_weights(HH, feats.device)
        out = torch.empty(n, C, HH, WW, device=feats.device, dtype=torch.float32)
        fn = L.lib().ag2v_crop_bbox_jj_fwd if jj else L.lib().ag2v_crop_bbox_fwd
        L.check(fn(L.ptr(feats), L.ptr(frame), L.ptr(boxes), L.ptr(ws_x), L.ptr(we_x),
                   L.ptr(ws_y), L.ptr(we_y), n, C, H, W, HH, WW, L.ptr(out), L.stream()))
        ctx.save_for_backward(frame, boxes)
        ctx.dims = (NF, C, H, W, HH, WW, bool(jj))
        return out

    @staticmethod
    def backward(ctx, dout):
        frame, boxes = ctx.saved_tensors
        NF, C, H, W, HH, WW, jj = ctx.dims
        n = boxes.shape[0]
        ws_x, we_x = _lerp_weights(WW, dout.device)
        ws_y, we_y = _lerp_weights(HH, dout.device)
        dfeats = torch.zeros(NF, C, H, W, device=dout.device, dtype=torch.float32)
        fn = L.lib().ag2v_crop_bbox_jj_bwd if jj else L.lib().ag2v_crop_bbox_bwd
        L.check(fn(L.ptr(L.f32c(dout)), L.ptr(frame), L.ptr(boxes), L.ptr(ws_x), L.ptr(we_x),
                   L.ptr(ws_y), L.ptr(we_y), n, C, H, W, HH, WW, L.ptr(dfeats), L.stream()))
        return dfeats, None, None, None, None, None


def crop_bbox(feats, bbox, HH, WW=None, backend='cudnn'):
    """feats [N,C,H,W], bbox [N,4] xywh -> [N,C,HH,WW]  (bilinear.py:102-131).  backend 'cudnn' is grid_sample
    (align_corners=True, zero padding) on the box mapped to [-1, 1]; 'jj' is bilinear_sample (bilinear.py:134-189):
    coordinates scaled by the image size, taps clamped to the image."""
    if backend not in ('cudnn', 'jj'):
        raise ValueError('crop_bbox: backend must be "cudnn" or "jj" (got %r)' % (backend,))
    N = feats.size(0)
    assert bbox.size(0) == N and bbox.size(1) == 4
    WW = HH if WW is None else WW
    frame = torch.arange(N, device=feats.device, dtype=torch.int32)
    return _CropFn.apply(feats, frame, bbox, int(HH), int(WW), backend == 'jj')


def crop_bbox_batch(imgs, objs, bbox, HH, WW=None, vocab=None, backend='cudnn'):
    """imgs [B,N,C,H,W], objs [B,O,A], bbox [B,N,O,4] xywh -> (list of per-clip crops
    [sum_frames n, C, HH, WW], list of flattened attribute rows)  (bilinear.py:29-44,67-99).
    Dummy / padding objects and all-zero boxes are dropped; order is frame-major, object-minor.
    ``backend`` is accepted and ignored like the reference does (bilinear.py:95 always samples with 'cudnn')."""
    L.need_cuda(imgs, objs, bbox)
    B, N, C, H, W = imgs.shape
    WW = HH if WW is None else WW
    O = objs.shape[1]
    keep = real_object_mask(objs, vocab).view(B, 1, O) & (bbox != 0).any(dim=-1)          # [B,N,O]
    idx = keep.nonzero()                  # the variable-length return value needs the counts on the host
    frame = (idx[:, 0] * N + idx[:, 1]).to(torch.int32)
    boxes = bbox[idx[:, 0], idx[:, 1], idx[:, 2]]
    crops = _CropFn.apply(imgs.reshape(B * N, C, H, W), frame, boxes, int(HH), int(WW))
    attrs = objs[idx[:, 0], idx[:, 2]]
    counts = torch.bincount(idx[:, 0], minlength=B).tolist()
    return list(torch.split(crops, counts)), [a.reshape(-1) for a in torch.split(attrs, counts)]

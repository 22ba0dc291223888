"""Independent closed form of ``boxes_to_layout`` in numpy (CPU, fp32).

TEST INFRASTRUCTURE — see ``oracle/__init__.py``.  ``oracle.ops.boxes_to_layout``
restates the reference with the torch primitives it calls (``F.grid_sample``,
``scatter_add``); this module is the second opinion that does NOT go through
``grid_sample``: sampling a per-object constant 8x8 map with bilinear
interpolation, zeros padding and ``align_corners=True`` (models/layout.py:52-53)
is separable, ``out[d,i,j] = sum_o v[o,d] * wy_o(i) * wx_o(j)``, with the 1-D
weight of SURVEY.md appendix A.2 built from the reference's own coordinate
arithmetic (``_boxes_to_grid``, layout.py:98-130), every step rounded to fp32 in
the reference's order.  The *support* {wx != 0} x {wy != 0} is the bit-exact
part of the contract.
"""
import numpy as np
import torch

f32 = np.float32


def axis_weights(start, extent, n, src=8):
    """1-D weights of one object along one axis: lin = torch.linspace(0,1,n) (the reference
    builds it on the CPU, layout.py:116-117); g = ((lin - start) / extent) * 2 - 1 (:119-128);
    ix = ((g + 1) / 2) * (src - 1) (grid_sample un-normalisation, align_corners=True); taps
    floor(ix), floor(ix)+1 with weights (x_e - ix), (ix - x_w), out-of-range taps dropped."""
    lin = torch.linspace(0, 1, steps=n).numpy().astype(f32)
    with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
        g = ((lin - f32(start)) / f32(extent)).astype(f32)
        g = (g * f32(2)).astype(f32) - f32(1)
        ix = ((g + f32(1)) / f32(2)).astype(f32) * f32(src - 1)
        ix = ix.astype(f32)
        x_w = np.floor(ix)
        x_e = x_w + f32(1)
        w_w = (x_e - ix).astype(f32)
        w_e = (ix - x_w).astype(f32)
        ok_w = (x_w >= 0) & (x_w <= src - 1)
        ok_e = (x_e >= 0) & (x_e <= src - 1)
        w = np.where(ok_w, w_w, f32(0)) + np.where(ok_e, w_e, f32(0))
    w = np.where(np.isfinite(ix), w, f32(0)).astype(f32)         # CUDA grid_sample: non-finite coordinates sample nothing
    return w


def boxes_to_layout(vecs, boxes, H, W=None, pooling='sum'):
    """vecs [O,D], boxes [O,4] xywh (torch or numpy) -> (out [1,D,H,W] float32 numpy, support [O,H,W] bool).
    All-zero boxes are skipped (layout.py:40-42)."""
    W = H if W is None else W
    vecs = np.asarray(vecs, dtype=f32)
    boxes = np.asarray(boxes, dtype=f32)
    O, D = vecs.shape
    out = np.zeros((D, H, W), dtype=np.float64)
    support = np.zeros((O, H, W), dtype=bool)
    n_legal = 0
    for o in range(O):
        if not boxes[o].any():
            continue
        n_legal += 1
        x0, y0, ww, hh = boxes[o]
        wx, wy = axis_weights(x0, ww, W), axis_weights(y0, hh, H)
        m = np.outer(wy, wx)
        support[o] = m != 0
        out += vecs[o][:, None, None].astype(np.float64) * m[None].astype(np.float64)
    if pooling == 'avg':
        out /= max(n_legal, 1)
    elif pooling != 'sum':
        raise ValueError('Invalid pooling "%s"' % pooling)
    return out.astype(f32)[None], support

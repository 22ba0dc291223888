"""Layout composition on sm_100a (drop-in for models/layout.py).

``boxes_to_layout`` keeps the signature of layout.py:28-63.  The batched form
``boxes_to_layout_batched`` composes every (clip, frame) of a batch in one
launch and takes the callers' object masks (models/utils.py:95-102) as a device
tensor instead of boolean-index compaction, so there is no host sync.
"""
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
        L.check(lib.ag2v_boxes_to_layout_fwd(L.ptr(vecs), L.ptr(boxes), L.ptr(valid), L.ptr(lin_x), L.ptr(lin_y),
                                             N, O, D, H, W, int(avg), L.ptr(ws), L.ptr(out), L.stream()))
        ctx.save_for_backward(ws)
        ctx.dims = (N, O, D, H, W, int(avg))
        return out

    @staticmethod
    def backward(ctx, dout):
        (ws,) = ctx.saved_tensors
        N, O, D, H, W, avg = ctx.dims
        dout = L.f32c(dout)
        dvecs = torch.zeros(N, O, D, device=dout.device, dtype=torch.float32)
        L.check(L.lib().ag2v_boxes_to_layout_bwd(L.ptr(dout), None, None, None, None, N, O, D, H, W, avg, 0,
                                                 L.ptr(ws), L.ptr(dvecs), L.stream()))
        # boxes are data on the training path (meta_models.py:53 passes ground truth
        # or detached predictions), so no gradient is produced for them.
        return dvecs, None, None, None, None, None


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

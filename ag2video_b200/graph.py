"""Action-graph convolution on sm_100a (drop-in for models/graph_models/graph.py).

``GraphTripleConv`` keeps the reference constructor, ``forward`` signature and
state-dict keys (``net1.{0,2}.{weight,bias}``, ``net2.{0,2}.{weight,bias}``) of
graph.py:16-107; the forward and backward passes are each one cooperative CUDA
kernel (csrc/k1_gcn.cu).  ``GraphTripleConvNet`` is the ``gconvs`` loop of
models/graph_models/model.py:54-57,163-164.
"""
import torch
import torch.nn as nn

from . import _lib as L


def gcn_layer_bytes(B, O, E, Din, Dp, H, Dout, Dpo):
    """Algorithmic bytes of one forward layer call: parameters + inputs + outputs, fp32."""
    params = (2 * Din + Dp) * H + H + H * (2 * H + Dpo) + (2 * H + Dpo) + H * H + H + H * Dout + Dout
    return 4.0 * (params + B * O * Din + B * E * Dp + 2 * B * E + B * O * Dout + B * E * Dpo)


class _GcnLayerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, obj, pred, edges, ind, W1a, b1a, W1b, b1b, W2a, b2a, W2b, b2b, want_p):
        L.need_cuda(obj, pred, edges, ind, W1a)
        obj, pred = L.f32c(obj), L.f32c(pred)
        edges = edges.long().contiguous()
        ind = ind.contiguous()
        ind = ind.view(torch.uint8) if ind.dtype == torch.bool else (ind != 0).view(torch.uint8)
        W = [L.f32c(w) for w in (W1a, b1a, W1b, b1b, W2a, b2a, W2b, b2b)]
        B, O, Din = obj.shape
        E, Dp = pred.shape[1], pred.shape[2]
        H, Dout = W[0].shape[0], W[6].shape[0]
        Dpo = W[2].shape[0] - 2 * H
        if tuple(edges.shape) != (B, E, 2) or tuple(ind.shape) != (B, E) or W[0].shape[1] != 2 * Din + Dp:
            raise RuntimeError('GraphTripleConv: inconsistent shapes obj %s pred %s edges %s ind %s net1 %s' % (
                tuple(obj.shape), tuple(pred.shape), tuple(edges.shape), tuple(ind.shape), tuple(W[0].shape)))
        lib = L.lib()
        new_obj = torch.empty(B, O, Dout, device=obj.device, dtype=torch.float32)
        new_p = torch.empty(B, E, Dpo, device=obj.device, dtype=torch.float32)
        saved = torch.empty(lib.ag2v_gcn_layer_saved_floats(B, O, E, H, Dpo), device=obj.device, dtype=torch.float32)
        with L.timed('k1_gcn_fwd', gcn_layer_bytes(B, O, E, Din, Dp, H, Dout, Dpo), (B * E, Din)):
            L.check(lib.ag2v_gcn_layer_fwd(L.ptr(obj), L.ptr(pred), L.ptr(edges), L.ptr(ind), *[L.ptr(w) for w in W],
                                           B, O, E, Din, Dp, H, Dout, Dpo, L.ptr(new_obj), L.ptr(new_p), L.ptr(saved),
                                           L.stream()))
        ctx.dims = (B, O, E, Din, Dp, H, Dout, Dpo)
        ctx.want_p = want_p
        ctx.save_for_backward(obj, pred, edges, ind, W[0], W[2], W[4], W[6], new_obj, saved)
        return new_obj, new_p

    @staticmethod
    def backward(ctx, d_obj, d_p):
        obj, pred, edges, ind, W1a, W1b, W2a, W2b, new_obj, saved = ctx.saved_tensors
        B, O, E, Din, Dp, H, Dout, Dpo = ctx.dims
        lib = L.lib()
        dev = obj.device
        d_obj = L.f32c(d_obj) if d_obj is not None else torch.zeros_like(new_obj)
        d_p = L.f32c(d_p) if (d_p is not None and ctx.want_p) else None
        ws = torch.empty(lib.ag2v_gcn_layer_bwd_workspace_floats(B, O, E, Din, Dp, H, Dpo), device=dev, dtype=torch.float32)
        g_obj, g_pred = torch.empty_like(obj), torch.empty_like(pred)
        gW = [torch.empty_like(W1a), torch.empty(H, device=dev), torch.empty_like(W1b), torch.empty(2 * H + Dpo, device=dev),
              torch.empty_like(W2a), torch.empty(H, device=dev), torch.empty_like(W2b), torch.empty(Dout, device=dev)]
        with L.timed('k1_gcn_bwd', 2.0 * gcn_layer_bytes(B, O, E, Din, Dp, H, Dout, Dpo), (B * E, Din)):
            L.check(lib.ag2v_gcn_layer_bwd(L.ptr(obj), L.ptr(pred), L.ptr(edges), L.ptr(ind), L.ptr(W1a), L.ptr(W1b),
                                           L.ptr(W2a), L.ptr(W2b), L.ptr(new_obj), L.ptr(saved), L.ptr(d_obj), L.ptr(d_p),
                                           B, O, E, Din, Dp, H, Dout, Dpo, L.ptr(ws), L.ptr(g_obj), L.ptr(g_pred),
                                           *[L.ptr(g) for g in gW], L.stream()))
        return (g_obj, g_pred, None, None, *gW, None)


def _mlp2(d_in, d_hid, d_out):
    # models/layers.py:6-25 with batch_norm='none': weights live at indices 0 and 2
    return nn.Sequential(nn.Linear(d_in, d_hid), nn.ReLU(), nn.Linear(d_hid, d_out), nn.ReLU())


class GraphTripleConv(nn.Module):
    """Same constructor and forward contract as graph.py:16-107 (``pooling`` is
    accepted and ignored exactly like the reference: always a masked average)."""

    def __init__(self, obj_input_dim, object_output_dim, predicate_input_dim, predicate_output_dim, hidden_dim,
                 num_attributes=None, loc_dim=4, pooling='avg', mlp_normalization='none', return_new_p_vecs=True):
        super().__init__()
        assert pooling in ['sum', 'avg'], 'Invalid pooling "%s"' % pooling
        if mlp_normalization != 'none':
            # layers.py:13-14 puts BatchNorm1d(hidden) behind a Linear that sees [B, T, features] (graph.py:67-68):
            # BatchNorm1d reads dim 1 (T) as channels, so the reference itself raises for 'batch' unless T == hidden
            raise NotImplementedError('mlp_normalization=%r: the reference\'s batched graph layer cannot run it either '
                                      '(BatchNorm1d on [B, T, features]); only "none" exists' % (mlp_normalization,))
        self.return_new_p_vecs = return_new_p_vecs
        self.hidden_dim = hidden_dim
        self.num_attributes = num_attributes
        self.predicate_output_dim = predicate_output_dim
        self.pooling = pooling
        self.net1 = _mlp2(2 * obj_input_dim + predicate_input_dim, hidden_dim, 2 * hidden_dim + predicate_output_dim)
        self.net2 = _mlp2(hidden_dim, hidden_dim, object_output_dim)
        for net in (self.net1, self.net2):          # graph.py:10-13
            for m in net:
                if isinstance(m, nn.Linear):
                    nn.init.kaiming_normal_(m.weight)

    def forward(self, obj_vecs, pred_vecs, edges, pred_indicators):
        new_obj, new_p = _GcnLayerFn.apply(
            obj_vecs, pred_vecs, edges, pred_indicators,
            self.net1[0].weight, self.net1[0].bias, self.net1[2].weight, self.net1[2].bias,
            self.net2[0].weight, self.net2[0].bias, self.net2[2].weight, self.net2[2].bias,
            self.return_new_p_vecs)
        return new_obj, (new_p if self.return_new_p_vecs else pred_vecs)


class GraphTripleConvNet(nn.Module):
    """``layers`` is a list of GraphTripleConv keyword dicts; forward runs them in
    order on the same edges (model.py:163-164, discriminator.py:308-309)."""

    def __init__(self, layers):
        super().__init__()
        self.gconvs = nn.ModuleList([GraphTripleConv(**kw) for kw in layers])

    def forward(self, obj_vecs, pred_vecs, edges, pred_indicators):
        for layer in self.gconvs:
            obj_vecs, pred_vecs = layer(obj_vecs, pred_vecs, edges, pred_indicators)
        return obj_vecs, pred_vecs

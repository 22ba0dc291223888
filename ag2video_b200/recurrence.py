"""The Acts2LayoutModel recurrence as one persistent kernel per direction (csrc/k1r_recur.cu; SURVEY.md
section 8 row f2; reference loop: models/graph_models/model.py:126-169).

``run(models, embs, box0, preds, edges, ind)`` evaluates, for every model of ``models`` (same graph data,
different weights - the generator step runs ``acts_to_boxes`` and ``acts_to_objs`` side by side) and every clip,

    for t in 1 .. T-1:  x = obj_vecs_net([emb | boxes[t-1]]);  x, p = gconv_l(x, p, edges[t], ind[t]) for all l;
                        boxes[t] = boxes[t-1] + box_net(x)

in ONE launch (one thread-block cluster per (model, clip) chain).  Autograd sees one node per model, so a model
whose outputs receive no gradient (``acts_to_boxes`` in the generator step) costs nothing in the backward;
a model's backward is one chain launch for the data gradients plus one grouped GEMM for all weight gradients.
"""
import ctypes
from types import SimpleNamespace

import torch

from . import _lib as L
from ._lib import c_i, c_p, c_sz

_D10 = [c_i] * 10
L.register('ag2v_recur_cluster_size', c_i, _D10)
L.register('ag2v_recur_cluster_fits', c_i, _D10 + [c_i])
L.register('ag2v_recur_max_active_clusters', c_i, [c_i])
L.register('ag2v_recur_num_params', c_i, [c_i])
L.register('ag2v_recur_set_profile', c_i, [c_p])
L.register('ag2v_recur_set_core', c_i, [c_i])
L.register('ag2v_recur_get_core', c_i, [])
L.register('ag2v_recur_pack_floats', c_sz, _D10 + [c_i])
L.register('ag2v_recur_saved_floats', c_sz, _D10 + [c_i])
L.register('ag2v_recur_z_floats', c_sz, _D10 + [c_i])
L.register('ag2v_recur_pack', c_i, _D10 + [c_i, c_p, c_p, c_p])
L.register('ag2v_recur_fwd', c_i, _D10 + [c_i, c_i, c_i, c_p] + [c_p] * 8 + [c_p])
L.register('ag2v_recur_bwd', c_i, _D10 + [c_i, c_i, c_i, c_i] + [c_p] * 14 + [c_p])
L.register('ag2v_recur_wgrad', c_i, _D10 + [c_i, c_i, c_i] + [c_p] * 4 + [c_p])

ENABLED = True          # tests flip this to compare with the layer-by-layer path (graph.GraphTripleConv per call)
_CLUSTERS = {}


def _ptr_array(tensors):
    return (ctypes.c_void_p * len(tensors))(*[None if t is None else t.data_ptr() for t in tensors])


def model_params(m):
    """Parameters of an Acts2LayoutModel in the kernel's order (k1r_recur.cu: param_index_box)."""
    ps = [m.obj_vecs_net[0].weight, m.obj_vecs_net[2].weight]
    for g in m.gconvs:
        ps += [g.net1[0].weight, g.net1[0].bias, g.net1[2].weight, g.net1[2].bias,
               g.net2[0].weight, g.net2[0].bias, g.net2[2].weight, g.net2[2].bias]
    ps += [m.box_net[0].weight, m.box_net[0].bias, m.box_net[2].weight, m.box_net[2].bias]
    return ps


def model_dims(m, O, E, T):
    """(O, E, T, Kx, De, Dp, H, Dout, Dpo, NL) or None when the module is not the shape the kernel is written for."""
    g0 = m.gconvs[0]
    De = m.obj_vecs_net[2].weight.shape[0]
    Kx = m.obj_vecs_net[0].weight.shape[1] - 4
    H = g0.net1[0].weight.shape[0]
    Dout = g0.net2[2].weight.shape[0]
    Dpo = g0.net1[2].weight.shape[0] - 2 * H
    Dp = g0.net1[0].weight.shape[1] - 2 * De
    NL = len(m.gconvs)
    if Kx != De or m.obj_vecs_net[0].weight.shape[0] != De or m.box_net[0].weight.shape != (H, Dout) or m.box_net[2].weight.shape != (4, H):
        return None
    for l, g in enumerate(m.gconvs):
        din, dp = (De, Dp) if l == 0 else (Dout, Dpo)
        if (g.net1[0].weight.shape != (H, 2 * din + dp) or g.net1[2].weight.shape != (2 * H + Dpo, H)
                or g.net2[0].weight.shape != (H, H) or g.net2[2].weight.shape != (Dout, H)):
            return None
    return (int(O), int(E), int(T), int(Kx), int(De), int(Dp), int(H), int(Dout), int(Dpo), int(NL))


def cluster_size(dims):
    """Cluster size the kernel would use for these sizes on this device, 0 = not supported here."""
    if not ENABLED or dims is None:
        return 0
    key = dims[:2] + dims[3:]
    if key not in _CLUSTERS:
        lib = L.lib()
        cs = lib.ag2v_recur_cluster_size(*dims)
        # a device (or partition) without room for a cluster of that size: try the smaller fitting sizes
        while cs >= 1 and not (lib.ag2v_recur_cluster_fits(*dims, cs) and lib.ag2v_recur_max_active_clusters(cs) >= 1):
            cs //= 2
        _CLUSTERS[key] = cs
    return _CLUSTERS[key]


def _packed(module, dims, cs):
    from .spade import _cache_of
    params = model_params(module)

    def build():
        lib = L.lib()
        ps = [L.f32c(p.detach()) for p in params]
        pack = torch.empty(lib.ag2v_recur_pack_floats(*dims, cs), device=ps[0].device, dtype=torch.float32)
        L.check(lib.ag2v_recur_pack(*dims, cs, _ptr_array(ps), L.ptr(pack), L.stream()))
        return pack
    cache = _cache_of(module)
    cache.forward_begin()
    return cache.get('recur%d.%d' % (cs, L.lib().ag2v_recur_get_core()), params, build)


def recur_bytes(dims, nc):
    """Algorithmic bytes of one forward launch: the packed weights once per chain and timestep (they are
    re-streamed from L2 for every step of the recurrence) + activations written."""
    O, E, T, Kx, De, Dp, H, Dout, Dpo, NL = dims
    w = Kx * De + De * De + H * Dout + 4 * H
    act = 2 * O * De + O * H
    for l in range(NL):
        din, dp = (De, Dp) if l == 0 else (Dout, Dpo)
        w += H * (2 * din + dp) + (2 * H + Dpo) * H + H * H + Dout * H
        act += E * (2 * din + dp) + E * H + E * (2 * H + Dpo) + 2 * O * H + O * Dout
    return 4.0 * nc * (T - 1) * (w + act)


class _RecurFn(torch.autograd.Function):
    """One model's share of a fused launch: forward hands out the precomputed slices, backward runs the
    model's own chain + grouped weight-gradient launches."""

    @staticmethod
    def forward(ctx, emb, box0, pred, shared, index, module, *params):
        B = emb.shape[0]
        ctx.shared, ctx.index, ctx.module = shared, index, module
        ctx.save_for_backward(emb, *params)
        ctx.n_params = len(params)
        sl = slice(index * B, (index + 1) * B)
        return shared.objv[sl], shared.boxes[sl]

    @staticmethod
    def backward(ctx, d_objv, d_boxes):
        from .spade import _cache_of
        sh, m = ctx.shared, ctx.index
        emb, params = ctx.saved_tensors[0], ctx.saved_tensors[1:]
        dims, cs, B, NC = sh.dims, sh.cs, sh.B, sh.NC
        O, E, T, Kx, De, Dp, H, Dout, Dpo, NL = dims
        lib = L.lib()
        dev = emb.device
        _cache_of(ctx.module).backward_seen()
        if sh.z is None:            # gradients of every Linear's output, shared by the models of the launch
            sh.z = torch.empty(lib.ag2v_recur_z_floats(*dims, NC), device=dev, dtype=torch.float32)
        f32 = dict(device=dev, dtype=torch.float32)
        d_emb, d_box0 = torch.empty(B, O, Kx, **f32), torch.empty(B, O, 4, **f32)
        d_pred = torch.empty(B, T, E, Dp, **f32)
        dw0box, dwb2, dbb2 = torch.empty(B, De, 4, **f32), torch.empty(B, 4, H, **f32), torch.empty(B, 4, **f32)
        d_objv = L.f32c(d_objv) if d_objv is not None else None
        d_boxes = L.f32c(d_boxes) if d_boxes is not None else None
        boxes_m = sh.boxes[m * B:(m + 1) * B]
        with L.timed('k1r_recur_bwd', 2.0 * recur_bytes(dims, B), (B, T)):
            L.check(lib.ag2v_recur_bwd(*dims, cs, NC, m * B, B, L.ptr(sh.packs[m]), L.ptr(sh.saved), L.ptr(boxes_m),
                                       L.ptr(d_objv), L.ptr(d_boxes), L.ptr(sh.edges), L.ptr(sh.ind), L.ptr(sh.z),
                                       L.ptr(d_emb), L.ptr(d_box0), L.ptr(d_pred), L.ptr(dw0box), L.ptr(dwb2), L.ptr(dbb2),
                                       L.stream()))
        grads = [torch.empty_like(p, memory_format=torch.contiguous_format) for p in params]
        emb_c = L.f32c(emb)
        with L.timed('k1r_recur_wgrad', 2.0 * B * (T - 1) * E * sum(p.numel() for p in params), (B, T)):
            L.check(lib.ag2v_recur_wgrad(*dims, NC, m * B, B, L.ptr(sh.saved), L.ptr(sh.z), L.ptr(emb_c), _ptr_array(grads),
                                         L.stream()))
        grads[0][:, Kx:] = dw0box.sum(dim=0)
        grads[-2].copy_(dwb2.sum(dim=0))
        grads[-1].copy_(dbb2.sum(dim=0))
        return (d_emb, d_box0 if ctx.needs_input_grad[1] else None, d_pred, None, None, None, *grads)


def run(models, embs, box0, preds, edges, ind, needs_grad=None):
    """models: list of Acts2LayoutModel sharing the graph data; embs[m] [B,O,Kx], preds[m] [B,T,E,Dp] per model;
    box0 [B,O,4], edges [B,T,E,2] int64, ind [B,T,E] bool.  Returns [(obj_vecs [B,T,O,Dout], boxes [B,T,O,4])]
    per model, or None when the kernel does not cover these sizes (the caller then walks the layers)."""
    B, T, E = preds[0].shape[:3]
    O = embs[0].shape[1]
    dims = model_dims(models[0], O, E, T)
    if any(model_dims(m, O, E, T) != dims for m in models[1:]):
        return None
    cs = cluster_size(dims)
    if cs == 0:
        return None
    L.need_cuda(embs[0], box0, preds[0], edges, ind)
    lib = L.lib()
    dev = embs[0].device
    M = len(models)
    NC = M * B
    Dout = dims[7]
    packs = [_packed(m, dims, cs) for m in models]
    with torch.no_grad():
        emb_all = L.f32c(torch.cat([e.detach() for e in embs], dim=0) if M > 1 else embs[0].detach())
        pred_all = L.f32c(torch.cat([p.detach() for p in preds], dim=0) if M > 1 else preds[0].detach())
        box0_c = L.f32c(box0.detach())
        edges_c = edges.long().contiguous()
        ind_c = ind.contiguous()
        ind_c = ind_c.view(torch.uint8) if ind_c.dtype == torch.bool else (ind_c != 0).view(torch.uint8)
        objv = torch.empty(NC, T, O, Dout, device=dev, dtype=torch.float32)
        boxes = torch.empty(NC, T, O, 4, device=dev, dtype=torch.float32)
        saved = torch.empty(lib.ag2v_recur_saved_floats(*dims, NC), device=dev, dtype=torch.float32)
        with L.timed('k1r_recur_fwd', recur_bytes(dims, NC), (NC, T)):
            L.check(lib.ag2v_recur_fwd(*dims, cs, NC, M, _ptr_array(packs), L.ptr(emb_all), L.ptr(box0_c), L.ptr(pred_all),
                                       L.ptr(edges_c), L.ptr(ind_c), L.ptr(objv), L.ptr(boxes), L.ptr(saved), L.stream()))
    shared = SimpleNamespace(dims=dims, cs=cs, B=B, NC=NC, packs=packs, saved=saved, z=None, objv=objv, boxes=boxes,
                             edges=edges_c, ind=ind_c)
    out = []
    for i, m in enumerate(models):
        if needs_grad is not None and not needs_grad[i]:
            out.append((objv[i * B:(i + 1) * B], boxes[i * B:(i + 1) * B]))        # forward only: no autograd node
        else:
            out.append(_RecurFn.apply(embs[i], box0, preds[i], shared, i, m, *model_params(m)))
    return out

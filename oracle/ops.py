"""Oracle restatement of the hot-path operators (CPU, fp32, torch primitives).

TEST INFRASTRUCTURE — see ``oracle/__init__.py``.  Every function cites the
reference lines (relative to the reference checkout) whose behaviour it restates.
State-dict key names equal the reference's so reference weights load with
``strict=True``.
"""
import re

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils import spectral_norm


# --------------------------------------------------------------------------
# K1: action-graph convolution
# --------------------------------------------------------------------------

def _mlp2(d_in, d_hid, d_out):
    """Linear-ReLU-Linear-ReLU as produced by models/layers.py:6-25 with
    batch_norm='none', final_nonlinearity='relu' (indices 0 and 2 hold weights)."""
    return nn.Sequential(nn.Linear(d_in, d_hid), nn.ReLU(), nn.Linear(d_hid, d_out), nn.ReLU())


class GraphTripleConv(nn.Module):
    """models/graph_models/graph.py:16-107.

    net1 on [subject | predicate | object] rows for ALL edges (graph.py:70-71);
    pooling only over edges whose indicator is set (graph.py:80-84): subject
    parts first, then object parts, each in edge order (graph.py:90-91); divide
    by the incidence count where it is > 0 (graph.py:93-99; the ``pooling``
    argument is ignored); net2 on every node (graph.py:103).
    """

    def __init__(self, obj_input_dim, object_output_dim, predicate_input_dim, predicate_output_dim,
                 hidden_dim, num_attributes=None, loc_dim=4, pooling='avg', mlp_normalization='none',
                 return_new_p_vecs=True):
        super().__init__()
        assert pooling in ('sum', 'avg')
        assert mlp_normalization == 'none', 'oracle restates the live configuration only'
        self.hidden_dim = hidden_dim
        self.predicate_output_dim = predicate_output_dim
        self.return_new_p_vecs = return_new_p_vecs
        self.net1 = _mlp2(2 * obj_input_dim + predicate_input_dim, hidden_dim,
                          2 * hidden_dim + predicate_output_dim)
        self.net2 = _mlp2(hidden_dim, hidden_dim, object_output_dim)
        for net in (self.net1, self.net2):  # graph.py:10-13,35,39
            for m in net:
                if isinstance(m, nn.Linear):
                    nn.init.kaiming_normal_(m.weight)

    def forward(self, obj_vecs, pred_vecs, edges, pred_indicators):
        B, O, _ = obj_vecs.shape
        E = pred_vecs.shape[1]
        H, P = self.hidden_dim, self.predicate_output_dim
        s_idx = edges[..., 0].long()
        o_idx = edges[..., 1].long()
        rows = torch.arange(B, device=obj_vecs.device).view(B, 1).expand(B, E)
        triples = torch.cat([obj_vecs[rows, s_idx], pred_vecs, obj_vecs[rows, o_idx]], dim=-1)
        hidden = self.net1(triples)
        new_s, new_p, new_o = hidden[..., :H], hidden[..., H:H + P], hidden[..., H + P:]

        pooled = []
        for b in range(B):
            keep = pred_indicators[b].bool()
            acc = torch.zeros(O, H, dtype=obj_vecs.dtype, device=obj_vecs.device)
            acc = acc.index_add(0, s_idx[b][keep], new_s[b][keep])
            acc = acc.index_add(0, o_idx[b][keep], new_o[b][keep])
            cnt = torch.zeros(O, dtype=obj_vecs.dtype, device=obj_vecs.device)
            one = torch.ones(int(keep.sum()), dtype=obj_vecs.dtype, device=obj_vecs.device)
            cnt = cnt.index_add(0, s_idx[b][keep], one).index_add(0, o_idx[b][keep], one)
            denom = torch.where(cnt > 0, cnt, torch.ones_like(cnt))
            pooled.append(acc / denom.view(O, 1))
        new_obj = self.net2(torch.stack(pooled, 0))
        return new_obj, (new_p if self.return_new_p_vecs else pred_vecs)


class GraphTripleConvNet(nn.Module):
    """The ``gconvs`` loop of models/graph_models/model.py:54-57,163-164 (and
    discriminator.py:246-250,308-309): layer after layer on the same edges."""

    def __init__(self, layers):
        super().__init__()
        self.gconvs = nn.ModuleList([GraphTripleConv(**kw) for kw in layers])

    def forward(self, obj_vecs, pred_vecs, edges, pred_indicators):
        for layer in self.gconvs:
            obj_vecs, pred_vecs = layer(obj_vecs, pred_vecs, edges, pred_indicators)
        return obj_vecs, pred_vecs


# --------------------------------------------------------------------------
# K2: layout composition
# --------------------------------------------------------------------------

def boxes_to_grid(boxes, H, W):
    """models/layout.py:98-130: per-object affine sampling grid, boxes are xywh."""
    O = boxes.shape[0]
    bx = boxes.reshape(O, 4, 1, 1)
    x0, y0, ww, hh = bx[:, 0], bx[:, 1], bx[:, 2], bx[:, 3]
    xs = torch.linspace(0, 1, steps=W).view(1, 1, W).to(boxes)
    ys = torch.linspace(0, 1, steps=H).view(1, H, 1).to(boxes)
    gx = ((xs - x0) / ww).expand(O, H, W)
    gy = ((ys - y0) / hh).expand(O, H, W)
    return torch.stack([gx, gy], dim=3).mul(2).sub(1)


def _sum_objects(sampled, pooling):
    """models/layout.py:205-237: scatter_add on dim 0 with an all-zero index, i.e.
    a sequential sum in object order; 'avg' divides by the object count."""
    O = sampled.shape[0]
    out = torch.zeros((1,) + tuple(sampled.shape[1:]), dtype=sampled.dtype, device=sampled.device)
    for o in range(O):
        out[0] = out[0] + sampled[o]
    if pooling == 'avg':
        out = out / float(max(O, 1))
    elif pooling != 'sum':
        raise ValueError('Invalid pooling "%s"' % pooling)
    return out


def boxes_to_layout(vecs, boxes, H, W=None, pooling='sum'):
    """models/layout.py:28-63."""
    legal = (boxes != 0).any(dim=-1)          # layout.py:40-42
    boxes, vecs = boxes[legal], vecs[legal]
    O, D = vecs.shape
    W = H if W is None else W
    if O == 0:
        if pooling not in ('sum', 'avg'):
            raise ValueError('Invalid pooling "%s"' % pooling)
        return torch.zeros(1, D, H, W, dtype=vecs.dtype, device=vecs.device)
    grid = boxes_to_grid(boxes, H, W)
    src = vecs.view(O, D, 1, 1).expand(O, D, 8, 8)   # layout.py:52
    sampled = F.grid_sample(src, grid, align_corners=True)
    return _sum_objects(sampled, pooling)


def masks_to_layout(vecs, boxes, masks, H, W=None, pooling='sum', test_mode=False):
    """models/layout.py:66-95 with _pool_mask_samples (layout.py:164-202).
    No zero-box filter here.  ``test_mode`` composites objects in ascending
    order of their sampled mass, first writer wins at threshold 0.5."""
    O, D = vecs.shape
    M = masks.shape[1]
    assert tuple(masks.shape) == (O, M, M)
    W = H if W is None else W
    grid = boxes_to_grid(boxes, H, W)
    src = vecs.view(O, D, 1, 1) * masks.float().view(O, 1, M, M)
    sampled = F.grid_sample(src, grid, align_corners=True)
    if not test_mode:
        out = _sum_objects(sampled, 'sum')
    else:
        clean = F.grid_sample(masks.float().view(O, 1, M, M), grid, align_corners=True)
        import numpy as np
        mass = [torch.sum(sampled[j]).item() for j in range(O)]
        order = list(np.argsort(mass))                       # layout.py:188-189
        painted = torch.zeros(H, W, dtype=sampled.dtype, device=sampled.device)
        canvas = torch.zeros(D, H, W, dtype=sampled.dtype, device=sampled.device)
        for j in order:
            take = (painted == 0).float() * (clean[j, 0] > 0.5).float()
            painted = painted + take
            canvas = canvas + sampled[j] * take
        out = canvas.unsqueeze(0)
    if pooling != 'sum':
        raise ValueError('Invalid pooling "%s"' % pooling)
    return out


def remove_dummy_objects(objs, vocab):
    """models/utils.py:95-102: keep rows that are neither padding (attr 0 == 0)
    nor the __image__ dummy."""
    image_id = vocab['object_name_to_idx']['__image__']
    first = objs[:, 0]
    return (first != 0) & (first != image_id)


def _sample_clamped(feats, X, Y):
    """models/bilinear.py:134-189 (bilinear_sample, the 'jj' backend): X, Y [N,HH,WW] in [0,1] scaled by the image
    size; the four taps are clamped into the image and weighted by their distance to the UNclamped coordinate;
    summed as w1*v1 + w2*v2 + w3*v3 + w4*v4 with v2 the tap below and v3 the tap to the right."""
    N, C, H, W = feats.shape
    X, Y = X * W, Y * H
    xa = X.floor().clamp(0, W - 1)
    xb = (xa + 1).clamp(0, W - 1)
    ya = Y.floor().clamp(0, H - 1)
    yb = (ya + 1).clamp(0, H - 1)
    n = torch.arange(N).view(N, 1, 1)

    def tap(yy, xx):                       # feats[n, :, yy, xx] -> [N,C,HH,WW]
        return feats[n, :, yy.long(), xx.long()].permute(0, 3, 1, 2)

    def wgt(a, b):
        return (a * b).unsqueeze(1)

    return (wgt(xb - X, yb - Y) * tap(ya, xa) + wgt(xb - X, Y - ya) * tap(yb, xa)
            + wgt(X - xa, yb - Y) * tap(ya, xb) + wgt(X - xa, Y - ya) * tap(yb, xb))


def crop_bbox(feats, bbox, HH, WW=None, backend='cudnn'):
    """models/bilinear.py:102-131 with tensor_linspace (bilinear.py:192-221) and xywh_to_points
    (models/metrics.py:20-24); backend 'cudnn' (grid_sample) or 'jj' (bilinear_sample)."""
    WW = HH if WW is None else WW
    N = feats.shape[0]
    pts = bbox.clone()
    pts[:, 2] = bbox[:, 0] + bbox[:, 2]
    pts[:, 3] = bbox[:, 1] + bbox[:, 3]
    if backend == 'cudnn':
        pts = 2 * pts - 1
    x0, y0, x1, y1 = pts[:, 0], pts[:, 1], pts[:, 2], pts[:, 3]

    def lerp(a, b, steps):
        wa = torch.linspace(1, 0, steps=steps).to(a).view(1, steps)
        wb = torch.linspace(0, 1, steps=steps).to(a).view(1, steps)
        return wa * a.view(N, 1) + wb * b.view(N, 1)

    X = lerp(x0, x1, WW).view(N, 1, WW).expand(N, HH, WW)
    Y = lerp(y0, y1, HH).view(N, HH, 1).expand(N, HH, WW)
    if backend == 'jj':
        return _sample_clamped(feats, X, Y)
    return F.grid_sample(feats, torch.stack([X, Y], dim=3), align_corners=True)


def crop_bbox_batch(imgs, objs, bbox, HH, WW=None, vocab=None, backend='cudnn'):
    """models/bilinear.py:29-44,67-99: per clip b and frame i, crop every real,
    non-zero box (object order); returns per-clip crops and flattened attribute rows."""
    assert backend == 'cudnn', 'only the live backend is restated'
    B, N, C, H, W = imgs.shape
    crops_b, objs_b = [], []
    for b in range(B):
        real = remove_dummy_objects(objs[b], vocab)
        frames, boxes, attrs = [], [], []
        for i in range(N):
            bb = bbox[b, i][real]
            legal = (bb != 0).any(dim=-1)
            bb = bb[legal]
            attrs.append(objs[b][real][legal].reshape(-1))
            frames.append(imgs[b, i].unsqueeze(0).expand(bb.shape[0], C, H, W))
            boxes.append(bb)
        crops_b.append(crop_bbox(torch.cat(frames, 0).contiguous(), torch.cat(boxes, 0), HH, WW))
        objs_b.append(torch.cat(attrs, 0))
    return crops_b, objs_b


# --------------------------------------------------------------------------
# K3: SPADE
# --------------------------------------------------------------------------

class _ParamFreeBatchNorm(nn.Module):
    """SynchronizedBatchNorm2d(affine=False) on one device = F.batch_norm
    (sync_batchnorm/batchnorm.py:63-68); buffers named as nn.BatchNorm2d's."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1):
        super().__init__()
        self.eps, self.momentum = eps, momentum
        self.register_buffer('running_mean', torch.zeros(num_features))
        self.register_buffer('running_var', torch.ones(num_features))
        self.register_buffer('num_batches_tracked', torch.tensor(0, dtype=torch.long))

    def forward(self, x):
        # the reference's forward override never bumps num_batches_tracked
        return F.batch_norm(x, self.running_mean, self.running_var, None, None,
                            self.training, self.momentum, self.eps)


class SPADE(nn.Module):
    """models/spade_models/networks/normalization.py:66-110."""

    def __init__(self, config_text, norm_nc, label_nc):
        super().__init__()
        assert config_text.startswith('spade')
        parsed = re.search(r'spade(\D+)(\d)x\d', config_text)
        kind, ks = str(parsed.group(1)), int(parsed.group(2))
        if kind not in ('syncbatch', 'batch'):
            raise ValueError('%s is not a recognized param-free norm type in SPADE' % kind)
        self.param_free_norm = _ParamFreeBatchNorm(norm_nc)
        pw = ks // 2
        self.mlp_shared = nn.Sequential(nn.Conv2d(label_nc, 128, kernel_size=ks, padding=pw), nn.ReLU())
        self.mlp_gamma = nn.Conv2d(128, norm_nc, kernel_size=ks, padding=pw)
        self.mlp_beta = nn.Conv2d(128, norm_nc, kernel_size=ks, padding=pw)

    def forward(self, x, segmap):
        normalized = self.param_free_norm(x)
        seg_r = F.interpolate(segmap, size=x.shape[2:], mode='nearest')
        actv = self.mlp_shared(seg_r)
        return normalized * (1 + self.mlp_gamma(actv)) + self.mlp_beta(actv)


class SPADEResnetBlock(nn.Module):
    """models/spade_models/networks/architecture.py:21-68."""

    def __init__(self, fin, fout, opt):
        super().__init__()
        self.learned_shortcut = fin != fout
        fmid = min(fin, fout)
        self.conv_0 = nn.Conv2d(fin, fmid, kernel_size=3, padding=1)
        self.conv_1 = nn.Conv2d(fmid, fout, kernel_size=3, padding=1)
        if self.learned_shortcut:
            self.conv_s = nn.Conv2d(fin, fout, kernel_size=1, bias=False)
        if 'spectral' in opt.norm_G:
            self.conv_0 = spectral_norm(self.conv_0)
            self.conv_1 = spectral_norm(self.conv_1)
            if self.learned_shortcut:
                self.conv_s = spectral_norm(self.conv_s)
        cfg = opt.norm_G.replace('spectral', '')
        self.norm_0 = SPADE(cfg, fin, opt.semantic_nc)
        self.norm_1 = SPADE(cfg, fmid, opt.semantic_nc)
        if self.learned_shortcut:
            self.norm_s = SPADE(cfg, fin, opt.semantic_nc)

    def forward(self, x, seg):
        x_s = self.conv_s(self.norm_s(x, seg)) if self.learned_shortcut else x
        dx = self.conv_0(F.leaky_relu(self.norm_0(x, seg), 0.2))
        dx = self.conv_1(F.leaky_relu(self.norm_1(dx, seg), 0.2))
        return x_s + dx

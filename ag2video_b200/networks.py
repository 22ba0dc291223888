"""The callers that drive the hot path, rebuilt around the sm_100a operators.

These modules mirror the reference's module tree (same attribute names and
state-dict keys, so reference checkpoints load after stripping DataParallel's
``module.`` infix) but are laid out for the kernels instead of translated:

* ``Acts2LayoutModel``  (models/graph_models/model.py:23-174): the per-timestep
  recurrence; each graph layer is one fused kernel.
* ``Layout2VidGenerator`` (models/spade_models/networks/generator.py:11-93): all
  B*F layouts of a batch come from ONE launch with a device-side object mask
  (no boolean-index compaction, no ``.item()`` sync); activations run in
  ``torch.channels_last`` so the SPADE kernels read them in place.
* ``SPADEGenerator`` (spade_generator.py:8-81): one ``SharedSeg`` per call.
* ``FlowsGenerator`` (flows_generator.py:13-109) and ``conv_dim_in`` are plain
  convolutions and stay on cuDNN (out of scope).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils import spectral_norm

from .graph import GraphTripleConv
from .layout import boxes_to_layout_batched, layout_conv3x3, layout_tables
from .spade import SPADEResnetBlock, SharedSeg, bn_act, plain_conv3x3, plain_conv3x3_usable
from .specnorm import SpectralNormGroup, conv_scaled, conv_unscaled
from .thinconv import thin_conv3x3

CL = torch.channels_last


class AttributeEmbeddings(nn.Module):
    """models/attribute_embed.py:16-46."""

    def __init__(self, attributes, embedding_dim):
        super().__init__()
        self.n_attr = len(attributes)
        if self.n_attr > 1:
            self.attribute_fc_gen = nn.Linear(self.n_attr * embedding_dim, self.n_attr * embedding_dim)
        for i, name in enumerate(list(attributes)):
            self.add_module('att_emb_%d' % i, nn.Embedding(max(attributes[name].values()) + 1, embedding_dim))

    def forward(self, objs):
        v = torch.cat([getattr(self, 'att_emb_%d' % k)(objs[..., k]) for k in range(objs.shape[-1])], dim=-1)
        return self.attribute_fc_gen(v) if self.n_attr > 1 else v


def real_object_mask(objs, vocab):
    """models/utils.py:95-102 for a whole batch: [B, O] bool, True for real objects."""
    first = objs[..., 0]
    return (first != 0) & (first != vocab['object_name_to_idx']['__image__'])


class Acts2LayoutModel(nn.Module):
    def __init__(self, opt):
        super().__init__()
        v = opt.vocab
        emb, gdim, hid = opt.embedding_dim, opt.gconv_dim, opt.gconv_hidden_dim
        n_attr = len(v['attributes'])
        obj_in = n_attr * emb
        self.vocab, self.embedding_dim = v, emb
        self.only_temporal = bool(getattr(opt, 'only_temporal', 0))
        self.pad_act = v['action_name_to_idx']['__padding__']
        self.pad_pred = v['pred_name_to_idx']['__padding__']
        self.attribute_embedding = AttributeEmbeddings(v['attributes'], emb)
        self.pred_embeddings = nn.Embedding(len(v['pred_idx_to_name']), emb)
        self.acts_embeddings = nn.Embedding(len(v['action_idx_to_name']), emb)
        first = dict(obj_input_dim=obj_in, object_output_dim=gdim, predicate_input_dim=emb, predicate_output_dim=gdim,
                     hidden_dim=hid, num_attributes=n_attr, mlp_normalization=opt.mlp_normalization,
                     pooling=opt.gconv_pooling, loc_dim=4)
        rest = dict(first, obj_input_dim=gdim, predicate_input_dim=gdim)
        self.gconvs = nn.ModuleList([GraphTripleConv(**(first if i == 0 else rest)) for i in range(opt.gconv_num_layers)])
        self.box_net = nn.Sequential(nn.Linear(gdim, hid), nn.ReLU(), nn.Linear(hid, 4))
        self.obj_vecs_net = nn.Sequential(nn.Linear(obj_in + 4, obj_in, bias=False), nn.ReLU(),
                                          nn.Linear(obj_in, obj_in, bias=False), nn.ReLU())

    def prepare(self, objs, triplets, actions):
        """Everything of model.py:108-158 that does not depend on the recurrence, for all timesteps at once:
        (emb [B,O,De], pred_vecs [B,T,E,Dp], edges [B,T,E,2], ind [B,T,E], actions_data)."""
        B, T = triplets.shape[:2]
        A = actions.shape[1]
        dev = actions.device
        act = actions.unsqueeze(1).expand(B, T, A, 7)
        f1, f2 = act[..., 3].float(), act[..., 4].float()
        steps = torch.arange(T, device=dev, dtype=torch.float32).view(1, T, 1)
        rel_t = (steps / T) * (f2 - f1 + 1e-6) + f1                                   # model.py:118
        a_id = torch.where((rel_t >= 0) & (rel_t <= 1), act[..., 1], act[..., 1].new_full((), float(self.pad_act)))
        temporal_triplets = torch.stack([act[..., 0], a_id, act[..., 2]], dim=-1).long()
        x_end, y_end = act[..., 5], act[..., 6]
        act_vecs = self.acts_embeddings(temporal_triplets[..., 1])
        act_vecs = torch.cat([act_vecs[..., :-3], x_end.unsqueeze(-1), y_end.unsqueeze(-1), rel_t.unsqueeze(-1)], dim=-1)
        edges = torch.stack([temporal_triplets[..., 0], temporal_triplets[..., 2]], dim=-1)   # (no list index: that is a host->device copy)
        ind = temporal_triplets[..., 1] != self.pad_act
        pred_vecs = act_vecs
        if not self.only_temporal:
            edges = torch.cat([torch.stack([triplets[..., 0], triplets[..., 2]], dim=-1), edges], dim=2)
            ind = torch.cat([triplets[..., 1] != self.pad_pred, ind], dim=2)
            pred_vecs = torch.cat([self.pred_embeddings(triplets[..., 1]), act_vecs], dim=2)
        emb = self.attribute_embedding(objs)
        locs = torch.stack([x_end, y_end], dim=-1)
        return emb, pred_vecs, edges, ind, [triplets, temporal_triplets, rel_t, locs]

    def recurrence_layerwise(self, emb, pred_vecs, edges, ind, box0):
        """The frame loop of model.py:126-169 with one fused kernel per graph layer (csrc/k1_gcn.cu): the general
        path, used when the persistent recurrence kernel does not cover the sizes."""
        T = pred_vecs.shape[1]
        # time-major and unbound once: every per-timestep operand is then a contiguous view (no copy per
        # layer call) and the backward of the T slices is ONE stack instead of T zero-filled select_backwards
        edges_t = edges.transpose(0, 1).contiguous().unbind(0)
        ind_t = ind.transpose(0, 1).contiguous().unbind(0)
        preds_t = pred_vecs.transpose(0, 1).contiguous().unbind(0)
        boxes = [box0]
        per_t = [emb.new_zeros(emb.shape[0], emb.shape[1], self.gconvs[-1].net2[2].weight.shape[0])]
        for t in range(1, T):
            obj_vecs = self.obj_vecs_net(torch.cat([emb, boxes[-1]], dim=-1))
            p_vecs = preds_t[t]
            for layer in self.gconvs:
                obj_vecs, p_vecs = layer(obj_vecs, p_vecs, edges_t[t], ind_t[t])
            per_t.append(obj_vecs)
            boxes.append(boxes[-1] + self.box_net(obj_vecs))                           # model.py:168
        return torch.stack(per_t, dim=1), torch.stack(boxes, dim=1)

    def forward(self, objs, triplets, actions, boxes_gt=None, test_mode=False):
        return acts2layout_forward([self], objs, triplets, actions, boxes_gt)[0]


def acts2layout_forward(models, objs, triplets, actions, boxes_gt, needs_grad=None):
    """``Acts2LayoutModel.forward`` for one or several models on the same clips (the generator step evaluates
    ``acts_to_boxes`` and ``acts_to_objs`` side by side): the whole recurrence of all of them is ONE launch of
    the persistent kernel (recurrence.run); sizes it does not cover walk the layers instead.
    ``needs_grad[i] = False`` evaluates model i without an autograd graph (its outputs are plain tensors)."""
    from . import recurrence
    needs_grad = [True] * len(models) if needs_grad is None else list(needs_grad)
    prep = []
    for m, g in zip(models, needs_grad):
        with torch.set_grad_enabled(g and torch.is_grad_enabled()):
            prep.append(m.prepare(objs, triplets, actions))
    box0 = boxes_gt[:, 0]
    edges, ind = prep[0][2], prep[0][3]
    res = None
    if all(isinstance(g, GraphTripleConv) for m in models for g in m.gconvs) and objs.is_cuda:
        res = recurrence.run(models, [p[0] for p in prep], box0, [p[1] for p in prep], edges, ind, needs_grad)
    if res is None:
        res = []
        for m, p, g in zip(models, prep, needs_grad):
            with torch.set_grad_enabled(g and torch.is_grad_enabled()):
                res.append(m.recurrence_layerwise(p[0], p[1], p[2], p[3], box0))
    return [(ov, bx, p[4]) for (ov, bx), p in zip(res, prep)]


class _BN2d(nn.Module):
    """SynchronizedBatchNorm2d(affine=True) (sync_batchnorm/batchnorm.py:63-68) with the
    LeakyReLU that follows it in every use folded in (``slope``); ``groups`` as in SPADE.forward.
    Statistics / normalisation / backward run on the K3 element-wise kernels (spade.bn_act)."""

    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer('running_mean', torch.zeros(c))
        self.register_buffer('running_var', torch.ones(c))
        self.register_buffer('num_batches_tracked', torch.tensor(0, dtype=torch.long))

    def forward(self, x, groups=1, slope=1.0, in_scale=None):
        return bn_act(x, self.weight, self.bias, self.running_mean, self.running_var, self.training, 0.1, 1e-5,
                      slope=slope, groups=groups, in_scale=in_scale)


def _sn_conv_bn(cin, cout, stride=1):
    return nn.Sequential(spectral_norm(nn.Conv2d(cin, cout, 3, stride=stride, padding=1, bias=False)), _BN2d(cout))


class _FlowResBlock(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv_0 = spectral_norm(nn.Conv2d(c, c, 3, padding=1))
        self.conv_1 = spectral_norm(nn.Conv2d(c, c, 3, padding=1))
        self.bn_0, self.bn_1 = _BN2d(c), _BN2d(c)

    def forward(self, x, groups=1):
        dx = conv_scaled(self.conv_0, self.bn_0(x, groups, 0.2))
        return x + conv_scaled(self.conv_1, self.bn_1(dx, groups, 0.2))


class FlowsGenerator(nn.Module):
    def __init__(self, opt):
        super().__init__()
        cin = opt.gconv_dim * 4 * opt.n_frames_G + (opt.n_frames_G - 1) * 3
        nf, nd = opt.nff, opt.n_downsample_F
        ch = [min(1024, nf * 2 ** i) for i in range(nd + 1)]
        down = [_sn_conv_bn(cin, nf), nn.LeakyReLU(0.2)]
        for i in range(nd):
            down += [_sn_conv_bn(ch[i], ch[i + 1], stride=2), nn.LeakyReLU(0.2)]
        up = []
        for i in reversed(range(nd)):
            up += [nn.Upsample(scale_factor=2), _sn_conv_bn(ch[i + 1], ch[i]), nn.LeakyReLU(0.2)]
        self.flow_multiplier = opt.flow_multiplier
        self.down_flow = nn.Sequential(*down)
        self.res_flow = nn.Sequential(*[_FlowResBlock(ch[-1]) for _ in range(opt.n_blocks_F)])
        self.up_flow = nn.Sequential(*up)
        self.conv_flow = nn.Sequential(nn.Conv2d(nf, 2, 3, padding=1))
        self.conv_w = nn.Sequential(nn.Conv2d(nf, 1, 3, padding=1), nn.Sigmoid())
        self.__dict__['_sn'] = SpectralNormGroup(self)

    @staticmethod
    def _stage(conv_bn, x, groups, first=None):
        """conv -> BN -> LeakyReLU(0.2); ``first`` = the convolution's output computed elsewhere."""
        if first is not None:
            return conv_bn[1](first[0], groups, 0.2, in_scale=first[1])
        z, scale = conv_unscaled(conv_bn[0], x)
        return conv_bn[1](z, groups, 0.2, in_scale=scale)

    def features(self, label, groups=1, first=None):
        """up_flow(res_flow(down_flow(label))).  ``first``: (output of down_flow[0]'s convolution,
        its pending 1/sigma scale or None) when the caller evaluated it itself (fused layout conv)."""
        x = self._stage(self.down_flow[0], label, groups, first)
        for i in range(2, len(self.down_flow), 2):
            x = self._stage(self.down_flow[i], x, groups)
        for block in self.res_flow:
            x = block(x, groups)
        for i in range(0, len(self.up_flow), 3):
            x = self._stage(self.up_flow[i + 1], self.up_flow[i](x), groups)
        return x

    def forward(self, label, groups=1):
        if groups == 1:
            self._sn.refresh_stale()
        feat = self.features(label, groups)
        return self.conv_w(feat), self.conv_flow(feat) * self.flow_multiplier


class SPADEGenerator(nn.Module):
    def __init__(self, opt):
        super().__init__()
        if opt.num_upsampling_layers != 'normal':
            raise NotImplementedError('num_upsampling_layers=%s' % opt.num_upsampling_layers)
        nf = opt.ngf
        self.sw = opt.image_size[0] // 32
        self.sh = round(self.sw / opt.aspect_ratio)
        self.fc = nn.Conv2d(opt.semantic_nc, 16 * nf, 3, padding=1)
        self.head_0 = SPADEResnetBlock(16 * nf, 16 * nf, opt)
        self.G_middle_0 = SPADEResnetBlock(16 * nf, 16 * nf, opt)
        self.G_middle_1 = SPADEResnetBlock(16 * nf, 16 * nf, opt)
        self.up_0 = SPADEResnetBlock(16 * nf, 8 * nf, opt)
        self.up_1 = SPADEResnetBlock(8 * nf, 4 * nf, opt)
        self.up_2 = SPADEResnetBlock(4 * nf, 2 * nf, opt)
        self.up_3 = SPADEResnetBlock(2 * nf, nf, opt)
        self.conv_img = nn.Conv2d(nf, 3, 3, padding=1)

    def forward(self, layout, groups=1):
        seg = SharedSeg.wrap(layout)                       # one NHWC copy + one gradient buffer for all 18 SPADEs
        up = lambda z: F.interpolate(z, scale_factor=2, mode='nearest')
        x = seg.nearest(self.sh, self.sw)                    # = F.interpolate(layout, size=(sh, sw))
        # fc on the tcgen05 implicit GEMM like the block convolutions (channels_last weights); the torch-layout
        # model of the tests (opt.channels_last = False) keeps the module call
        x = plain_conv3x3(self.fc, x) if plain_conv3x3_usable(self.fc, x) else self.fc(x)
        x = self.head_0(x, seg, groups)
        x = self.G_middle_0(up(x), seg, groups)
        x = self.G_middle_1(x, seg, groups)
        for name in ('up_0', 'up_1', 'up_2', 'up_3'):      # up(x) is read on the fly by the SPADE kernels
            x = getattr(self, name)(x, seg, groups, upsample=True)
        return thin_conv3x3(self.conv_img, x, slope_in=0.2, act_out='tanh')    # tanh(conv_img(leaky_relu(x, 0.2)))


_GRID = {}


def _base_grid(h, w, device):
    """get_grid of models/utils.py:127-140 (CPU linspace, as there), cached per device:
    a fresh host->device copy every call would drain the launch queue three times a step."""
    key = (h, w, str(device))
    if key not in _GRID:
        hor = torch.linspace(-1.0, 1.0, w).view(1, 1, 1, w).expand(1, 1, h, w)
        ver = torch.linspace(-1.0, 1.0, h).view(1, 1, h, 1).expand(1, 1, h, w)
        _GRID[key] = torch.cat([hor, ver], 1).to(device)
    return _GRID[key]


def flow_warp(image, flow):
    """models/utils.py:113-140 (border padding, align_corners=False)."""
    b, _, h, w = image.shape
    grid = _base_grid(h, w, image.device)
    flow = torch.cat([flow[:, 0:1] / ((w - 1.0) / 2.0), flow[:, 1:2] / ((h - 1.0) / 2.0)], dim=1)
    return F.grid_sample(image, (grid + flow).permute(0, 2, 3, 1), mode='bilinear', padding_mode='border',
                         align_corners=False)


class Layout2VidGenerator(nn.Module):
    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        self.attribute_embedding = AttributeEmbeddings(opt.vocab['attributes'], 384 // len(opt.vocab['attributes']))
        self.netG = SPADEGenerator(opt)
        self.flows_network = FlowsGenerator(opt)
        cin = opt.gconv_dim * 4 * opt.n_frames_G + 3
        self.conv_dim_in = nn.Sequential(_sn_conv_bn(cin, opt.semantic_nc), nn.LeakyReLU(0.2))
        self.channels_last = True
        # all 38 spectral-norm weights of a call in one launch (csrc/k5_specnorm.cu)
        self.__dict__['_sn'] = SpectralNormGroup(self)

    def build_layouts(self, objs, obj_vecs, boxes):
        """[B, F+1, D, H, H]: one launch for all (clip, frame) layouts (generator.py:36-54)."""
        B, T, O = boxes.shape[:3]
        H = self.opt.image_size[0]
        att = self.attribute_embedding(objs)
        vecs = torch.cat([att.unsqueeze(1).expand(B, T, O, att.shape[-1]), obj_vecs], dim=-1)
        valid = real_object_mask(objs, self.opt.vocab).unsqueeze(1).expand(B, T, O)
        seg = boxes_to_layout_batched(vecs.reshape(B * T, O, -1), boxes.reshape(B * T, O, 4),
                                      valid.reshape(B * T, O), H, H)
        seg = seg.view(B, T, -1, H, H)
        return torch.cat([seg, seg[:, -1:]], dim=1)

    def _generate(self, slots, tables, prev, groups):
        """One generator call on a batch of (previous frame, layout pair) items: flow network,
        warp, conv_dim_in, SPADE generator (generator.py:66-90).  The two convolutions that consume
        the layout run through its rank-1 structure (csrc/k4_layoutconv.cu): the [.,1024,H,W]
        layout pair, the 1027-channel concatenations and the dense 1027->32 / 1027->512
        convolutions never exist.  Returns (image, flow, confidence)."""
        D = slots[0].shape[-1]
        fn = self.flows_network
        flow_cb, in_cb = fn.down_flow[0], self.conv_dim_in[0]

        def layout_conv(conv_bn, img):
            entry = conv_bn[0].__dict__['_ag2v_sn_entry']
            if entry.scale is None:                      # weight / sigma was materialised (refresh)
                for hook in conv_bn[0]._forward_pre_hooks.values():
                    hook(conv_bn[0], None)
                w = conv_bn[0].weight
            else:                                        # sigma mode: raw weight, output scaled per image
                w = conv_bn[0].weight_orig
            base = F.conv2d(img, w[:, 2 * D:], padding=1).contiguous(memory_format=CL)
            return layout_conv3x3(w[:, :2 * D], slots, tables, base), entry.scale_g     # 1/sigma is folded into the BN

        feat = fn.features(None, groups, first=layout_conv(flow_cb, prev))
        flow = thin_conv3x3(fn.conv_flow[0], feat) * fn.flow_multiplier
        warped = flow_warp(prev[:, -3:], flow)
        diff = prev[:, -3:] - warped
        conf = ((diff * diff).sum(dim=1, keepdim=True) < 0.02).float()
        z, z_scale = layout_conv(in_cb, warped.contiguous(memory_format=CL))
        y = in_cb[1](z, groups, 0.2, in_scale=z_scale)
        return self.netG(y, groups) + warped, flow, conf

    def forward_fused(self, imgs_gt, objs, obj_vecs, layout, test_mode=False):
        """Same computation as ``forward`` on the fused kernels.  In training (ground-truth
        previous frames: not test_mode, bp_prev = 0) the T-1 generator calls of the reference's
        frame loop are independent given the inputs; they run as ONE call on a group-major batch
        of (T-1)*B images with per-group batch-norm statistics and per-group spectral-norm sigmas,
        i.e. exactly the arithmetic of T-1 successive calls."""
        B, T, O = layout.shape[:3]
        H = self.opt.image_size[0]
        n_prev = self.opt.n_frames_G - 1
        assert n_prev == 1, 'fused path is written for n_frames_G = 2 (the reference default)'
        att = self.attribute_embedding(objs)
        vecs = torch.cat([att.unsqueeze(1).expand(B, T, O, att.shape[-1]), obj_vecs], dim=-1)      # [B,T,O,512]
        valid = real_object_mask(objs, self.opt.vocab)                                               # [B,O]
        valid2 = torch.cat([valid, valid], dim=1)
        dev = imgs_gt.device
        sequential = test_mode or bool(self.opt.bp_prev) or not getattr(self, 'batch_frames', True)
        if not sequential and T > n_prev:
            G = T - n_prev                                                     # frame groups, n = g*B + b
            gm = lambda x: x.transpose(0, 1).reshape(G * B, *x.shape[2:])      # [B,G,...] -> group-major batch
            boxes2 = torch.cat([gm(layout[:, :T - 1]), gm(layout[:, 1:])], dim=1)
            tables = layout_tables(boxes2, valid2.repeat(G, 1), H, H)
            slots = [gm(vecs[:, :T - 1]), gm(vecs[:, 1:])]
            prev = gm(imgs_gt[:, :T - 1]).contiguous(memory_format=CL)
            self._sn.refresh_sigma(G, B)                                       # G power iterations, in frame order
            try:
                img, flow, cf = self._generate(slots, tables, prev, G)
            finally:
                self._sn.end_sigma()
            ug = lambda x: x.view(G, B, *x.shape[1:]).transpose(0, 1)          # back to [B,G,...]
            imgs = torch.cat([imgs_gt[:, :n_prev], ug(img)], dim=1)
            pad = lambda x: torch.cat([ug(x), torch.zeros(B, 1, *x.shape[1:], device=dev)], dim=1)
            return imgs, pad(flow), pad(cf)
        imgs_prev = imgs_gt[:, :n_prev]
        conf = torch.zeros(B, T, 1, H, H, device=dev)
        flows = torch.zeros(B, T, 2, H, H, device=dev)
        for t in range(n_prev, T):
            self._sn.refresh()             # one power iteration per generator call, like the reference's hooks
            # frame slots of seg_t: layout(t-1) -> channels [0, D), layout(t) -> channels [D, 2D)
            tables = layout_tables(torch.cat([layout[:, t - 1], layout[:, t]], dim=1), valid2, H, H)
            slots = [vecs[:, t - 1], vecs[:, t]]
            prev = imgs_prev[:, -n_prev:] if (test_mode or self.opt.bp_prev) else imgs_gt[:, t - n_prev:t]
            prev = prev.reshape(B, -1, H, H).contiguous(memory_format=CL)
            img, flow, cf = self._generate(slots, tables, prev, 1)
            conf[:, t - 1] = cf
            flows[:, t - 1] = flow
            imgs_prev = torch.cat([imgs_prev, img.unsqueeze(1)], dim=1)
        return imgs_prev, flows, conf

    def forward(self, imgs_gt, objs, obj_vecs, layout, imgs_prev=None, test_mode=False):
        if getattr(self, 'fuse_layout_conv', False) and self.channels_last:
            return self.forward_fused(imgs_gt, objs, obj_vecs, layout, test_mode=test_mode)
        seg = self.build_layouts(objs, obj_vecs, layout)
        n_prev = self.opt.n_frames_G - 1
        B, T = imgs_gt.shape[0], layout.shape[1]
        H = self.opt.image_size[0]
        imgs_prev = imgs_gt[:, :n_prev]
        conf = torch.zeros(B, T, 1, H, H, device=imgs_gt.device)
        flows = torch.zeros(B, T, 2, H, H, device=imgs_gt.device)
        for t in range(n_prev, T):
            self._sn.refresh()
            seg_t = seg[:, t - n_prev:t + 1].reshape(B, -1, H, H)
            prev = imgs_prev[:, -n_prev:] if (test_mode or self.opt.bp_prev) else imgs_gt[:, t - n_prev:t]
            prev = prev.reshape(B, -1, H, H)
            fmt = CL if self.channels_last else torch.contiguous_format
            weight, flow = self.flows_network(torch.cat([seg_t, prev], dim=1).contiguous(memory_format=fmt))
            warped = flow_warp(prev[:, -3:], flow)
            diff = prev[:, -3:] - warped
            conf[:, t - 1] = ((diff * diff).sum(dim=1, keepdim=True) < 0.02).float()
            flows[:, t - 1] = flow
            x = self.conv_dim_in(torch.cat([seg_t, warped], dim=1).contiguous(memory_format=fmt))
            img = self.netG(x) + warped
            imgs_prev = torch.cat([imgs_prev, img.unsqueeze(1)], dim=1)
        return imgs_prev, flows, conf


class AG2VideoModel(nn.Module):
    """models/meta_models.py:9-57; data parallelism is one process per GPU
    (ag2video_b200.dist) instead of the reference's in-process DataParallel."""

    def __init__(self, opt, device=None):
        super().__init__()
        self.acts_to_boxes = Acts2LayoutModel(opt)
        self.acts_to_objs = Acts2LayoutModel(opt)
        self.layout_to_video = Layout2VidGenerator(opt)
        if device is not None:
            self.to(device)
        # channels_last end to end: cuDNN then hands NHWC activations to the SPADE kernels
        # without a layout copy.  opt.channels_last = False keeps torch's default layout
        # (the SPADE ops convert internally); tests use it to share cuDNN algorithms with the
        # eager reference.
        self.channels_last = bool(getattr(opt, 'channels_last', True))
        self.layout_to_video.channels_last = self.channels_last
        # row f1: evaluate the layout's consumer convolutions through its rank-1 structure
        self.layout_to_video.fuse_layout_conv = bool(getattr(opt, 'fuse_layout_conv', True))
        if self.channels_last:
            self.to(memory_format=CL)

    def forward(self, imgs, objs, triplets, actions, boxes_gt=None, test_mode=False, use_gt=False, graph_only=False,
                boxes_pred_grad=True):
        """``boxes_pred_grad=False`` (not a reference argument): the returned ``boxes_pred`` carries no autograd graph.
        The training loop's generator step never differentiates it (train.py:446-459: its losses are image losses and
        ``optimizer_generator`` excludes ``acts_to_boxes``), and leaving ``acts_to_boxes`` out of that graph lets the
        graph step own its gradient accumulators on its own stream (Trainer.iteration)."""
        if graph_only:
            return self.acts_to_boxes(objs, triplets, actions, boxes_gt, test_mode)[1]
        # both graph models on the same clips: one launch of the recurrence kernel for the two of them
        (_, boxes_pred, _), (obj_vecs, _, actions_data) = acts2layout_forward(
            [self.acts_to_boxes, self.acts_to_objs], objs, triplets, actions, boxes_gt, [bool(boxes_pred_grad), True])
        boxes_in = boxes_gt if use_gt else boxes_pred.detach()
        imgs_pred, flows, conf = self.layout_to_video(imgs, objs, obj_vecs, boxes_in, test_mode=test_mode)
        return imgs_pred, boxes_pred, flows, conf, actions_data


def load_reference_state(model, state):
    """Load a reference checkpoint's ``model_state`` (scripts/train.py:528-543): drop
    the ``.module.`` infix of its DataParallel wrappers (meta_models.py:16-27)."""
    return model.load_state_dict({k.replace('.module.', '.'): v for k, v in state.items()}, strict=True)

"""Oracle restatement of the callers that fix the hot path's shapes (CPU, fp32).

TEST INFRASTRUCTURE — see ``oracle/__init__.py``.  These modules reproduce the
reference's module tree (same parameter / buffer names) so a reference
``state_dict`` loads after stripping the ``module.`` prefix that the
reference's DataParallel wrapper adds (models/meta_models.py:16-27).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils import spectral_norm

from . import ops


class AttributeEmbeddings(nn.Module):
    """models/attribute_embed.py:16-46: one table per attribute, concat, Linear."""

    def __init__(self, attributes, embedding_dim):
        super().__init__()
        n = len(attributes)
        if n > 1:
            self.attribute_fc_gen = nn.Linear(n * embedding_dim, n * embedding_dim)
        for i, name in enumerate(list(attributes)):
            self.add_module('att_emb_%d' % i,
                            nn.Embedding(max(attributes[name].values()) + 1, embedding_dim))
        self.n = n

    def forward(self, x):
        cols = [getattr(self, 'att_emb_%d' % k)(x[:, :, k]) for k in range(x.shape[-1])]
        out = torch.cat(cols, dim=-1)
        return self.attribute_fc_gen(out) if hasattr(self, 'attribute_fc_gen') else out


def _get(opt, name, default=None):
    return getattr(opt, name, default)


class Acts2LayoutModel(nn.Module):
    """models/graph_models/model.py:23-174 (mask_size == 0, the live setting)."""

    def __init__(self, opt):
        super().__init__()
        v = opt.vocab
        self.vocab = v
        emb, gdim, hid = opt.embedding_dim, opt.gconv_dim, opt.gconv_hidden_dim
        n_attr = len(v['attributes'])
        obj_in = n_attr * emb
        self.embedding_dim = emb
        self.only_temporal = bool(_get(opt, 'only_temporal', 0))
        self.attribute_embedding = AttributeEmbeddings(v['attributes'], emb)
        self.pred_embeddings = nn.Embedding(len(v['pred_idx_to_name']), emb)
        self.acts_embeddings = nn.Embedding(len(v['action_idx_to_name']), emb)
        first = dict(obj_input_dim=obj_in, object_output_dim=gdim, predicate_input_dim=emb,
                     predicate_output_dim=gdim, hidden_dim=hid, num_attributes=n_attr,
                     mlp_normalization=opt.mlp_normalization, pooling=opt.gconv_pooling, loc_dim=4)
        rest = dict(first, obj_input_dim=gdim, predicate_input_dim=gdim)
        self.gconvs = nn.ModuleList(
            [ops.GraphTripleConv(**(first if i == 0 else rest)) for i in range(opt.gconv_num_layers)])
        self.box_net = nn.Sequential(nn.Linear(gdim, hid), nn.ReLU(), nn.Linear(hid, 4))
        self.obj_vecs_net = nn.Sequential(nn.Linear(obj_in + 4, obj_in, bias=False), nn.ReLU(),
                                          nn.Linear(obj_in, obj_in, bias=False), nn.ReLU())

    def forward(self, objs, triplets, actions, boxes_gt=None, test_mode=False):
        B, T = triplets.shape[0], triplets.shape[1]
        pad_act = self.vocab['action_name_to_idx']['__padding__']
        pad_pred = self.vocab['pred_name_to_idx']['__padding__']
        act = actions.unsqueeze(1).expand(B, T, actions.shape[1], actions.shape[2])
        sa, a, oa, f1, f2, x_end, y_end = [act[..., k] for k in range(7)]
        t = torch.arange(T, dtype=torch.float32, device=actions.device).view(1, T, 1)
        rel_t = (t / T) * (f2.float() - f1.float() + 1e-6) + f1.float()        # model.py:118
        inside = (rel_t >= 0) & (rel_t <= 1)
        a = torch.where(inside, a, torch.full_like(a, float(pad_act)))          # model.py:119-121
        temporal_triplets = torch.stack([sa, a, oa], dim=-1).long()
        boxes_pred = [boxes_gt[:, 0]]
        emb = self.attribute_embedding(objs)
        per_t = [torch.zeros(objs.shape[0], objs.shape[1], self.embedding_dim, device=emb.device)]
        for ts in range(1, T):
            prev_boxes = boxes_pred[-1]
            obj_vecs = self.obj_vecs_net(torch.cat([emb, prev_boxes], dim=-1))
            at = temporal_triplets[:, ts]
            s_a, a_a, o_a = at[..., 0], at[..., 1], at[..., 2]
            act_vecs = self.acts_embeddings(a_a)
            act_vecs = torch.cat([act_vecs[..., :-3], x_end[:, ts].unsqueeze(-1),
                                  y_end[:, ts].unsqueeze(-1), rel_t[:, ts].unsqueeze(-1)], dim=-1)
            edges = torch.stack([s_a, o_a], dim=-1)
            ind = a_a != pad_act
            pred_vecs = act_vecs
            if not self.only_temporal:
                sp = triplets[:, ts]
                edges = torch.cat([torch.stack([sp[..., 0], sp[..., 2]], dim=-1), edges], dim=1)
                ind = torch.cat([sp[..., 1] != pad_pred, ind], dim=1)
                pred_vecs = torch.cat([self.pred_embeddings(sp[..., 1]), act_vecs], dim=1)
            for layer in self.gconvs:
                obj_vecs, pred_vecs = layer(obj_vecs, pred_vecs, edges, ind)
            per_t.append(obj_vecs)
            boxes_pred.append(prev_boxes + self.box_net(obj_vecs))
        boxes_pred = torch.stack(boxes_pred, dim=1)
        locs = torch.stack([x_end, y_end], dim=-1)
        return torch.stack(per_t, dim=1), boxes_pred, [triplets, temporal_triplets, rel_t, locs]


class _BN2d(nn.Module):
    """SynchronizedBatchNorm2d(affine=True) on one device
    (sync_batchnorm/batchnorm.py:63-68): F.batch_norm, no num_batches_tracked bump."""

    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer('running_mean', torch.zeros(c))
        self.register_buffer('running_var', torch.ones(c))
        self.register_buffer('num_batches_tracked', torch.tensor(0, dtype=torch.long))

    def forward(self, x):
        return F.batch_norm(x, self.running_mean, self.running_var, self.weight, self.bias,
                            self.training, 0.1, 1e-5)


def _sn_conv_bn(cin, cout, stride=1):
    """get_nonspade_norm_layer with 'spectralsyncbatch' (normalization.py:16-50):
    spectral-normed conv without bias followed by an affine (sync) batch norm."""
    return nn.Sequential(spectral_norm(nn.Conv2d(cin, cout, 3, stride=stride, padding=1, bias=False)),
                         _BN2d(cout))


class _FlowResBlock(nn.Module):
    """flows_generator.py:71-109 with norm='spectralsyncbatch' (plain BN, no SPADE)."""

    def __init__(self, c):
        super().__init__()
        self.conv_0 = spectral_norm(nn.Conv2d(c, c, 3, padding=1))
        self.conv_1 = spectral_norm(nn.Conv2d(c, c, 3, padding=1))
        self.bn_0, self.bn_1 = _BN2d(c), _BN2d(c)

    def forward(self, x):
        dx = self.conv_0(F.leaky_relu(self.bn_0(x), 0.2))
        dx = self.conv_1(F.leaky_relu(self.bn_1(dx), 0.2))
        return x + dx


class FlowsGenerator(nn.Module):
    """flows_generator.py:13-68 (flow_deconv off)."""

    def __init__(self, opt):
        super().__init__()
        n_prev = opt.n_frames_G - 1
        cin = opt.gconv_dim * 4 * opt.n_frames_G + n_prev * 3
        nf, nd = opt.nff, opt.n_downsample_F
        ch = [min(1024, nf * 2 ** i) for i in range(nd + 1)]
        act = lambda: nn.LeakyReLU(0.2)
        down = [_sn_conv_bn(cin, nf), act()]
        for i in range(nd):
            down += [_sn_conv_bn(ch[i], ch[i + 1], stride=2), act()]
        up = []
        for i in reversed(range(nd)):
            up += [nn.Upsample(scale_factor=2), _sn_conv_bn(ch[i + 1], ch[i]), act()]
        self.flow_multiplier = opt.flow_multiplier
        self.down_flow = nn.Sequential(*down)
        self.res_flow = nn.Sequential(*[_FlowResBlock(ch[-1]) for _ in range(opt.n_blocks_F)])
        self.up_flow = nn.Sequential(*up)
        self.conv_flow = nn.Sequential(nn.Conv2d(nf, 2, 3, padding=1))
        self.conv_w = nn.Sequential(nn.Conv2d(nf, 1, 3, padding=1), nn.Sigmoid())

    def forward(self, label):
        feat = self.up_flow(self.res_flow(self.down_flow(label)))
        return self.conv_w(feat), self.conv_flow(feat) * self.flow_multiplier


class SPADEGenerator(nn.Module):
    """spade_generator.py:8-81 ('normal' number of upsampling layers)."""

    def __init__(self, opt):
        super().__init__()
        assert opt.num_upsampling_layers == 'normal'
        nf = opt.ngf
        self.sw = opt.image_size[0] // 32
        self.sh = round(self.sw / opt.aspect_ratio)
        self.fc = nn.Conv2d(opt.semantic_nc, 16 * nf, 3, padding=1)
        blk = ops.SPADEResnetBlock
        self.head_0 = blk(16 * nf, 16 * nf, opt)
        self.G_middle_0 = blk(16 * nf, 16 * nf, opt)
        self.G_middle_1 = blk(16 * nf, 16 * nf, opt)
        self.up_0 = blk(16 * nf, 8 * nf, opt)
        self.up_1 = blk(8 * nf, 4 * nf, opt)
        self.up_2 = blk(4 * nf, 2 * nf, opt)
        self.up_3 = blk(2 * nf, nf, opt)
        self.conv_img = nn.Conv2d(nf, 3, 3, padding=1)

    def forward(self, layout):
        up = lambda z: F.interpolate(z, scale_factor=2, mode='nearest')
        x = self.fc(F.interpolate(layout, size=(self.sh, self.sw)))
        x = self.head_0(x, layout)
        x = self.G_middle_0(up(x), layout)
        x = self.G_middle_1(x, layout)
        for name in ('up_0', 'up_1', 'up_2', 'up_3'):
            x = getattr(self, name)(up(x), layout)
        return torch.tanh(self.conv_img(F.leaky_relu(x, 0.2)))


def flow_warp(image, flow):
    """models/utils.py:113-140: border padding, align_corners=False."""
    b, _, h, w = image.shape
    hor = torch.linspace(-1.0, 1.0, w).to(image.device).view(1, 1, 1, w).expand(b, 1, h, w)
    ver = torch.linspace(-1.0, 1.0, h).to(image.device).view(1, 1, h, 1).expand(b, 1, h, w)
    grid = torch.cat([hor, ver], 1)
    flow = torch.cat([flow[:, 0:1] / ((w - 1.0) / 2.0), flow[:, 1:2] / ((h - 1.0) / 2.0)], dim=1)
    return F.grid_sample(image, (grid + flow).permute(0, 2, 3, 1), mode='bilinear',
                         padding_mode='border', align_corners=False)


class Layout2VidGenerator(nn.Module):
    """models/spade_models/networks/generator.py:11-93."""

    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        self.attribute_embedding = AttributeEmbeddings(opt.vocab['attributes'],
                                                       384 // len(opt.vocab['attributes']))
        self.netG = SPADEGenerator(opt)
        self.flows_network = FlowsGenerator(opt)
        cin = opt.gconv_dim * 4 * opt.n_frames_G + 3
        self.conv_dim_in = nn.Sequential(_sn_conv_bn(cin, opt.semantic_nc), nn.LeakyReLU(0.2))

    def build_layouts(self, objs, obj_vecs, boxes):
        """generator.py:36-54: one layout per (clip, frame), plus a repeat of the last."""
        att = self.attribute_embedding(objs)
        H = self.opt.image_size[0]
        clips = []
        for b in range(obj_vecs.shape[0]):
            real = ops.remove_dummy_objects(objs[b], self.opt.vocab)
            frames = []
            for t in range(boxes.shape[1]):
                vecs = torch.cat([att[b][real], obj_vecs[b, t][real]], dim=1)
                frames.append(ops.boxes_to_layout(vecs, boxes[b, t][real], H, H))
            frames.append(frames[-1])
            clips.append(torch.cat(frames, dim=0))
        return torch.stack(clips, dim=0)

    def forward(self, imgs_gt, objs, obj_vecs, layout, imgs_prev=None, test_mode=False):
        seg = self.build_layouts(objs, obj_vecs, layout)
        n_prev = self.opt.n_frames_G - 1
        B, T = imgs_gt.shape[0], layout.shape[1]
        H = self.opt.image_size[0]
        imgs_prev = imgs_gt[:, :n_prev]
        conf = torch.zeros(B, T, 1, H, H, device=imgs_gt.device)
        flows = torch.zeros(B, T, 2, H, H, device=imgs_gt.device)
        for t in range(n_prev, T):
            seg_t = seg[:, t - n_prev:t + 1].reshape(B, -1, H, H)
            if test_mode or self.opt.bp_prev:
                prev = imgs_prev[:, -n_prev:]
            else:
                prev = imgs_gt[:, t - n_prev:t]
            prev = prev.reshape(B, -1, H, H)
            weight, flow = self.flows_network(torch.cat([seg_t, prev], dim=1))
            warped = flow_warp(prev[:, -3:], flow)
            diff = prev[:, -3:] - warped
            conf[:, t - 1] = ((diff * diff).sum(dim=1, keepdim=True) < 0.02).float()
            flows[:, t - 1] = flow
            x = self.conv_dim_in(torch.cat([seg_t, warped], dim=1))
            img = self.netG(x) + warped
            imgs_prev = torch.cat([imgs_prev, img.unsqueeze(1)], dim=1)
        return imgs_prev, flows, conf


class AG2VideoModel(nn.Module):
    """models/meta_models.py:9-57 without the DataParallel wrappers."""

    def __init__(self, opt):
        super().__init__()
        self.acts_to_boxes = Acts2LayoutModel(opt)
        self.acts_to_objs = Acts2LayoutModel(opt)
        self.layout_to_video = Layout2VidGenerator(opt)

    def forward(self, imgs, objs, triplets, actions, boxes_gt=None, test_mode=False, use_gt=False,
                graph_only=False):
        _, boxes_pred, actions_data = self.acts_to_boxes(objs, triplets, actions, boxes_gt, test_mode)
        if graph_only:
            return boxes_pred
        obj_vecs, _, actions_data = self.acts_to_objs(objs, triplets, actions, boxes_gt, test_mode)
        boxes_in = boxes_gt if use_gt else boxes_pred.detach()
        imgs_pred, flows, conf = self.layout_to_video(imgs, objs, obj_vecs, boxes_in, test_mode=test_mode)
        return imgs_pred, boxes_pred, flows, conf, actions_data


def strip_module_prefix(state):
    """Reference checkpoints carry '<sub>.module.' from DataParallel (meta_models.py:16-27)."""
    return {k.replace('.module.', '.'): v for k, v in state.items()}

"""Discriminator-side callers of the hot path (SURVEY.md section 8, rows ctx / f4).

``MultiscaleActionDiscriminator`` mirrors spade_models/networks/discriminator.py:212-322
(same module tree and state-dict keys).  What belongs to the path runs on the
sm_100a kernels:

* its 2-layer action-graph network (``GraphTripleConv``, K1) per frame,
* its 256-channel layouts are never built: the first PatchGAN convolution of every scale
  (4x4, stride 2, 259 -> 64) is evaluated through the layout's rank-1 structure
  (``layout.layout_sconv``, K7), the coarser scales from the average-pooled separable masks.
  ``rank1_stem = False`` keeps the dense form: K2 writes all (clip, frame) layouts in ONE
  launch straight into the channel slice of the ``cat([img, seg])`` buffer
  (``layout.layout_cat``) - the reference materialises the layout, then copies it.

The conditioning (graph vectors -> fc -> per-object layout vectors) depends only on
the discriminator's parameters and on data, not on the image, so one evaluation
(``condition``) is shared by the fake and the real pass of a loss; the reference
recomputes it per pass (discriminator.py:317-337).  The PatchGAN stacks
(``NLayerActionDiscriminator``, :326-372) are plain 4x4 library convolutions with
torch's spectral norm and instance norm and stay on cuDNN.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils import spectral_norm

from .graph import GraphTripleConv
from .layout import layout_cat, layout_sconv, layout_tables, layout_tables_avgpool, pooled_size
from .networks import AttributeEmbeddings, real_object_mask


class NLayerActionDiscriminator(nn.Module):
    """discriminator.py:326-372 with norm_D = 'spectral' + 'instance' (args.py:105)."""

    def __init__(self, opt):
        super().__init__()
        norm = getattr(opt, 'norm_D', 'spectralinstance')
        if norm != 'spectralinstance':
            raise NotImplementedError('norm_D=%r: only the reference default "spectralinstance" is built' % norm)
        kw, padw, nf = 4, 2, opt.ndf
        input_nc = (opt.gconv_dim if getattr(opt, 'use_actions_loss', 1) else opt.semantic_nc) * 2 + 3
        seq = [[nn.Conv2d(input_nc, nf, kw, stride=2, padding=padw), nn.LeakyReLU(0.2, False)]]
        for n in range(1, opt.n_layers_D):
            prev, nf = nf, min(nf * 2, 512)
            stride = 1 if n == opt.n_layers_D - 1 else 2
            conv = spectral_norm(nn.Conv2d(prev, nf, kw, stride=stride, padding=padw, bias=False))
            seq.append([nn.Sequential(conv, nn.InstanceNorm2d(nf, affine=False)), nn.LeakyReLU(0.2, False)])
        seq.append([nn.Conv2d(nf, 1, kw, stride=1, padding=padw)])
        for n, layers in enumerate(seq):
            self.add_module('model%d' % n, nn.Sequential(*layers))

    @staticmethod
    def _level(sub, x):
        """One level of the trunk.  [Sequential(conv, InstanceNorm2d), LeakyReLU(0.2)] (discriminator.py:373-380) runs as
        conv (library, NHWC) -> instance norm + LeakyReLU on the grouped batch-norm kernels (one group per image)."""
        first = sub[0]
        if (x.is_cuda and isinstance(first, nn.Sequential) and len(first) == 2 and isinstance(first[1], nn.InstanceNorm2d)
                and not first[1].affine and not first[1].track_running_stats and len(sub) == 2
                and isinstance(sub[1], nn.LeakyReLU) and first[0].out_channels % 4 == 0):
            from .spade import instance_norm_act
            return instance_norm_act(first[0](x), first[1].eps, sub[1].negative_slope)
        return sub(x)

    def forward(self, x):
        outs = []
        for sub in self.children():
            x = self._level(sub, x)
            outs.append(x)
        return outs

    def from_stem(self, y):
        """Levels 1.. given the output of model0 (the stem evaluated elsewhere)."""
        outs = [y]
        for sub in list(self.children())[1:]:
            y = self._level(sub, y)
            outs.append(y)
        return outs


class MultiscaleActionDiscriminator(nn.Module):
    def __init__(self, opt):
        super().__init__()
        v = opt.vocab
        self.vocab = v
        self.image_size = opt.image_size[0]
        emb, gdim, hid = opt.embedding_dim, opt.gconv_dim, opt.gconv_hidden_dim
        n_attr = len(v['attributes'])
        obj_in = n_attr * emb
        self.pad_act = v['action_name_to_idx']['__padding__']
        self.num_D = opt.num_D
        # evaluate the PatchGAN stems through the layout's rank-1 structure (K7); False = materialise the
        # layout inside the cat([img, seg]) buffer (K2) and run the dense library convolution
        self.rank1_stem = bool(getattr(opt, 'rank1_stem', True)) and opt.ndf == 64
        for i in range(opt.num_D):
            self.add_module('discriminator_%d' % i, NLayerActionDiscriminator(opt))
        self.attribute_embedding = AttributeEmbeddings(v['attributes'], emb)
        self.pred_embeddings = nn.Embedding(len(v['pred_idx_to_name']), emb)       # in the state dict, never read
        self.acts_embeddings = nn.Embedding(len(v['action_idx_to_name']), emb)
        first = dict(obj_input_dim=obj_in, object_output_dim=gdim, predicate_input_dim=emb, predicate_output_dim=gdim,
                     hidden_dim=hid, num_attributes=n_attr, mlp_normalization=opt.mlp_normalization,
                     pooling=opt.gconv_pooling, loc_dim=4)
        rest = dict(first, obj_input_dim=gdim, predicate_input_dim=gdim)
        self.gconvs = nn.ModuleList([GraphTripleConv(**first), GraphTripleConv(**rest)])
        self.obj_vecs_net = nn.Sequential(nn.Linear(emb + 4, obj_in, bias=False), nn.ReLU(),
                                          nn.Linear(obj_in, obj_in, bias=False), nn.ReLU())
        self.pre_obj_vecs_net = nn.Sequential(nn.Linear(obj_in, emb, bias=False), nn.ReLU(),
                                              nn.Linear(emb, emb, bias=False), nn.ReLU())
        self.fc_objs_vecs = nn.Linear(gdim + opt.semantic_nc, gdim * 2)
        # the spectral norms of the PatchGAN trunks (torch's pre-forward hooks: ~10 tiny launches per module and call)
        # run as ONE launch for all trunks per forward call, with the hooks' per-call semantics (csrc/k5_specnorm.cu)
        from .specnorm import SpectralNormGroup
        self.__dict__['_sn'] = SpectralNormGroup(self)

    def get_obj_vecs(self, objs, layout_boxes, actions_data):
        """discriminator.py:273-313: [B,T,O,gconv_dim]; the object vectors are carried across frames.
        Everything outside the recurrence (embeddings, edges, indicators) is built for all frames at once."""
        _, temporal_triplets, rel_t, locs = actions_data
        T = layout_boxes.shape[1]
        a = temporal_triplets[..., 1].long()
        act_vecs = self.acts_embeddings(a)
        act_vecs = torch.cat([act_vecs[..., :-3], locs[..., 0:1], locs[..., 1:2], rel_t.unsqueeze(-1)], dim=-1)
        edges = torch.stack([temporal_triplets[..., 0], temporal_triplets[..., 2]], dim=-1).long()
        # time-major, unbound once: contiguous per-frame views, one stack in the backward (see Acts2LayoutModel)
        edges_t = edges.transpose(0, 1).contiguous().unbind(0)
        ind_t = (a != self.pad_act).transpose(0, 1).contiguous().unbind(0)
        acts_t = act_vecs.transpose(0, 1).contiguous().unbind(0)
        obj_vecs = self.pre_obj_vecs_net(self.attribute_embedding(objs))
        per_t = []
        for t in range(T):
            obj_vecs = self.obj_vecs_net(torch.cat([obj_vecs, layout_boxes[:, t]], dim=-1))
            p_vecs = acts_t[t]
            for layer in self.gconvs:
                obj_vecs, p_vecs = layer(obj_vecs, p_vecs, edges_t[t], ind_t[t])
            per_t.append(obj_vecs)
        return torch.stack(per_t, dim=1)

    def condition(self, objs, layout_boxes, actions_data):
        """Everything the PatchGAN stems need from the graph side, for every (clip, frame):
        (vecs [B*T,O,2*gconv_dim], boxes [B*T,O,4], valid [B*T,O], tables per scale) -
        discriminator.py:317-331 without the boolean-index compaction (the object mask travels to the
        kernel instead).  ``tables`` = the separable mask tables at every scale (K2 tables, then their
        3x3/stride-2 average pools): data only, shared by all passes."""
        B, T, O = layout_boxes.shape[:3]
        obj_vecs = self.get_obj_vecs(objs, layout_boxes, actions_data)
        att = self.attribute_embedding(objs)
        vecs = self.fc_objs_vecs(torch.cat([att.unsqueeze(1).expand(B, T, O, att.shape[-1]), obj_vecs], dim=-1))
        valid = real_object_mask(objs, self.vocab).unsqueeze(1).expand(B, T, O)
        vecs, boxes, valid = vecs.reshape(B * T, O, -1), layout_boxes.reshape(B * T, O, 4), valid.reshape(B * T, O)
        tables = None
        if self.rank1_stem:
            H = self.image_size
            tables = [layout_tables(boxes, valid, H, H)]
            for _ in range(1, self.num_D):
                tables.append(layout_tables_avgpool(tables[-1], B * T, O, H, H))
                H = pooled_size(H)
        return vecs, boxes, valid, tables

    @staticmethod
    def _pool(x):
        return F.avg_pool2d(x, kernel_size=3, stride=2, padding=[1, 1], count_include_pad=False)

    def forward(self, img, objs, layout_boxes, actions_data, cond=None):
        """List (scale) of lists (level) of PatchGAN outputs, discriminator.py:317-353.
        ``cond`` = a ``condition(...)`` result to reuse (same objs / boxes / actions)."""
        if cond is None:
            cond = self.condition(objs, layout_boxes, actions_data)
        vecs, boxes, valid, tables = cond
        H = self.image_size
        img = img.reshape(-1, *img.shape[2:])
        if img.is_cuda:
            self._sn.refresh()             # one power iteration per trunk convolution and call, like the hooks
        nets = [D for name, D in self.named_children() if name.startswith('discriminator')]
        result = []
        if not self.rank1_stem:
            # the layout is written straight into the channel slice of cat([img, seg]) (K2, batch-strided)
            x = layout_cat(img, vecs, boxes, valid, H, H)
            for i, D in enumerate(nets):
                result.append(D(x))
                if i + 1 < len(nets):       # (the reference also pools after the last scale and drops it)
                    x = self._pool(x)
            return result
        # rank-1 stem (csrc/k7_layoutconv_strided.cu): conv(cat[img, seg]) = conv(img; W[:, :3]) + b + the
        # layout part evaluated from its separable masks; the pooled scales use pooled images / pooled masks
        for i, D in enumerate(nets):
            conv = D.model0[0]
            w = conv.weight
            base = F.conv2d(img, w[:, :3], conv.bias, stride=conv.stride, padding=conv.padding)
            base = base.contiguous(memory_format=torch.channels_last)
            y = layout_sconv(w[:, 3:], vecs, tables[i], base, H, H, stride=conv.stride[0], pad=conv.padding[0])
            result.append(D.from_stem(D.model0[1](y)))
            if i + 1 < len(nets):
                img, H = self._pool(img), pooled_size(H)
        return result


class MetaDiscriminatorModel(nn.Module):
    """models/meta_models.py:60-72: the image discriminator and its Adam optimiser."""

    def __init__(self, opt, device=None, fused=None):
        super().__init__()
        self.img_discriminator = MultiscaleActionDiscriminator(opt)
        if device is not None:
            self.img_discriminator.to(device)
        if bool(getattr(opt, 'channels_last', True)):
            # the PatchGAN stems hand NHWC activations on (K7); NHWC weights keep cuDNN from converting the trunk's
            # activations and weights back and forth (144 nchwToNhwc / nhwcToNchw launches per iteration in round 1)
            # (the stems keep torch's layout: their 3 image channels go through a library convolution that is
            # faster on NCHW weights, the layout channels through K7)
            for name, D in self.img_discriminator.named_children():
                if name.startswith('discriminator'):
                    for sub in list(D.children())[1:]:
                        sub.to(memory_format=torch.channels_last)
        self.img_discriminator.train()
        cuda = next(self.img_discriminator.parameters()).is_cuda
        fused = cuda if fused is None else fused
        self.optimizer_d_img = torch.optim.Adam(list(self.img_discriminator.parameters()), lr=opt.learning_rate,
                                                betas=(opt.beta1, 0.999), **(dict(fused=True, capturable=True) if fused else {}))

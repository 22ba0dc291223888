"""Oracle restatement of the discriminator-side callers of the hot path and of the
three losses of one training iteration (CPU, fp32).

TEST INFRASTRUCTURE — see ``oracle/__init__.py``.

``MultiscaleActionDiscriminator`` (spade_models/networks/discriminator.py:212-322)
calls ``GraphTripleConv`` (2 layers, action edges only) once per frame and
``boxes_to_layout`` (D = 2*gconv_dim = 256 channels) once per (clip, frame); its
PatchGAN stacks (``NLayerActionDiscriminator``, discriminator.py:326-372) are
plain 4x4 library convolutions.  ``LossModel``
(spade_models/loss_model.py:13-149) turns its outputs into the generator,
discriminator and graph losses; scripts/train.py:440-493 fixes their order.
Module / parameter names follow the reference so its ``state_dict`` loads.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils import spectral_norm

from . import ops
from .networks import AttributeEmbeddings, flow_warp


class NLayerActionDiscriminator(nn.Module):
    """discriminator.py:326-372 with norm_D = 'spectralinstance' (args.py:105):
    conv4x4 s2 + LeakyReLU | (SN conv4x4 (no bias) + InstanceNorm) + LeakyReLU x3 | conv4x4 -> 1."""

    def __init__(self, opt):
        super().__init__()
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

    def forward(self, x):
        outs = []
        for sub in self.children():
            x = sub(x)
            outs.append(x)
        return outs


class MultiscaleActionDiscriminator(nn.Module):
    """discriminator.py:212-322."""

    def __init__(self, opt):
        super().__init__()
        v = opt.vocab
        self.vocab = v
        self.image_size = opt.image_size[0]
        emb, gdim, hid = opt.embedding_dim, opt.gconv_dim, opt.gconv_hidden_dim
        n_attr = len(v['attributes'])
        obj_in = n_attr * emb
        for i in range(opt.num_D):
            self.add_module('discriminator_%d' % i, NLayerActionDiscriminator(opt))
        self.attribute_embedding = AttributeEmbeddings(v['attributes'], emb)
        self.pred_embeddings = nn.Embedding(len(v['pred_idx_to_name']), emb)        # never read (:234)
        self.acts_embeddings = nn.Embedding(len(v['action_idx_to_name']), emb)
        first = dict(obj_input_dim=obj_in, object_output_dim=gdim, predicate_input_dim=emb,
                     predicate_output_dim=gdim, hidden_dim=hid, num_attributes=n_attr,
                     mlp_normalization=opt.mlp_normalization, pooling=opt.gconv_pooling, loc_dim=4)
        rest = dict(first, obj_input_dim=gdim, predicate_input_dim=gdim)
        self.gconvs = nn.ModuleList([ops.GraphTripleConv(**first), ops.GraphTripleConv(**rest)])
        self.obj_vecs_net = nn.Sequential(nn.Linear(emb + 4, obj_in, bias=False), nn.ReLU(),
                                          nn.Linear(obj_in, obj_in, bias=False), nn.ReLU())
        self.pre_obj_vecs_net = nn.Sequential(nn.Linear(obj_in, emb, bias=False), nn.ReLU(),
                                              nn.Linear(emb, emb, bias=False), nn.ReLU())
        self.fc_objs_vecs = nn.Linear(gdim + opt.semantic_nc, gdim * 2)

    def get_obj_vecs(self, objs, boxes, actions_data):
        """:273-313 — the object vectors are carried from one frame to the next."""
        _, temporal_triplets, rel_t, locs = actions_data
        pad_act = self.vocab['action_name_to_idx']['__padding__']
        obj_vecs = self.pre_obj_vecs_net(self.attribute_embedding(objs))
        per_t = []
        for t in range(boxes.shape[1]):
            obj_vecs = self.obj_vecs_net(torch.cat([obj_vecs, boxes[:, t]], dim=-1))
            at = temporal_triplets[:, t]
            s, a, o = at[..., 0], at[..., 1], at[..., 2]
            act_vecs = self.acts_embeddings(a.long())
            act_vecs = torch.cat([act_vecs[..., :-3], locs[:, t, :, 0:1], locs[:, t, :, 1:2],
                                  rel_t[:, t].unsqueeze(-1)], dim=-1)
            edges, ind, pred_vecs = torch.stack([s, o], dim=-1).long(), a != pad_act, act_vecs
            for layer in self.gconvs:
                obj_vecs, pred_vecs = layer(obj_vecs, pred_vecs, edges, ind)
            per_t.append(obj_vecs)
        return torch.stack(per_t, dim=1)

    def build_seg(self, objs, boxes, actions_data):
        """:317-337 — per (clip, frame) layout of fc([attribute embedding, graph vector])."""
        obj_vecs = self.get_obj_vecs(objs, boxes, actions_data)
        att = self.attribute_embedding(objs)
        clips = []
        for b in range(objs.shape[0]):
            keep = ops.remove_dummy_objects(objs[b], self.vocab)
            frames = []
            for t in range(boxes.shape[1]):
                vecs = self.fc_objs_vecs(torch.cat([att[b][keep], obj_vecs[b, t][keep]], dim=1))
                frames.append(ops.boxes_to_layout(vecs, boxes[b, t][keep], self.image_size, self.image_size))
            clips.append(torch.cat(frames, dim=0))
        return torch.stack(clips, dim=0)

    def forward(self, img, objs, boxes, actions_data):
        seg = self.build_seg(objs, boxes, actions_data)
        x = torch.cat([img, seg], dim=2)
        x = x.reshape(-1, *x.shape[2:])
        result = []
        for name, D in self.named_children():
            if name.startswith('discriminator'):
                result.append(D(x))
                x = F.avg_pool2d(x, kernel_size=3, stride=2, padding=[1, 1], count_include_pad=False)
        return result


def hinge_loss(preds, target_is_real, for_discriminator):
    """GANLoss(gan_mode='hinge') over the multiscale list (networks/loss.py:61-97):
    the last output of every scale, averaged over scales."""
    total = 0
    for scale in preds:
        x = scale[-1]
        if not for_discriminator:
            l = -torch.mean(x)
        elif target_is_real:
            l = -torch.mean(torch.min(x - 1, torch.zeros_like(x)))
        else:
            l = -torch.mean(torch.min(-x - 1, torch.zeros_like(x)))
        total = total + l
    return total / len(preds)


class LossModel(nn.Module):
    """spade_models/loss_model.py:13-149 with --no_vgg_loss (VGG19 weights are a download)."""

    def __init__(self, opt, netD_img):
        super().__init__()
        self.opt = opt
        self.netD_img = netD_img

    def _relevant(self, batch, model_out):
        imgs_pred, _, _, _, actions_data = model_out
        n = self.opt.n_frames_G - 1
        return (batch['imgs'][:, n:], batch['boxes'][:, n:], imgs_pred[:, n:], [a[:, n:] for a in actions_data])

    def compute_graph_loss(self, batch, boxes_pred):
        """:40-60."""
        objs, boxes = batch['objs'], batch['boxes']
        w = getattr(self.opt, 'bbox_pred_loss_weight', 10)
        l = F.smooth_l1_loss(boxes_pred[:, 1:].reshape(-1, 4), boxes[:, 1:].reshape(-1, 4), reduction='none') * w
        flat = objs.unsqueeze(1).repeat(1, boxes.shape[1] - 1, 1, 1).view(-1, objs.shape[-1])
        real = (flat.sum(1) != 0).float().unsqueeze(1)
        out = {'bbox_pred': (l * real).mean()}
        out['total_loss'] = out['bbox_pred']
        return out

    def compute_generator_loss(self, batch, model_out):
        """:62-105."""
        opt = self.opt
        imgs, objs = batch['imgs'], batch['objs']
        flows_pred = model_out[2]
        n = opt.n_frames_G - 1
        r_imgs, r_boxes, r_pred, r_act = self._relevant(batch, model_out)
        out = {}
        fake = self.netD_img(r_pred, objs, r_boxes, r_act)
        out['GAN_Img'] = hinge_loss(fake, True, False) * getattr(opt, 'discriminator_img_loss_weight', 1.0)
        if not getattr(opt, 'no_ganFeat_loss', False):
            real = self.netD_img(r_imgs, objs, r_boxes, r_act)
            feat = 0
            for i in range(len(fake)):
                for j in range(len(fake[i]) - 1):
                    feat = feat + F.l1_loss(fake[i][j], real[i][j].detach()) * getattr(opt, 'lambda_feat', 10.0) / len(fake)
            out['GAN_Feat'] = feat
        b, t, c, h, w = imgs.shape
        prev = imgs[:, n - 1:-1].reshape(-1, c, h, w)
        nxt = imgs[:, n:].reshape(-1, c, h, w)
        warped = flow_warp(prev, flows_pred[:, n - 1:-1].reshape(-1, 2, h, w))
        out['loss_F_Warp'] = F.l1_loss(warped, nxt) * getattr(opt, 'lambda_F_warp', 10.0)
        out['total_loss'] = torch.stack(list(out.values()), dim=0).sum()
        return out

    def compute_discriminator_loss(self, batch, model_out):
        """:107-133."""
        r_imgs, r_boxes, r_pred, r_act = self._relevant(batch, model_out)
        fake = self.netD_img(r_pred.detach(), batch['objs'], r_boxes, r_act)
        real = self.netD_img(r_imgs, batch['objs'], r_boxes, r_act)
        out = {'D_img_fake': hinge_loss(fake, False, True), 'D_img_real': hinge_loss(real, True, True)}
        out['total_img_loss'] = out['D_img_fake'] + out['D_img_real']
        return out

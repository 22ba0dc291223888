"""The three losses of one training iteration on the product path
(spade_models/loss_model.py:13-149, order fixed by scripts/train.py:440-493).

``LossModel(opt, discriminator)`` keeps the reference constructor and the
``forward(batch, model_out, mode)`` dispatch with the same mode strings and the
same dictionary keys.  Two things are evaluated differently from the reference
without changing a single result:

* the discriminator's conditioning (action-graph vectors and layout vectors) is
  computed once per ITERATION: shared by the fake and the real pass of a loss, and
  by the generator loss and the discriminator loss of the same ``model_out``;
* in the generator loss nothing needs a gradient with respect to the
  discriminator's parameters (train.py:523 zeroes them before the discriminator
  step) and the real pass only supplies detached feature targets
  (loss_model.py:84), so the conditioning and the real pass run without autograd
  and the fake pass back-propagates to the image only.

``--no_vgg_loss`` is implied: the VGG19 perceptual loss needs downloaded weights.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .networks import flow_warp


def hinge_loss(preds, target_is_real, for_discriminator):
    """GANLoss(gan_mode='hinge') over the multiscale output (networks/loss.py:61-97): the last
    level of every scale, averaged over scales."""
    total = 0
    for scale in preds:
        x = scale[-1]
        if not for_discriminator:
            assert target_is_real, "The generator's hinge loss must be aiming for real"
            l = -x.mean()
        elif target_is_real:
            l = -torch.clamp(x - 1, max=0).mean()
        else:
            l = -torch.clamp(-x - 1, max=0).mean()
        total = total + l
    return total / len(preds)


def _unpack(batch):
    if isinstance(batch, dict):
        return batch.get('imgs'), batch['objs'], batch['boxes']
    return batch[0], batch[1], batch[2]


class LossModel(nn.Module):
    def __init__(self, opt, discriminator):
        super().__init__()
        if getattr(opt, 'gan_mode', 'hinge') != 'hinge':
            raise NotImplementedError('gan_mode=%r: only the reference default "hinge" is built' % opt.gan_mode)
        if not getattr(opt, 'no_vgg_loss', False):
            # the reference's default (--no_vgg_loss is store_true, data/args.py) adds G_losses['VGG'] from a
            # downloaded VGG19 (networks/architecture.py:96); training without it would silently change the objective
            raise NotImplementedError('the VGG perceptual loss needs downloaded VGG19 weights and is not built: '
                                      'pass --no_vgg_loss (opt.no_vgg_loss = True)')
        self.opt = opt
        self.discriminator = discriminator
        self.netD_img = discriminator.img_discriminator

    def _relevant(self, batch, model_out):
        imgs, objs, boxes = _unpack(batch)
        imgs_pred, actions_data = model_out[0], model_out[4]
        n = self.opt.n_frames_G - 1
        return objs, imgs[:, n:], boxes[:, n:], imgs_pred[:, n:], [a[:, n:] for a in actions_data]

    def compute_graph_loss(self, batch, boxes_pred):
        """loss_model.py:40-60: masked smooth-L1 on the predicted boxes of frames 1.."""
        _, objs, boxes = _unpack(batch)
        w = getattr(self.opt, 'bbox_pred_loss_weight', 10)
        T = boxes.shape[1] - 1
        l = F.smooth_l1_loss(boxes_pred[:, 1:].reshape(-1, 4), boxes[:, 1:].reshape(-1, 4), reduction='none') * w
        real = (objs.sum(-1) != 0).unsqueeze(1).expand(objs.shape[0], T, objs.shape[1]).reshape(-1, 1)
        out = {'bbox_pred': (l * real.to(l.dtype)).mean()}
        out['total_loss'] = torch.stack(list(out.values()), dim=0).sum()
        return out

    def compute_generator_loss(self, batch, model_out):
        """loss_model.py:62-105."""
        opt = self.opt
        imgs = _unpack(batch)[0]
        flows_pred = model_out[2]
        n = opt.n_frames_G - 1
        objs, r_imgs, r_boxes, r_pred, r_act = self._relevant(batch, model_out)
        D = self.netD_img
        # The conditioning depends on the discriminator's parameters and on data only, and those parameters do not
        # move between this loss and the discriminator loss of the same iteration (train.py:446-464): evaluate it
        # ONCE, with its autograd graph, use a detached copy here and hand the attached one to
        # compute_discriminator_loss (keyed on the generator output it belongs to).
        cond_full = D.condition(objs, r_boxes, r_act) if torch.is_grad_enabled() else None
        self._cond_cache = (model_out[0], cond_full) if cond_full is not None else None
        params = [p for p in D.parameters() if p.requires_grad]
        for p in params:                   # no gradient w.r.t. the discriminator in the generator step
            p.requires_grad_(False)
        try:
            if cond_full is not None:
                cond = (cond_full[0].detach(),) + tuple(cond_full[1:])
            else:
                with torch.no_grad():
                    cond = D.condition(objs, r_boxes, r_act)
            fake = D(r_pred, objs, r_boxes, r_act, cond=cond)
            out = {'GAN_Img': hinge_loss(fake, True, False) * getattr(opt, 'discriminator_img_loss_weight', 1.0)}
            if not getattr(opt, 'no_ganFeat_loss', False):
                with torch.no_grad():
                    real = D(r_imgs, objs, r_boxes, r_act, cond=cond)
                feat = 0
                for i in range(len(fake)):
                    for j in range(len(fake[i]) - 1):
                        feat = feat + F.l1_loss(fake[i][j], real[i][j]) * (getattr(opt, 'lambda_feat', 10.0) / len(fake))
                out['GAN_Feat'] = feat
        finally:
            for p in params:
                p.requires_grad_(True)
        b, t, c, h, w = imgs.shape
        prev = imgs[:, n - 1:-1].reshape(-1, c, h, w)
        nxt = imgs[:, n:].reshape(-1, c, h, w)
        warped = flow_warp(prev, flows_pred[:, n - 1:-1].reshape(-1, 2, h, w))
        out['loss_F_Warp'] = F.l1_loss(warped, nxt) * getattr(opt, 'lambda_F_warp', 10.0)
        out['total_loss'] = torch.stack(list(out.values()), dim=0).sum()
        return out

    def compute_discriminator_loss(self, batch, model_out):
        """loss_model.py:107-133."""
        objs, r_imgs, r_boxes, r_pred, r_act = self._relevant(batch, model_out)
        D = self.netD_img
        cached = getattr(self, '_cond_cache', None)
        self._cond_cache = None
        if cached is not None and cached[0] is model_out[0]:
            cond = cached[1]               # built by compute_generator_loss on the same model_out, graph attached
        else:
            cond = D.condition(objs, r_boxes, r_act)
        fake = D(r_pred.detach(), objs, r_boxes, r_act, cond=cond)
        real = D(r_imgs, objs, r_boxes, r_act, cond=cond)
        out = {'D_img_fake': hinge_loss(fake, False, True), 'D_img_real': hinge_loss(real, True, True)}
        out['total_img_loss'] = torch.stack(list(out.values()), dim=0).sum()
        return out

    def forward(self, batch, model_out, mode):
        if mode == 'compute_discriminator_loss':
            return self.compute_discriminator_loss(batch, model_out)
        if mode == 'compute_generator_loss':
            return self.compute_generator_loss(batch, model_out)
        if mode == 'compute_graph_loss':
            return self.compute_graph_loss(batch, model_out)
        raise ValueError('unknown mode %r' % mode)

"""Host logic of Trainer on CPU stand-ins (non-fused optimisers): the NaN guard of scripts/train.py:450-453
and checkpoints in the reference's dictionary format (train.py:528-543, :27-55)."""
import io
from types import SimpleNamespace

import torch

from test_dist_gloo import _TinyD, _TinyModel, _tiny_batches


class _Losses(torch.nn.Module):
    def __init__(self, d, poison=False):
        super().__init__()
        self.discriminator = d
        self.netD_img = d.img_discriminator
        self.poison = poison

    def forward(self, batch, out, mode):
        D = self.netD_img
        if mode == 'compute_generator_loss':
            gan = -D(out[0]).mean()
            if self.poison:
                gan = gan * float('nan')
            feat = (out[0] - batch['imgs']).abs().mean()
            return {'GAN_Img': gan, 'GAN_Feat': feat, 'total_loss': gan + feat}
        if mode == 'compute_discriminator_loss':
            return {'total_img_loss': torch.relu(1 + D(out[0].detach())).mean() + torch.relu(1 - D(batch['imgs'])).mean()}
        return {'total_loss': torch.nn.functional.smooth_l1_loss(out, batch['boxes'])}


def _trainer(poison=False, seed=0):
    from ag2video_b200.trainer import Trainer
    torch.manual_seed(seed)
    model, d = _TinyModel(), _TinyD()
    opt = SimpleNamespace(learning_rate=1e-2, beta1=0.5, vocab={'dummy': 1})
    return Trainer(opt, model, d, _Losses(d, poison), world=1, fused=False), model, d


def _flat(mods):
    return torch.cat([p.detach().flatten().clone() for m in mods for p in m.parameters()])


def test_nan_generator_loss_skips_generator_and_discriminator_updates():
    tr, model, d = _trainer(poison=True)
    clip, graph = _tiny_batches(3, 2)
    gen0 = _flat([model.acts_to_objs, model.layout_to_video])
    d0, graph0 = _flat([d.img_discriminator]), _flat([model.acts_to_boxes])
    G, D, GG = tr.iteration(clip, graph)
    assert torch.isnan(G['GAN_Img']) and D is None
    assert torch.equal(_flat([model.acts_to_objs, model.layout_to_video]), gen0)      # no poisoned weights
    assert torch.equal(_flat([d.img_discriminator]), d0)
    assert not torch.equal(_flat([model.acts_to_boxes]), graph0)                       # the graph step still runs
    assert all(torch.isfinite(p).all() for p in model.parameters())
    assert len(tr.optimizer_generator.state) == 0                                      # Adam state untouched


def test_checkpoint_round_trip_in_the_reference_format():
    from ag2video_b200.trainer import from_reference_keys, to_reference_keys
    tr, model, d = _trainer()
    clip, graph = _tiny_batches(5, 2)
    tr.iteration(clip, graph)
    tr.iteration(clip, graph)
    ck = tr.state_dict()
    assert set(ck) >= {'model_state', 'gans_model_state', 'd_img_state', 'd_img_optim_state', 'optim_state_gen',
                       'optim_state_graph', 'vocab', 'counters'}
    assert ck['counters'] == {'t': 2, 'epoch': 0}
    # the reference wraps its three sub-models (and the loss model) in DataParallel: '.module.' / 'module.' keys
    assert 'acts_to_boxes.module.weight' in ck['model_state'] and 'layout_to_video.module.bias' in ck['model_state']
    assert all(k.startswith('module.') for k in ck['gans_model_state'])
    assert from_reference_keys(to_reference_keys(model.state_dict())).keys() == model.state_dict().keys()
    buf = io.BytesIO()
    torch.save(ck, buf)
    buf.seek(0)
    tr2, model2, d2 = _trainer(seed=1)
    info = tr2.load_state_dict(torch.load(buf, weights_only=False))
    assert info['optim_state_gen_loaded'] and tr2.t == 2
    assert torch.equal(_flat([model2, d2.img_discriminator]), _flat([model, d.img_discriminator]))
    # both continue identically (optimiser moments and step counts restored)
    tr.iteration(clip, graph)
    tr2.iteration(clip, graph)
    assert torch.equal(_flat([model2, d2.img_discriminator]), _flat([model, d.img_discriminator]))
    # a reference-written checkpoint has no portable generator-optimiser order: it is skipped, the rest loads
    ref_ck = {k: v for k, v in ck.items() if k != 'optim_state_gen_order'}
    tr3, _, _ = _trainer(seed=2)
    assert tr3.load_state_dict(ref_ck)['optim_state_gen_loaded'] is False


def test_loss_model_refuses_the_default_vgg_objective():
    import pytest
    from ag2video_b200.config import make_opt
    from ag2video_b200.losses import LossModel
    opt = make_opt(64, no_vgg_loss=False)
    with pytest.raises(NotImplementedError):
        LossModel(opt, SimpleNamespace(img_discriminator=torch.nn.Linear(1, 1)))


def test_optimizer_step_invalidates_weight_packs():
    """a pack built by a forward that is NOT followed by a backward (validation pass) must not
    survive the next optimiser step."""
    import ag2video_b200.spade as sp
    tr, model, d = _trainer()
    cache = sp._PackCache()
    w = torch.nn.Parameter(torch.ones(2))
    built = []
    get = lambda: cache.get('fwd', (w,), lambda: built.append(1) or len(built))
    cache.forward_begin(); assert get() == 1
    cache.forward_begin(); assert get() == 1                # second forward, no backward in between: reused
    clip, graph = _tiny_batches(9, 2)
    tr.iteration(clip, graph)                               # optimiser steps -> post-step hook
    cache.forward_begin(); assert get() == 2

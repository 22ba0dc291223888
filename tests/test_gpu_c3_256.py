"""End-to-end parity at the BASELINE config 3 shapes (256x256) on a bounded sample (1 clip x 2 frames,
one generated frame): the product path against the REFERENCE's own outputs
(tests/golden/generator256.pt, losses256.pt, produced by tests/golden/make_golden.py generator256).

Bars (north_star / SURVEY.md 8d): loss <= 1e-3; boxes <= 1e-3; pixels: reported, bar 3e-2 (18 stacked
TF32 SPADE layers with variance-preserving random weights; the fp32 eager-GPU path itself is ~4e-3 from
the CPU golden); graph-step results (no TF32 on that path) <= 1e-4.
"""
import pytest
import torch

from _util import det_state, golden, max_rel, rel_l2
from ag2video_b200.config import make_opt, synthetic_batch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _exact_library_convs():
    import ag2video_b200.spade as sp
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, sp.CONV_IMPL)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, sp.CONV_IMPL = old


def _model(seed):
    from ag2video_b200.networks import AG2VideoModel
    m = AG2VideoModel(make_opt(256, batch_size=1))
    m.load_state_dict(det_state(m.state_dict(), seed), strict=True)
    return m.cuda().to(memory_format=torch.channels_last).train()


@pytest.mark.parametrize('mode', ['validation_3xtf32', 'product_tf32'])
def test_generator256_vs_reference_golden_loss_1e3(mode):
    import ag2video_b200.spade as sp
    sp.CONV_IMPL = 3 if mode == 'validation_3xtf32' else 0
    c = golden('generator256.pt')
    m = _model(c['seed'])
    b = synthetic_batch(B=1, F=2, image_size=256, seed=c['batch_seed'], device='cuda')
    imgs_pred, boxes_pred, flows, conf, _ = m(b['imgs'], b['objs'], b['triplets'], b['actions'],
                                              boxes_gt=b['boxes'], use_gt=True)
    e_img, e_box, e_flow = max_rel(imgs_pred, c['imgs_pred']), max_rel(boxes_pred, c['boxes_pred']), max_rel(flows, c['flows'])
    loss = (imgs_pred - b['imgs']).abs().mean() + (boxes_pred - b['boxes'])[:, 1:].abs().mean()
    e_loss = abs(float(loss.detach()) - float(c['loss'])) / abs(float(c['loss']))
    print('generator256 [%s]: loss %.2e  boxes %.2e  pixels max %.2e rel-L2 %.2e  flows %.2e'
          % (mode, e_loss, e_box, e_img, rel_l2(imgs_pred, c['imgs_pred']), e_flow))
    assert e_loss <= 1e-3 and e_box <= 1e-3
    assert e_img <= (1e-2 if mode == 'validation_3xtf32' else 3e-2)
    loss.backward()
    grads = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    # graph-side parameters of acts_to_boxes do not pass through the SPADE stack: tight
    for k, n in c['grad_norms'].items():
        if k.startswith('acts_to_boxes') and n > 1e-6:
            assert abs(float(grads[k].norm()) - n) <= 1e-3 * n, (k, float(grads[k].norm()), n)
    worst = 0.0
    for k, v in c['grad_picks'].items():
        if float(v.abs().max()) < 1e-6:
            continue
        e = rel_l2(grads[k].contiguous().flatten()[:4096], v)
        print('  %-66s rel-L2 %.2e' % (k, e))
        worst = max(worst, e)
    # Gross-error check only: through ~40 ReLU / LeakyReLU layers with random weights, gates within rounding
    # of zero flip, and the L1 loss on ONE generated frame adds sign() (measured on B200: 3.6e-2 in the
    # 3xTF32 validation mode, 1.4e-1 with TF32 operands).  The strict statement is
    # tests/test_gpu_spade_gates.py: every SPADE gradient within 1e-3 on the branch actually taken.
    assert worst <= (0.1 if mode == 'validation_3xtf32' else 0.2)
    # (the flow network's gradients pass through a hard threshold on one frame: looser)
    bad = [(k, float(grads[k].norm()), n) for k, n in c['grad_norms'].items()
           if n > 1e-4 and abs(float(grads[k].norm()) - n) > (0.25 if 'flows_network' in k else 0.1) * n]
    assert not bad, bad[:5]


def test_iteration256_losses_vs_reference_golden_1e3():
    """Trainer.iteration at 256x256 (generator, discriminator and graph step) against the reference's
    seven loss values; the graph step (K1 recurrence only) also against its boxes and full gradients."""
    from ag2video_b200.discriminator import MetaDiscriminatorModel
    from ag2video_b200.losses import LossModel
    from ag2video_b200.trainer import Trainer
    c = golden('losses256.pt')
    dev = torch.device('cuda', 0)
    opt = make_opt(256, batch_size=1)
    m = _model(c['seed_g'])
    meta = MetaDiscriminatorModel(opt, device=dev)
    meta.img_discriminator.load_state_dict(det_state(meta.img_discriminator.state_dict(), c['seed_d']), strict=True)
    tr = Trainer(opt, m, meta, LossModel(opt, meta))
    b = synthetic_batch(B=1, F=2, image_size=256, seed=c['batch_seed'], device='cuda')
    bg = synthetic_batch(B=1, F=16, image_size=256, seed=c['graph_batch_seed'], with_images=False, device='cuda')
    # the graph step alone first (does not touch the weights): boxes and gradients, no TF32 on this path
    bp = m(None, bg['objs'], bg['triplets'], bg['actions'], boxes_gt=bg['boxes'], graph_only=True)
    assert max_rel(bp, c['graph_boxes_pred']) <= 1e-4
    GG = tr.gans_model(bg, bp, mode='compute_graph_loss')
    GG['total_loss'].backward()
    grads = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    for k, v in c['graph_grad_full'].items():
        assert max_rel(grads[k], v) <= 1e-3, (k, max_rel(grads[k], v))
    for k, n in c['graph_grad_norms'].items():
        if n > 1e-6:
            assert abs(float(grads[k].norm()) - n) <= 1e-3 * n, k
    m.zero_grad(set_to_none=True)
    G, D, GG = tr.iteration(b, bg)
    torch.cuda.synchronize()
    errs = {}
    for name, got, want in (('G', G, c['G']), ('D', D, c['D']), ('graph', GG, c['graph'])):
        for k, v in want.items():
            errs['%s.%s' % (name, k)] = abs(float(got[k]) - float(v)) / abs(float(v))
    print('iteration256 losses vs reference: ' + ', '.join('%s %.2e' % kv for kv in errs.items()))
    assert errs['graph.total_loss'] <= 1e-4 and errs['D.D_img_real'] <= 1e-3, errs
    assert errs['G.total_loss'] <= 1e-3 and errs['D.total_img_loss'] <= 1e-3, errs
    assert max(errs.values()) <= 1e-2, errs

"""Pin the oracle: every restated function against fixtures produced by the
reference code itself (tests/golden/make_golden.py).  CPU only."""
import types

import pytest
import torch

from _util import det_state, golden, max_rel
from ag2video_b200.config import make_opt, synthetic_batch
from oracle import networks as onet
from oracle import ops as oops

TOL = 2e-6      # same torch build, same primitives: differences are summation-order noise


def _grads(module):
    return {k: p.grad for k, p in module.named_parameters() if p.grad is not None}


def _check_grads(got, want, tol=TOL):
    assert set(got) == set(want)
    for k in want:
        assert max_rel(got[k], want[k]) <= tol, k


def test_gconv_matches_reference():
    g = golden('gconv.pt')
    layer = oops.GraphTripleConv(**g['dims'])
    layer.load_state_dict(g['state'], strict=True)
    obj, pred = g['obj'].clone().requires_grad_(), g['pred'].clone().requires_grad_()
    new_obj, new_p = layer(obj, pred, g['edges'], g['ind'])
    assert torch.equal(new_obj, g['new_obj']) or max_rel(new_obj, g['new_obj']) <= TOL
    assert max_rel(new_p, g['new_p']) <= TOL
    ((new_obj * g['c1']).sum() + (new_p * g['c2']).sum()).backward()
    assert max_rel(obj.grad, g['dobj']) <= TOL
    assert max_rel(pred.grad, g['dpred']) <= TOL
    _check_grads(_grads(layer), g['dparams'])


def test_gconv_net_matches_reference():
    g = golden('gconv_net.pt')
    net = oops.GraphTripleConvNet(g['layers'])
    net.load_state_dict(g['state'], strict=True)
    obj, pred = g['obj'].clone().requires_grad_(), g['pred'].clone().requires_grad_()
    o, p = net(obj, pred, g['edges'], g['ind'])
    assert max_rel(o, g['new_obj']) <= TOL and max_rel(p, g['new_p']) <= TOL
    ((o * g['c1']).sum() + (p * g['c2']).sum()).backward()
    assert max_rel(obj.grad, g['dobj']) <= TOL and max_rel(pred.grad, g['dpred']) <= TOL
    _check_grads(_grads(net), g['dparams'])


@pytest.mark.parametrize('name', ['mixed32', 'demo64', 'rect', 'avg', 'padbox', 'cater128'])
def test_boxes_to_layout_matches_reference(name):
    c = golden('layout.pt')[name]
    vecs = c['vecs'].clone().requires_grad_()
    out = oops.boxes_to_layout(vecs, c['boxes'], c['H'], c['W'], pooling=c['pooling'])
    assert torch.equal(out != 0, c['out'] != 0), 'pixel support must be bit-exact'
    assert max_rel(out, c['out']) <= TOL
    (out * c['cot']).sum().backward()
    assert max_rel(vecs.grad, c['dvecs']) <= TOL


@pytest.mark.parametrize('name', ['m5_train', 'm5_test', 'm16_train', 'm16_test'])
def test_masks_to_layout_matches_reference(name):
    c = golden('masks_layout.pt')[name]
    vecs = c['vecs'].clone().requires_grad_()
    out = oops.masks_to_layout(vecs, c['boxes'], c['masks'], c['H'], test_mode=c['test_mode'])
    assert torch.equal(out != 0, c['out'] != 0)
    assert max_rel(out, c['out']) <= TOL
    (out * c['cot']).sum().backward()
    assert max_rel(vecs.grad, c['dvecs']) <= TOL


def test_crop_bbox_batch_matches_reference():
    from ag2video_b200.config import cater_vocab
    c = golden('crop.pt')
    imgs = c['imgs'].clone().requires_grad_()
    crops, flat = oops.crop_bbox_batch(imgs, c['objs'], c['boxes'], c['HH'], vocab=cater_vocab())
    assert len(crops) == len(c['crops'])
    for got, want, gf, wf in zip(crops, c['crops'], flat, c['objs_flat']):
        assert got.shape == want.shape and max_rel(got, want) <= TOL
        assert torch.equal(gf, wf)
    sum((a * k).sum() for a, k in zip(crops, c['cots'])).backward()
    assert max_rel(imgs.grad, c['dimgs']) <= TOL


@pytest.mark.parametrize('name', ['c16_r8', 'c8_r16', 'dy_c16_r8', 'dy_c8_r16'])
def test_spade_matches_reference(name):
    c = golden('spade.pt')[name]
    m = oops.SPADE('spadesyncbatch3x3', c['C'], c['L'])
    m.load_state_dict(c['state'], strict=True)
    m.train()
    x, seg = c['x'].clone().requires_grad_(), c['seg'].clone().requires_grad_()
    out = m(x, seg)
    assert max_rel(out, c['out']) <= TOL
    (out * c['cot']).sum().backward()
    assert max_rel(x.grad, c['dx']) <= 5e-6 and max_rel(seg.grad, c['dseg']) <= 5e-6
    _check_grads(_grads(m), c['dparams'], 5e-6)
    for k, v in c['state_after'].items():
        assert max_rel(m.state_dict()[k].float(), v.float()) <= TOL, k
    m.eval()
    with torch.no_grad():
        assert max_rel(m(x, seg), c['out_eval']) <= TOL


@pytest.mark.parametrize('name', ['b16_8', 'b8_8', 'dy_b16_8', 'dy_b8_8'])
def test_spade_resnet_block_matches_reference(name):
    c = golden('spade_block.pt')[name]
    opt = types.SimpleNamespace(norm_G='spectralspadesyncbatch3x3', semantic_nc=8)
    m = oops.SPADEResnetBlock(c['fin'], c['fout'], opt)
    m.load_state_dict(c['state'], strict=True)
    m.train()
    x, seg = c['x'].clone().requires_grad_(), c['seg'].clone().requires_grad_()
    out = m(x, seg)
    assert max_rel(out, c['out']) <= TOL
    (out * c['cot']).sum().backward()
    assert max_rel(x.grad, c['dx']) <= 5e-6 and max_rel(seg.grad, c['dseg']) <= 5e-6
    _check_grads(_grads(m), c['dparams'], 5e-6)
    for k, v in c['state_after'].items():
        assert max_rel(m.state_dict()[k].float(), v.float()) <= TOL, k


def test_acts2layout_matches_reference():
    c = golden('acts2layout.pt')
    opt = make_opt(32, **c['over'])
    m = onet.Acts2LayoutModel(opt)
    m.load_state_dict(det_state(m.state_dict(), c['seed']), strict=True)
    b = c['batch']
    obj_vecs, boxes_pred, extra = m(b['objs'], b['triplets'], b['actions'], b['boxes'])
    assert torch.equal(extra[1], c['temporal_triplets'])
    assert max_rel(extra[2], c['rel_t']) <= TOL
    assert max_rel(obj_vecs, c['obj_vecs']) <= TOL and max_rel(boxes_pred, c['boxes_pred']) <= TOL
    ((obj_vecs * c['c1']).sum() + (boxes_pred * c['c2']).sum()).backward()
    _check_grads(_grads(m), c['dparams'], 5e-6)


def test_generator64_matches_reference():
    """BASELINE config 1 end to end: the whole oracle generator against the
    reference's AG2VideoModel at 64x64, batch 2, 4 frames."""
    c = golden('generator64.pt')
    torch.manual_seed(0)
    opt = make_opt(64, batch_size=2)
    m = onet.AG2VideoModel(opt)
    m.load_state_dict(det_state(m.state_dict(), c['seed']), strict=True)
    m.train()
    b = synthetic_batch(B=2, F=4, image_size=64, seed=c['batch_seed'])
    imgs_pred, boxes_pred, flows, conf, _ = m(b['imgs'], b['objs'], b['triplets'], b['actions'],
                                              boxes_gt=b['boxes'], use_gt=True)
    assert max_rel(imgs_pred, c['imgs_pred']) <= 2e-5
    assert max_rel(boxes_pred, c['boxes_pred']) <= 2e-5
    assert max_rel(flows, c['flows']) <= 2e-4
    loss = (imgs_pred - b['imgs']).abs().mean() + (boxes_pred - b['boxes'])[:, 1:].abs().mean()
    assert abs(float(loss) - float(c['loss'])) <= 1e-5 * abs(float(c['loss']))
    loss.backward()
    grads = _grads(m)
    for k, v in c['grad_picks'].items():
        assert max_rel(grads[k].flatten()[:4096], v) <= 5e-4, k
    bad = [k for k, n in c['grad_norms'].items()
           if abs(float(grads[k].norm()) - n) > 2e-3 * max(n, 1e-6)]
    assert not bad, bad[:5]


def test_losses64_match_reference():
    """One training iteration's losses (scripts/train.py:446-493) at BASELINE config 1: the oracle
    discriminator + LossModel against the reference's, on the oracle generator's output."""
    from oracle import losses as oloss
    c = golden('losses64.pt')
    opt = make_opt(64, batch_size=2)
    m = onet.AG2VideoModel(opt)
    m.load_state_dict(det_state(m.state_dict(), c['seed_g']), strict=True)
    m.train()
    netD = oloss.MultiscaleActionDiscriminator(opt)
    netD.load_state_dict(det_state(netD.state_dict(), c['seed_d']), strict=True)
    netD.train()
    lm = oloss.LossModel(opt, netD)
    b = synthetic_batch(B=2, F=4, image_size=64, seed=c['batch_seed'])
    out = m(b['imgs'], b['objs'], b['triplets'], b['actions'], boxes_gt=b['boxes'], use_gt=True)
    n = opt.n_frames_G - 1
    with torch.no_grad():
        d_real = netD(b['imgs'][:, n:], b['objs'], b['boxes'][:, n:], [a[:, n:] for a in out[4]])
    for s, scale in enumerate(d_real):
        assert max_rel(scale[-1], c['d_real_last'][s]) <= 2e-5
        for j, o in enumerate(scale):
            assert max_rel(o.flatten()[:1024], c['d_real_picks'][s][j]) <= 2e-5, (s, j)
            assert abs(float(o.norm()) - c['d_real_norms'][s][j]) <= 1e-5 * c['d_real_norms'][s][j]
    G = lm.compute_generator_loss(b, out)
    # the warp loss bilinearly samples WHITE-NOISE frames at flow-displaced positions: the 2e-4
    # summation-order noise of the flow network (see test_generator64) moves it by ~1e-3, so it is
    # checked strictly on the reference's own flows and loosely on the oracle generator's
    strict = lm.compute_generator_loss(b, (out[0], out[1], golden('generator64.pt')['flows'], out[3], out[4]))
    assert abs(float(strict['loss_F_Warp']) - float(c['G']['loss_F_Warp'])) <= 2e-5 * abs(float(c['G']['loss_F_Warp']))
    for k, v in c['G'].items():
        tol = 5e-3 if k in ('loss_F_Warp', 'total_loss') else 2e-5
        assert abs(float(G[k]) - float(v)) <= tol * abs(float(v)), k
    m.zero_grad(); netD.zero_grad()
    G['total_loss'].backward()
    grads = _grads(m)
    for k, v in c['g_grad_picks'].items():
        assert max_rel(grads[k].flatten()[:4096], v) <= 1e-3, k
    bad = [k for k, nrm in c['g_grad_norms'].items() if abs(float(grads[k].norm()) - nrm) > 5e-3 * max(nrm, 1e-6)]
    assert not bad, bad[:5]
    netD.zero_grad()
    D = lm.compute_discriminator_loss(b, out)
    for k, v in c['D'].items():
        assert abs(float(D[k]) - float(v)) <= 2e-5 * abs(float(v)), k
    D['total_img_loss'].backward()
    dg = _grads(netD)
    for k, v in c['d_grad_picks'].items():
        # hinge gates of the D loss flip for the few logits within 2e-5 of +-1 (the fake frames carry
        # the generator's summation-order noise): gradients move by O(flipped share), not O(eps)
        assert max_rel(dg[k].flatten()[:2048], v) <= 1e-2, k
    bad = [k for k, nrm in c['d_grad_norms'].items() if abs(float(dg[k].norm()) - nrm) > 5e-3 * max(nrm, 1e-6)]
    assert not bad, bad[:5]
    bg = synthetic_batch(B=2, F=16, image_size=64, seed=c['graph_batch_seed'], with_images=False)
    boxes_pred = m(None, bg['objs'], bg['triplets'], bg['actions'], boxes_gt=bg['boxes'], graph_only=True)
    GG = lm.compute_graph_loss(bg, boxes_pred)
    assert abs(float(GG['total_loss']) - float(c['graph']['total_loss'])) <= 1e-5 * abs(float(c['graph']['total_loss']))
    m.zero_grad()
    GG['total_loss'].backward()
    gg = _grads(m)
    bad = [k for k, nrm in c['graph_grad_norms'].items() if abs(float(gg[k].norm()) - nrm) > 1e-3 * max(nrm, 1e-6)]
    assert not bad, bad[:5]


@pytest.mark.parametrize('name', ['mixed32', 'demo64', 'rect', 'avg', 'cater128'])
def test_closed_form_layout_matches_reference(name):
    """The numpy closed form (no grid_sample) against the reference's boxes_to_layout: identical
    pixel support, values to fp32 summation noise - an independent pin of K2's contract."""
    import numpy as np
    from oracle import closed_form
    c = golden('layout.pt')[name]
    out, support = closed_form.boxes_to_layout(c['vecs'], c['boxes'], c['H'], c['W'], pooling=c['pooling'])
    ref = c['out'].numpy()
    assert out.shape == ref.shape
    # support of the whole layout: a pixel is touched iff some object covers it (generic vecs: no cancellation)
    assert np.array_equal(support.any(axis=0), (ref[0] != 0).any(axis=0))
    assert max_rel(torch.from_numpy(out), c['out']) <= 2e-6
    # per-object support against the restated torch path on single-object layouts
    for o in range(c['boxes'].shape[0]):
        if not bool(c['boxes'][o].any()):
            continue
        one = oops.boxes_to_layout(torch.ones(1, 1), c['boxes'][o:o + 1], c['H'], c['W'])
        assert np.array_equal(support[o], (one[0, 0] != 0).numpy()), o


@pytest.mark.parametrize('name', ['all_masked', 'single_node', 'one_live_edge'])
def test_gconv_edge_cases_match_reference(name):
    g = golden('gconv_edge.pt')
    c = g['cases'][name]
    layer = oops.GraphTripleConv(**g['dims'])
    layer.load_state_dict(g['state'], strict=True)
    obj, pred = c['obj'].clone().requires_grad_(), c['pred'].clone().requires_grad_()
    new_obj, new_p = layer(obj, pred, c['edges'], c['ind'])
    assert max_rel(new_obj, c['new_obj']) <= TOL and max_rel(new_p, c['new_p']) <= TOL
    ((new_obj * c['c1']).sum() + (new_p * c['c2']).sum()).backward()
    assert max_rel(obj.grad, c['dobj']) <= TOL and max_rel(pred.grad, c['dpred']) <= TOL
    _check_grads(_grads(layer), c['dparams'])


def test_generator256_and_losses256_match_reference():
    """BASELINE config 3 shapes (256x256) on a bounded sample (1 clip x 2 frames): the oracle generator and
    one iteration's three losses against the reference's (tests/golden/generator256.pt, losses256.pt)."""
    from oracle import losses as oloss
    c, cl = golden('generator256.pt'), golden('losses256.pt')
    opt = make_opt(256, batch_size=1)
    m = onet.AG2VideoModel(opt)
    state0 = det_state(m.state_dict(), c['seed'])
    m.load_state_dict(state0, strict=True)
    m.train()
    b = synthetic_batch(B=1, F=2, image_size=256, seed=c['batch_seed'])
    imgs_pred, boxes_pred, flows, conf, _ = m(b['imgs'], b['objs'], b['triplets'], b['actions'],
                                              boxes_gt=b['boxes'], use_gt=True)
    assert max_rel(imgs_pred, c['imgs_pred']) <= 5e-4 and max_rel(boxes_pred, c['boxes_pred']) <= 2e-5
    loss = (imgs_pred - b['imgs']).abs().mean() + (boxes_pred - b['boxes'])[:, 1:].abs().mean()
    assert abs(float(loss) - float(c['loss'])) <= 1e-5 * abs(float(c['loss']))
    loss.backward()
    grads = _grads(m)
    bad = [k for k, n in c['grad_norms'].items() if abs(float(grads[k].norm()) - n) > 5e-3 * max(n, 1e-6)]
    assert not bad, bad[:5]
    # the iteration's losses on a fresh copy of the weights
    m.load_state_dict(state0, strict=True)
    netD = oloss.MultiscaleActionDiscriminator(opt)
    netD.load_state_dict(det_state(netD.state_dict(), cl['seed_d']), strict=True)
    netD.train()
    lm = oloss.LossModel(opt, netD)
    out = m(b['imgs'], b['objs'], b['triplets'], b['actions'], boxes_gt=b['boxes'], use_gt=True)
    G = lm.compute_generator_loss(b, out)
    for k, v in cl['G'].items():
        tol = 5e-3 if k in ('loss_F_Warp', 'total_loss') else 1e-4
        assert abs(float(G[k]) - float(v)) <= tol * abs(float(v)), (k, float(G[k]), float(v))
    D = lm.compute_discriminator_loss(b, out)
    for k, v in cl['D'].items():
        assert abs(float(D[k]) - float(v)) <= 1e-4 * abs(float(v)), k
    bg = synthetic_batch(B=1, F=16, image_size=256, seed=cl['graph_batch_seed'], with_images=False)
    bp = m(None, bg['objs'], bg['triplets'], bg['actions'], boxes_gt=bg['boxes'], graph_only=True)
    assert max_rel(bp, cl['graph_boxes_pred']) <= 2e-5
    GG = lm.compute_graph_loss(bg, bp)
    assert abs(float(GG['total_loss']) - float(cl['graph']['total_loss'])) <= 1e-5 * abs(float(cl['graph']['total_loss']))
    m.zero_grad()
    GG['total_loss'].backward()
    gg = _grads(m)
    for k, v in cl['graph_grad_full'].items():
        assert max_rel(gg[k], v) <= 1e-4, k


def test_k2b_config4_oracle_matches_reference():
    """masks_to_layout / crop_bbox_batch at the BASELINE config 4 sizes: the oracle reproduces the
    reference's outputs (tests/golden/k2b_c4.pt)."""
    from _util import c4_masks, det_tensor
    from ag2video_b200.config import cater_vocab
    g = golden('k2b_c4.pt')
    for name in ('m16_train', 'm256_train', 'm256_test'):
        c = g[name]
        out = oops.masks_to_layout(det_tensor('k2bc4.%s' % name, (10, 2), 8), c['boxes'], c4_masks(name, c['M']), 256,
                                   test_mode=c['test_mode'])
        assert max_rel(out, c['out']) <= TOL, name
    b = synthetic_batch(B=2, F=4, image_size=256, seed=g['crop']['batch_seed'], n_objects=10)
    crops, flat = oops.crop_bbox_batch(b['imgs'], b['objs'], b['boxes'], 32, vocab=cater_vocab())
    for got, want, gf, wf in zip(crops, g['crop']['crops'], flat, g['crop']['objs_flat']):
        assert max_rel(got, want) <= TOL and torch.equal(gf, wf)


# ---- gradients for inputs that are data on the training path, and the 'jj' sampling backend ------------------
@pytest.mark.parametrize('name', ['layout_sum32', 'layout_rect', 'layout_avg64'])
def test_boxes_to_layout_box_gradient_matches_reference(name):
    c = golden('input_grads.pt')[name]
    vecs, boxes = c['vecs'].clone().requires_grad_(), c['boxes'].clone().requires_grad_()
    out = oops.boxes_to_layout(vecs, boxes, c['H'], c['W'], pooling=c['pooling'])
    (out * c['cot']).sum().backward()
    assert max_rel(vecs.grad, c['dvecs']) <= TOL and max_rel(boxes.grad, c['dboxes']) <= TOL
    assert float(c['dboxes'].abs().max()) > 0


@pytest.mark.parametrize('name', ['masks_m5', 'masks_m16'])
def test_masks_to_layout_input_gradients_match_reference(name):
    c = golden('input_grads.pt')[name]
    vecs, boxes, masks = (c[k].clone().requires_grad_() for k in ('vecs', 'boxes', 'masks'))
    out = oops.masks_to_layout(vecs, boxes, masks, c['H'])
    (out * c['cot']).sum().backward()
    assert max_rel(vecs.grad, c['dvecs']) <= TOL and max_rel(boxes.grad, c['dboxes']) <= TOL
    assert max_rel(masks.grad, c['dmasks']) <= TOL


def test_crop_bbox_jj_backend_matches_reference():
    """bilinear_sample (bilinear.py:134-189): clamped taps, two boxes reach outside the image."""
    c = golden('input_grads.pt')['crop_jj']
    feats = c['feats'].clone().requires_grad_()
    crops = oops.crop_bbox(feats, c['bbox'], c['HH'], c['WW'], backend='jj')
    assert max_rel(crops, c['crops']) <= TOL
    (crops * c['cot']).sum().backward()
    assert max_rel(feats.grad, c['dfeats']) <= TOL

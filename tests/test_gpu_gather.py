"""K2b parity on the GPU: masks_to_layout and crop_bbox_batch through the C ABI
vs golden vectors produced by the reference.  Tap indices / support bit-exact,
values <= 1e-5 relative."""
import pytest
import torch

from _util import golden, max_rel
from ag2video_b200.config import cater_vocab

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.mark.parametrize('name', ['m5_train', 'm5_test', 'm16_train', 'm16_test'])
def test_masks_to_layout_golden(name):
    from ag2video_b200.layout import masks_to_layout
    c = golden('masks_layout.pt')[name]
    vecs = c['vecs'].cuda().requires_grad_()
    out = masks_to_layout(vecs, c['boxes'].cuda(), c['masks'].cuda(), c['H'], test_mode=c['test_mode'])
    assert torch.equal(out.cpu() != 0, c['out'] != 0)
    assert max_rel(out, c['out']) <= TOL
    if not c['test_mode']:
        (out * c['cot'].cuda()).sum().backward()
        assert max_rel(vecs.grad, c['dvecs']) <= TOL


def test_crop_bbox_batch_golden():
    from ag2video_b200.bilinear import crop_bbox_batch
    c = golden('crop.pt')
    imgs = c['imgs'].cuda().requires_grad_()
    crops, flat = crop_bbox_batch(imgs, c['objs'].cuda(), c['boxes'].cuda(), c['HH'], vocab=cater_vocab())
    assert len(crops) == len(c['crops'])
    for got, want, gf, wf in zip(crops, c['crops'], flat, c['objs_flat']):
        assert got.shape == want.shape
        assert max_rel(got, want) <= TOL
        assert torch.equal(gf.cpu(), wf)
    sum((a * k.cuda()).sum() for a, k in zip(crops, c['cots'])).backward()
    assert max_rel(imgs.grad, c['dimgs']) <= TOL


def test_crop_bbox_single_matches_batch_path():
    from ag2video_b200.bilinear import crop_bbox
    g = torch.Generator().manual_seed(0)
    feats = torch.randn(3, 4, 20, 28, generator=g).cuda()
    bbox = torch.tensor([[0.1, 0.2, 0.5, 0.4], [0.0, 0.0, 1.0, 1.0], [0.6, 0.5, 0.6, 0.7]]).cuda()
    out = crop_bbox(feats, bbox, 8, 6)
    from oracle import ops as oops
    ref = oops.crop_bbox(feats.cpu(), bbox.cpu(), 8, 6)
    assert max_rel(out, ref) <= TOL


# ---- BASELINE config 4 sizes (10 objects, 256x256; masks 16 / 256 wide; crops 256 -> 32, B=2, N=4) ----------
@pytest.mark.parametrize('name', ['m16_train', 'm256_train', 'm256_test'])
def test_masks_to_layout_config4_golden_tol1e5(name):
    """masks_to_layout at the C4 sizes against the reference's own output (tests/golden/k2b_c4.pt):
    pixel support bit-exact, values and dvecs <= 1e-5."""
    from _util import c4_masks, det_tensor
    from ag2video_b200.layout import masks_to_layout
    c = golden('k2b_c4.pt')[name]
    masks = c4_masks(name, c['M']).cuda()
    vecs = det_tensor('k2bc4.%s' % name, (10, 2), 8).cuda().requires_grad_()
    out = masks_to_layout(vecs, c['boxes'].cuda(), masks, 256, test_mode=c['test_mode'])
    assert torch.equal(out.cpu() != 0, c['out'] != 0)
    assert max_rel(out, c['out']) <= TOL
    if not c['test_mode']:
        cot = det_tensor('k2bc4.cot.%s' % name, out.shape, 8).cuda()
        (out * cot).sum().backward()
        assert max_rel(vecs.grad, c['dvecs']) <= TOL


@pytest.mark.parametrize('M', [16, 256])
def test_masks_to_layout_config4_full_width_vs_oracle_tol1e5(M):
    """The full C4 case (D=512 channels, 134 MB output) against the CPU oracle."""
    from _util import c4_masks, det_tensor
    from ag2video_b200.layout import masks_to_layout
    from oracle import ops as oops
    c = golden('k2b_c4.pt')['m%d_train' % M]
    masks = c4_masks('full%d' % M, M)
    vecs = det_tensor('k2bc4.full.%d' % M, (10, 512), 8)
    want = oops.masks_to_layout(vecs, c['boxes'], masks, 256)
    got = masks_to_layout(vecs.cuda(), c['boxes'].cuda(), masks.cuda(), 256)
    assert torch.equal(got.cpu() != 0, want != 0)
    assert max_rel(got, want) <= TOL


def test_crop_bbox_batch_config4_golden_tol1e5():
    """crop_bbox_batch at the C4 sizes (B=2, N=4 frames of 256x256, 10 objects + dummy, 32x32 crops)
    against the reference's output; the input gradient against its sub-sampled golden and norm."""
    from _util import det_tensor
    from ag2video_b200.bilinear import crop_bbox_batch
    from ag2video_b200.config import synthetic_batch
    c = golden('k2b_c4.pt')['crop']
    b = synthetic_batch(B=2, F=4, image_size=256, seed=c['batch_seed'], n_objects=10, device='cuda')
    imgs = b['imgs'].clone().requires_grad_()
    crops, flat = crop_bbox_batch(imgs, b['objs'], b['boxes'], 32, vocab=cater_vocab())
    assert [tuple(x.shape) for x in crops] == [tuple(x.shape) for x in c['crops']]
    for got, want, gf, wf in zip(crops, c['crops'], flat, c['objs_flat']):
        assert max_rel(got, want) <= TOL
        assert torch.equal(gf.cpu(), wf)
    cots = [det_tensor('k2bc4.crop.cot.%d' % i, x.shape, 9).cuda() for i, x in enumerate(crops)]
    sum((a * k).sum() for a, k in zip(crops, cots)).backward()
    assert max_rel(imgs.grad[:, :, :, ::8, ::8], c['dimgs_pick']) <= TOL
    assert abs(float(imgs.grad.norm()) - c['dimgs_norm']) <= TOL * c['dimgs_norm']


# ---- gradients for inputs that are data on the training path, and the 'jj' sampling backend ------------------
@pytest.mark.parametrize('name', ['masks_m5', 'masks_m16'])
def test_masks_to_layout_mask_and_box_gradients_golden_tol1e5(name):
    """dmasks / dboxes as the reference's autograd produces them through F.grid_sample (tests/golden/input_grads.pt)."""
    from ag2video_b200.layout import masks_to_layout
    c = golden('input_grads.pt')[name]
    vecs, boxes, masks = (c[k].cuda().requires_grad_() for k in ('vecs', 'boxes', 'masks'))
    out = masks_to_layout(vecs, boxes, masks, c['H'])
    (out * c['cot'].cuda()).sum().backward()
    errs = (max_rel(vecs.grad, c['dvecs']), max_rel(boxes.grad, c['dboxes']), max_rel(masks.grad, c['dmasks']))
    print('masks_to_layout %s: dvecs %.2e dboxes %.2e dmasks %.2e' % ((name,) + errs))
    assert max(errs) <= TOL
    # only the inputs that ask for a gradient get one
    v2 = c['vecs'].cuda().requires_grad_()
    masks_to_layout(v2, c['boxes'].cuda(), c['masks'].cuda(), c['H']).sum().backward()
    assert v2.grad is not None


def test_crop_bbox_jj_backend_golden_tol1e5():
    from ag2video_b200.bilinear import crop_bbox
    c = golden('input_grads.pt')['crop_jj']
    feats = c['feats'].cuda().requires_grad_()
    crops = crop_bbox(feats, c['bbox'].cuda(), c['HH'], c['WW'], backend='jj')
    assert max_rel(crops, c['crops']) <= TOL
    (crops * c['cot'].cuda()).sum().backward()
    assert max_rel(feats.grad, c['dfeats']) <= TOL
    with pytest.raises(ValueError):
        crop_bbox(feats, c['bbox'].cuda(), 8, backend='other')

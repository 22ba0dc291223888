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

"""K2 parity on the GPU: boxes_to_layout through the C ABI vs golden vectors
(reference outputs) and vs the CPU oracle.  Pixel support: bit-exact.  Values
and dvecs: <= 1e-5 relative (north_star)."""
import pytest
import torch

from _util import golden, max_rel
from ag2video_b200.config import cater_vocab, synthetic_batch
from oracle import ops as oops

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.mark.parametrize('name', ['mixed32', 'demo64', 'rect', 'avg', 'padbox', 'cater128'])
def test_boxes_to_layout_golden(name):
    from ag2video_b200.layout import boxes_to_layout
    c = golden('layout.pt')[name]
    vecs = c['vecs'].cuda().requires_grad_()
    out = boxes_to_layout(vecs, c['boxes'].cuda(), c['H'], c['W'], pooling=c['pooling'])
    assert out.shape == c['out'].shape
    assert torch.equal(out.cpu() != 0, c['out'] != 0), 'pixel support must be bit-exact'
    assert max_rel(out, c['out']) <= TOL
    (out * c['cot'].cuda()).sum().backward()
    assert max_rel(vecs.grad, c['dvecs']) <= TOL


def test_invalid_pooling_raises():
    from ag2video_b200.layout import boxes_to_layout
    with pytest.raises(ValueError):
        boxes_to_layout(torch.zeros(1, 4).cuda(), torch.ones(1, 4).cuda(), 8, pooling='max')


def test_empty_and_all_zero_boxes():
    from ag2video_b200.layout import boxes_to_layout, boxes_to_layout_batched
    out = boxes_to_layout(torch.randn(3, 8).cuda(), torch.zeros(3, 4).cuda(), 16)
    assert out.shape == (1, 8, 16, 16) and float(out.abs().max()) == 0.0
    out = boxes_to_layout_batched(torch.zeros(2, 0, 8).cuda(), torch.zeros(2, 0, 4).cuda(), None, 16)
    assert out.shape == (2, 8, 16, 16) and float(out.abs().max()) == 0.0


def test_batched_full_size_vs_oracle():
    """BASELINE config 3/4 shape: D=512, 256x256, CATER boxes, masks from
    remove_dummy_objects; two (clip, frame) samples checked against the oracle."""
    from ag2video_b200.layout import boxes_to_layout_batched
    vocab = cater_vocab()
    b = synthetic_batch(B=2, F=2, image_size=8, seed=11, with_images=False)
    B, F, O = b['boxes'].shape[:3]
    D, H = 512, 256
    g = torch.Generator().manual_seed(2)
    vecs = torch.randn(B, F, O, D, generator=g)
    valid = torch.stack([oops.remove_dummy_objects(b['objs'][i], vocab) for i in range(B)])    # [B,O]
    valid_bt = valid.unsqueeze(1).expand(B, F, O).reshape(B * F, O)
    v_gpu = vecs.reshape(B * F, O, D).cuda().requires_grad_()
    out = boxes_to_layout_batched(v_gpu, b['boxes'].reshape(B * F, O, 4).cuda(), valid_bt.cuda(), H)
    cot = torch.randn(B * F, D, H, H, generator=g)
    (out * cot.cuda()).sum().backward()
    for n in (0, 3):
        bi, fi = divmod(n, F)
        v_ref = vecs[bi, fi][valid[bi]].clone().requires_grad_()
        ref = oops.boxes_to_layout(v_ref, b['boxes'][bi, fi][valid[bi]], H)
        assert torch.equal(out[n].cpu() != 0, ref[0] != 0)
        assert max_rel(out[n], ref[0]) <= TOL
        (ref[0] * cot[n]).sum().backward()
        assert max_rel(v_gpu.grad[n][valid[bi].cuda()], v_ref.grad) <= TOL
        assert float(v_gpu.grad[n][~valid[bi].cuda()].abs().max()) == 0.0


def test_linearity_and_determinism_full_size():
    """Size-independent properties at 256x256: layout(a*v1 + v2) == a*layout(v1) + layout(v2)
    within rounding, and two runs are bitwise identical."""
    from ag2video_b200.layout import boxes_to_layout_batched
    b = synthetic_batch(B=2, F=4, image_size=8, seed=5, with_images=False)
    boxes = b['boxes'].reshape(8, -1, 4).cuda()
    O = boxes.shape[1]
    g = torch.Generator().manual_seed(0)
    v1, v2 = torch.randn(8, O, 512, generator=g).cuda(), torch.randn(8, O, 512, generator=g).cuda()
    valid = (boxes[..., 2] > 0) & (boxes[..., 2] < 1)
    f = lambda v: boxes_to_layout_batched(v, boxes, valid, 256)
    a = f(v1)
    assert torch.equal(a, f(v1))
    lhs = f(2.0 * v1 + v2)
    rhs = 2.0 * a + f(v2)
    assert max_rel(lhs, rhs) <= 1e-5
    # checksum: sum over pixels equals sum_o v[o,d] * (sum wy)(sum wx)
    assert torch.isfinite(lhs).all()


@pytest.mark.parametrize('name', ['layout_sum32', 'layout_rect', 'layout_avg64'])
def test_boxes_to_layout_box_gradient_golden_tol1e5(name):
    """dboxes as the reference's autograd produces it through _boxes_to_grid and grid_sample's grid gradient
    (tests/golden/input_grads.pt); an all-zero box (dropped by the forward) gets zero."""
    from ag2video_b200.layout import boxes_to_layout
    c = golden('input_grads.pt')[name]
    vecs, boxes = c['vecs'].cuda().requires_grad_(), c['boxes'].cuda().requires_grad_()
    out = boxes_to_layout(vecs, boxes, c['H'], c['W'], pooling=c['pooling'])
    (out * c['cot'].cuda()).sum().backward()
    e_v, e_b = max_rel(vecs.grad, c['dvecs']), max_rel(boxes.grad, c['dboxes'])
    print('boxes_to_layout %s: dvecs %.2e dboxes %.2e' % (name, e_v, e_b))
    assert e_v <= TOL and e_b <= TOL
    assert float(boxes.grad[1].abs().max()) == 0.0


def test_boxes_to_layout_box_gradient_full_size_vs_oracle_tol1e5():
    """D=128, 64x64, 30 (clip, frame) samples of CATER boxes through the batched entry: dboxes of every sample
    against the CPU oracle's autograd."""
    from ag2video_b200.layout import boxes_to_layout_batched
    batch = synthetic_batch(B=2, F=15, image_size=8, seed=12, with_images=False)
    B, F, O = batch['boxes'].shape[:3]
    g = torch.Generator().manual_seed(3)
    vecs = torch.randn(B * F, O, 128, generator=g)
    cot = torch.randn(B * F, 128, 64, 64, generator=g)
    boxes = batch['boxes'].reshape(B * F, O, 4).clone()
    valid = torch.stack([oops.remove_dummy_objects(batch['objs'][i], cater_vocab()) for i in range(B)])
    valid = valid.unsqueeze(1).expand(B, F, O).reshape(B * F, O)
    bg = boxes.cuda().requires_grad_()
    out = boxes_to_layout_batched(vecs.cuda(), bg, valid.cuda(), 64)
    (out * cot.cuda()).sum().backward()
    worst = 0.0
    for n in range(B * F):
        br = boxes[n][valid[n]].clone().requires_grad_()
        ref = oops.boxes_to_layout(vecs[n][valid[n]], br, 64)
        (ref * cot[n:n + 1]).sum().backward()
        worst = max(worst, max_rel(bg.grad[n][valid[n].cuda()], br.grad))
        assert float(bg.grad[n][~valid[n].cuda()].abs().max() if (~valid[n]).any() else 0.0) == 0.0
    print('boxes_to_layout dboxes, 30 samples of 64x64 D=128 vs CPU oracle: %.2e' % worst)
    assert worst <= TOL

"""End-to-end parity on the GPU (BASELINE config 1 shape): the generator built on
the sm_100a operators vs the golden outputs of the reference's AG2VideoModel at
64x64, batch 2, 4 frames, identical deterministic weights and synthetic clip."""
import pytest
import torch

from _util import det_state, golden, max_rel
from ag2video_b200.config import make_opt, synthetic_batch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _exact_library_convs():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _model(seed, size):
    from ag2video_b200.networks import AG2VideoModel
    m = AG2VideoModel(make_opt(size, batch_size=2))
    m.load_state_dict(det_state(m.state_dict(), seed), strict=True)
    return m.cuda().to(memory_format=torch.channels_last).train()


@pytest.mark.parametrize('mode', ['validation_3xtf32', 'product_tf32'])
def test_generator64_matches_reference_golden(mode):
    """Two passes over the same golden: (1) the SPADE GEMMs in the 3xTF32 validation
    mode (fp32-class products) - this holds the whole pipeline's LOGIC to 1e-4 end to end;
    (2) the product path (TF32 operands, tcgen05): every operator is within 1e-3 on its own
    (tests/test_gpu_spade.py), end to end the rounding of 18 stacked SPADE layers with
    variance-preserving random weights reaches ~1e-2 on single pixels, while the loss
    stays within 1e-3 (the bar SURVEY.md 8d sets for the end-to-end step)."""
    import ag2video_b200.spade as sp
    old = sp.CONV_IMPL
    sp.CONV_IMPL = 3 if mode == 'validation_3xtf32' else 0
    try:
        _generator64(1e-4 if mode == 'validation_3xtf32' else 3e-2)
    finally:
        sp.CONV_IMPL = old


def _generator64(img_tol):
    c = golden('generator64.pt')
    m = _model(c['seed'], 64)
    b = synthetic_batch(B=2, F=4, image_size=64, seed=c['batch_seed'], device='cuda')
    imgs_pred, boxes_pred, flows, conf, _ = m(b['imgs'], b['objs'], b['triplets'], b['actions'],
                                              boxes_gt=b['boxes'], use_gt=True)
    e_img, e_box = max_rel(imgs_pred, c['imgs_pred']), max_rel(boxes_pred, c['boxes_pred'])
    print('imgs %.2e boxes %.2e flows %.2e' % (e_img, e_box, max_rel(flows, c['flows'])))
    assert e_img <= img_tol and e_box <= 1e-3
    loss = (imgs_pred - b['imgs']).abs().mean() + (boxes_pred - b['boxes'])[:, 1:].abs().mean()
    assert abs(float(loss.detach()) - float(c['loss'])) <= 1e-3 * abs(float(c['loss']))
    loss.backward()
    grads = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    # Gradients cross ~40 ReLU / LeakyReLU layers with random weights: a handful of
    # activations within TF32 rounding of zero switch branch (tests/test_gpu_spade.py
    # holds the kink-free cases to 1e-3), so here: relative L2 on sampled slices and
    # the norm of every parameter gradient.
    from _util import rel_l2
    worst = 0.0
    for k, v in c['grad_picks'].items():
        e = rel_l2(grads[k].contiguous().flatten()[:4096], v)
        print('%-70s %.2e' % (k, e))
        worst = max(worst, e)
    assert worst <= 3e-2
    bad = [(k, float(grads[k].norm()), n) for k, n in c['grad_norms'].items()
           if abs(float(grads[k].norm()) - n) > 2e-2 * max(n, 1e-6)]
    assert not bad, bad[:5]

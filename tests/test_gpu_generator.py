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


def _model(seed, size, channels_last=True):
    from ag2video_b200.networks import AG2VideoModel
    m = AG2VideoModel(make_opt(size, batch_size=2, channels_last=channels_last))
    m.load_state_dict(det_state(m.state_dict(), seed), strict=True)
    m = m.cuda().train()
    return m.to(memory_format=torch.channels_last) if channels_last else m


def test_generator64_logic_vs_eager_gpu_reference():
    """The whole pipeline's LOGIC end to end: our generator with the SPADE GEMMs in the
    3xTF32 validation mode (fp32-class products) against the oracle modules run eagerly
    on the same GPU.  Two fp32 implementations of a 40-layer network with batch norms do
    not agree to 1e-6: the eager GPU path itself sits 3.5e-3 (single pixels) from the
    CPU golden.  The check is therefore self-calibrating: our distance to the eager GPU
    path must not exceed twice the eager path's own distance to the CPU golden, and the
    loss must agree to 1e-4."""
    import ag2video_b200.spade as sp
    from oracle import networks as onet
    from _util import rel_l2
    c = golden('generator64.pt')
    old = sp.CONV_IMPL
    sp.CONV_IMPL = 3
    try:
        ref = onet.AG2VideoModel(make_opt(64, batch_size=2))
        ref.load_state_dict(det_state(ref.state_dict(), c['seed']), strict=True)
        ref = ref.cuda().train()
        m = _model(c['seed'], 64, channels_last=False)
        b = synthetic_batch(B=2, F=4, image_size=64, seed=c['batch_seed'], device='cuda')
        args = (b['imgs'], b['objs'], b['triplets'], b['actions'])
        want = ref(*args, boxes_gt=b['boxes'], use_gt=True)
        got = m(*args, boxes_gt=b['boxes'], use_gt=True)
        floor = max_rel(want[0], c['imgs_pred'])
        e_img, e_box = max_rel(got[0], want[0]), max_rel(got[1], want[1])
        print('logic check: imgs %.2e (eager-GPU-vs-CPU floor %.2e, rel-L2 %.2e) boxes %.2e'
              % (e_img, floor, rel_l2(got[0], want[0]), e_box))
        assert e_img <= 2 * floor + 1e-4 and e_box <= 2e-5
        lw = (want[0] - b['imgs']).abs().mean() + (want[1] - b['boxes'])[:, 1:].abs().mean()
        lg = (got[0] - b['imgs']).abs().mean() + (got[1] - b['boxes'])[:, 1:].abs().mean()
        assert abs(float(lg.detach()) - float(lw.detach())) <= 1e-4 * abs(float(lw.detach()))
        lw.backward()
        lg.backward()
        gw = {k: p.grad for k, p in ref.named_parameters() if p.grad is not None}
        gg = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
        # hot-path parameters only (the flow network's gradient passes through a hard
        # threshold); gates within fp32 noise of zero flip, so this is a gross-error check
        hot = [k for k in gw if ('netG' in k or 'acts_to' in k) and float(gw[k].norm()) > 1e-6]
        worst = max((rel_l2(gg[k], gw[k]), k) for k in hot)
        print('logic check: worst hot-path gradient rel-L2 %.2e (%s) over %d tensors' % (worst + (len(hot),)))
        assert worst[0] <= 0.15
    finally:
        sp.CONV_IMPL = old


@pytest.mark.parametrize('mode', ['validation_3xtf32', 'product_tf32'])
def test_generator64_matches_reference_golden(mode):
    """Against the golden produced by the reference on the CPU.  The plain eager GPU path
    (oracle modules on the GPU, fp32) is already 3.5e-3 away from this golden on single
    pixels - cuDNN's fp32 algorithms through 40 layers - so: (1) validation mode must sit
    at that floor; (2) the product path (TF32 operands, tcgen05; every operator within
    1e-3 on its own, tests/test_gpu_spade.py) accumulates the rounding of 18 stacked SPADE
    layers with variance-preserving random weights to ~1e-2 on single pixels, while the
    loss stays within 1e-3, the bar SURVEY.md 8d sets for the end-to-end step."""
    import ag2video_b200.spade as sp
    old = sp.CONV_IMPL
    sp.CONV_IMPL = 3 if mode == 'validation_3xtf32' else 0
    try:
        _generator64(6e-3 if mode == 'validation_3xtf32' else 3e-2)
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
        if float(v.abs().max()) < 1e-6:          # e.g. netG.fc.bias feeds a batch norm: its gradient is 0 in exact arithmetic
            continue
        e = rel_l2(grads[k].contiguous().flatten()[:4096], v)
        print('%-70s %.2e' % (k, e))
        worst = max(worst, e)
    assert worst <= 0.1                           # GPU vs CPU through 40 layers with flipping gates: gross-error check
    bad = [(k, float(grads[k].norm()), n) for k, n in c['grad_norms'].items()
           if n > 1e-4 and abs(float(grads[k].norm()) - n) > 0.1 * n]
    assert not bad, bad[:5]

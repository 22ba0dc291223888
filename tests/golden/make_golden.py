"""Generate the golden fixtures by running the REFERENCE code itself on CPU fp32.

Run in the build container only (needs the read-only checkout at /root/reference):

    python tests/golden/make_golden.py

The reference holds no tests or golden vectors for this path, so
these fixtures are the parity pin: inputs, deterministic weights and the
reference's outputs/gradients, produced with fixed seeds.  Shims (none touches
arithmetic): ``torch.tensor`` module alias for models/utils.py:2, and
``Tensor.cuda`` made a no-op for the hard-coded ``.cuda()`` calls
(generator.py:59-60, utils.py:117-140).
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
REF = os.environ.get('AG2V_REFERENCE', '/root/reference')

shim = types.ModuleType('torch.tensor')
shim.Tensor = torch.Tensor
sys.modules['torch.tensor'] = shim
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
torch.Tensor.cuda = lambda self, *a, **k: self
torch.Tensor.get_device = lambda self: 0

from _util import c4_masks, det_state, det_tensor, dyadic_seg, dyadic_spade_state  # noqa: E402
from ag2video_b200.config import cater_vocab, synthetic_batch  # noqa: E402

from models.graph_models.graph import GraphTripleConv  # noqa: E402
from models.layout import boxes_to_layout, masks_to_layout  # noqa: E402
from models.bilinear import crop_bbox_batch  # noqa: E402
from models.spade_models.networks.normalization import SPADE  # noqa: E402
from models.spade_models.networks.architecture import SPADEResnetBlock  # noqa: E402
from models.graph_models.model import Acts2LayoutModel  # noqa: E402
from models.meta_models import AG2VideoModel  # noqa: E402
from data.args import parser, init_args  # noqa: E402

torch.manual_seed(0)
torch.set_num_threads(8)


def save(name, obj):
    path = os.path.join(HERE, name)
    torch.save(obj, path)
    print('%-22s %8.1f KB' % (name, os.path.getsize(path) / 1024))


def load_det(module, seed):
    module.load_state_dict(det_state(module.state_dict(), seed), strict=True)


def grads_of(module):
    return {k.replace('.module.', '.'): p.grad.clone() for k, p in module.named_parameters()
            if p.grad is not None}


def ref_opt(argv):
    opt = parser.parse_args(argv + ['--gpu_ids', '-1', '--use_cuda', '0', '--no_vgg_loss'])
    opt.vocab = cater_vocab()
    init_args(opt)
    return opt


# ---------------------------------------------------------------- K1 ------
def gconv_case():
    dims = dict(obj_input_dim=24, object_output_dim=16, predicate_input_dim=16,
                predicate_output_dim=16, hidden_dim=32, num_attributes=4)
    layer = GraphTripleConv(**dims)
    load_det(layer, 11)
    B, O, E = 2, 5, 7
    obj = det_tensor('gconv.obj', (B, O, 24), 1).requires_grad_()
    pred = det_tensor('gconv.pred', (B, E, 16), 1).requires_grad_()
    # self loop, repeated nodes, node 4 isolated in clip 0, masked edges
    edges = torch.tensor([[[0, 1], [1, 1], [2, 0], [3, 2], [0, 3], [2, 2], [1, 0]],
                          [[4, 0], [0, 4], [1, 2], [2, 1], [3, 3], [4, 4], [0, 0]]])
    ind = torch.tensor([[1, 1, 0, 1, 1, 0, 1], [1, 0, 1, 1, 1, 1, 0]], dtype=torch.bool)
    new_obj, new_p = layer(obj, pred, edges, ind)
    c1, c2 = det_tensor('gconv.c1', new_obj.shape, 1), det_tensor('gconv.c2', new_p.shape, 1)
    ((new_obj * c1).sum() + (new_p * c2).sum()).backward()
    save('gconv.pt', dict(dims=dims, state=layer.state_dict(), obj=obj.detach(), pred=pred.detach(),
                          edges=edges, ind=ind, new_obj=new_obj.detach(), new_p=new_p.detach(),
                          c1=c1, c2=c2, dobj=obj.grad, dpred=pred.grad, dparams=grads_of(layer)))


def gconv_edge_case():
    """Edge cases of GraphTripleConv.forward run on the reference: no live edge at all (every node pools
    nothing, graph.py:93-99), a single node with self loops, and one live edge among masked ones."""
    dims = dict(obj_input_dim=24, object_output_dim=16, predicate_input_dim=16,
                predicate_output_dim=16, hidden_dim=32, num_attributes=4)
    layer = GraphTripleConv(**dims)
    load_det(layer, 13)
    cases = {}
    specs = {'all_masked': (2, 5, 6, lambda B, E: torch.zeros(B, E, dtype=torch.bool)),
             'single_node': (2, 1, 3, lambda B, E: torch.ones(B, E, dtype=torch.bool)),
             'one_live_edge': (1, 4, 5, lambda B, E: torch.tensor([[0, 0, 1, 0, 0]], dtype=torch.bool))}
    for name, (B, O, E, mk) in specs.items():
        g = torch.Generator().manual_seed(17)
        obj = det_tensor('gedge.obj.' + name, (B, O, 24), 1).requires_grad_()
        pred = det_tensor('gedge.pred.' + name, (B, E, 16), 1).requires_grad_()
        edges = torch.randint(0, O, (B, E, 2), generator=g)
        ind = mk(B, E)
        layer.zero_grad()
        new_obj, new_p = layer(obj, pred, edges, ind)
        c1, c2 = det_tensor('gedge.c1.' + name, new_obj.shape, 1), det_tensor('gedge.c2.' + name, new_p.shape, 1)
        ((new_obj * c1).sum() + (new_p * c2).sum()).backward()
        cases[name] = dict(obj=obj.detach(), pred=pred.detach(), edges=edges, ind=ind, new_obj=new_obj.detach(),
                           new_p=new_p.detach(), c1=c1, c2=c2, dobj=obj.grad.clone(), dpred=pred.grad.clone(),
                           dparams=grads_of(layer))
    save('gconv_edge.pt', dict(dims=dims, state=layer.state_dict(), cases=cases))


def gconv_net_case():
    first = dict(obj_input_dim=64, object_output_dim=16, predicate_input_dim=16,
                 predicate_output_dim=16, hidden_dim=32, num_attributes=4)
    rest = dict(first, obj_input_dim=16)
    layers = torch.nn.ModuleList([GraphTripleConv(**first), GraphTripleConv(**rest), GraphTripleConv(**rest)])
    holder = torch.nn.Module()
    holder.gconvs = layers
    load_det(holder, 12)
    B, O, E = 2, 6, 9
    g = torch.Generator().manual_seed(5)
    obj = det_tensor('net.obj', (B, O, 64), 1).requires_grad_()
    pred = det_tensor('net.pred', (B, E, 16), 1).requires_grad_()
    edges = torch.randint(0, O, (B, E, 2), generator=g)
    ind = torch.rand(B, E, generator=g) > 0.25
    o, p = obj, pred
    for l in layers:
        o, p = l(o, p, edges, ind)
    c1, c2 = det_tensor('net.c1', o.shape, 1), det_tensor('net.c2', p.shape, 1)
    ((o * c1).sum() + (p * c2).sum()).backward()
    save('gconv_net.pt', dict(layers=[first, rest, rest], state=holder.state_dict(), obj=obj.detach(),
                              pred=pred.detach(), edges=edges, ind=ind, new_obj=o.detach(),
                              new_p=p.detach(), c1=c1, c2=c2, dobj=obj.grad, dpred=pred.grad,
                              dparams=grads_of(holder)))


# ---------------------------------------------------------------- K2 ------
def layout_case():
    cases = {}
    boxes_a = torch.tensor([[0.10, 0.20, 0.30, 0.25], [0.00, 0.00, 0.00, 0.00], [0.55, 0.40, 0.219, 0.292],
                            [0.85, 0.80, 0.30, 0.30], [0.50, 0.50, 0.02, 0.02], [-0.10, 0.30, 0.40, 0.20]])
    # literal boxes of the reference's stale demo (models/layout.py:246-252), used as xywh
    boxes_demo = torch.tensor([[0.25, 0.125, 0.5, 0.875], [0, 0, 1, 0.25], [0.6125, 0, 0.875, 1],
                               [0, 0.8, 1, 1.0], [0.25, 0.125, 0.5, 0.875], [0.6125, 0, 0.875, 1]])
    boxes_pad = torch.tensor([[0.2, 0.2, 0.156, 0.208], [-1., -1., -1., -1.], [0.0, 0.0, 1.0, 1.0]])
    specs = [('mixed32', boxes_a, 8, 32, None, 'sum'), ('demo64', boxes_demo, 4, 64, None, 'sum'),
             ('rect', boxes_a, 5, 24, 40, 'sum'), ('avg', boxes_a, 3, 16, None, 'avg'),
             ('padbox', boxes_pad, 4, 32, None, 'sum'), ('cater128', None, 16, 128, None, 'sum')]
    for name, boxes, D, H, W, pooling in specs:
        if boxes is None:
            batch = synthetic_batch(B=1, F=1, image_size=8, seed=3, with_images=False)
            boxes = batch['boxes'][0, 0, :-1].clone()
        vecs = det_tensor('layout.%s' % name, (boxes.shape[0], D), 2).requires_grad_()
        out = boxes_to_layout(vecs, boxes, H, W, pooling=pooling)
        cot = det_tensor('layout.cot.%s' % name, out.shape, 2)
        (out * cot).sum().backward()
        cases[name] = dict(vecs=vecs.detach(), boxes=boxes, H=H, W=W, pooling=pooling,
                           out=out.detach(), cot=cot, dvecs=vecs.grad.clone())
    save('layout.pt', cases)


def masks_case():
    cases = {}
    g = torch.Generator().manual_seed(9)
    boxes = torch.tensor([[0.10, 0.20, 0.30, 0.25], [0.25, 0.30, 0.40, 0.40], [0.55, 0.10, 0.219, 0.292],
                          [0.05, 0.60, 0.50, 0.30]])
    for name, M, H, test_mode in [('m5_train', 5, 32, False), ('m5_test', 5, 32, True),
                                  ('m16_train', 16, 48, False), ('m16_test', 16, 48, True)]:
        masks = (torch.rand(4, M, M, generator=g) > 0.35).float()
        if M == 16:
            masks = masks * torch.rand(4, M, M, generator=g)
        vecs = det_tensor('masks.%s' % name, (4, 6), 3).requires_grad_()
        out = masks_to_layout(vecs, boxes, masks, H, test_mode=test_mode)
        cot = det_tensor('masks.cot.%s' % name, out.shape, 3)
        (out * cot).sum().backward()
        cases[name] = dict(vecs=vecs.detach(), boxes=boxes, masks=masks, H=H, test_mode=test_mode,
                           out=out.detach(), cot=cot, dvecs=vecs.grad.clone())
    save('masks_layout.pt', cases)


def crop_case():
    vocab = cater_vocab()
    batch = synthetic_batch(B=2, F=2, image_size=24, seed=21, n_objects=None)
    imgs = batch['imgs'].clone().requires_grad_()
    boxes = batch['boxes'].clone()
    boxes[0, 1, 1] = 0.0                      # an all-zero box is dropped (bilinear.py:81-83)
    crops, objs_flat = crop_bbox_batch(imgs, batch['objs'], boxes, 8, vocab=vocab)
    cots = [det_tensor('crop.cot.%d' % i, c.shape, 4) for i, c in enumerate(crops)]
    sum((c * k).sum() for c, k in zip(crops, cots)).backward()
    save('crop.pt', dict(imgs=imgs.detach(), objs=batch['objs'], boxes=boxes, HH=8,
                         crops=[c.detach() for c in crops], objs_flat=objs_flat, cots=cots,
                         dimgs=imgs.grad.clone()))


# ---------------------------------------------------------------- K3 ------
def spade_case():
    cases = {}
    for name, C, L, r, Hs, B in [('c16_r8', 16, 8, 8, 32, 2), ('c8_r16', 8, 16, 16, 16, 3),
                                 ('dy_c16_r8', 16, 8, 8, 32, 2), ('dy_c8_r16', 8, 16, 16, 16, 3)]:
        m = SPADE('spadesyncbatch3x3', C, L)
        load_det(m, 31)
        dyadic = name.startswith('dy_')
        if dyadic:          # exact TF32 products: gates agree, gradients can be compared strictly
            m.load_state_dict(dyadic_spade_state(m.state_dict(), 31), strict=True)
        m.train()
        x = det_tensor('spade.x.%s' % name, (B, C, r, r), 5).mul(1.5).add(0.3).requires_grad_()
        seg = (dyadic_seg('spade.seg.%s' % name, (B, L, Hs, Hs), 5, 0.5) if dyadic
               else det_tensor('spade.seg.%s' % name, (B, L, Hs, Hs), 5)).requires_grad_()
        state0 = {k: v.clone() for k, v in m.state_dict().items()}
        out = m(x, seg)
        cot = det_tensor('spade.cot.%s' % name, out.shape, 5)
        (out * cot).sum().backward()
        state1 = {k: v.clone() for k, v in m.state_dict().items()}
        m.eval()
        with torch.no_grad():
            out_eval = m(x, seg)
        cases[name] = dict(C=C, L=L, state=state0, state_after=state1, x=x.detach(), seg=seg.detach(),
                           out=out.detach(), cot=cot, dx=x.grad.clone(), dseg=seg.grad.clone(),
                           dparams=grads_of(m), out_eval=out_eval)
    save('spade.pt', cases)


def block_case():
    cases = {}
    opt = types.SimpleNamespace(norm_G='spectralspadesyncbatch3x3', semantic_nc=8)
    for name, fin, fout in [('b16_8', 16, 8), ('b8_8', 8, 8), ('dy_b16_8', 16, 8), ('dy_b8_8', 8, 8)]:
        m = SPADEResnetBlock(fin, fout, opt)
        load_det(m, 41)
        dyadic = name.startswith('dy_')
        if dyadic:
            m.load_state_dict(dyadic_spade_state(m.state_dict(), 41), strict=True)
        m.train()
        x = det_tensor('block.x.%s' % name, (2, fin, 8, 8), 6).requires_grad_()
        seg = (dyadic_seg('block.seg.%s' % name, (2, 8, 16, 16), 6, 0.5) if dyadic
               else det_tensor('block.seg.%s' % name, (2, 8, 16, 16), 6)).requires_grad_()
        state0 = {k: v.clone() for k, v in m.state_dict().items()}
        out = m(x, seg)
        cot = det_tensor('block.cot.%s' % name, out.shape, 6)
        (out * cot).sum().backward()
        cases[name] = dict(fin=fin, fout=fout, state=state0, x=x.detach(), seg=seg.detach(),
                           out=out.detach(), cot=cot, dx=x.grad.clone(), dseg=seg.grad.clone(),
                           dparams=grads_of(m),
                           state_after={k: v.clone() for k, v in m.state_dict().items()})
    save('spade_block.pt', cases)


# ------------------------------------------------------------ callers -----
def acts2layout_case():
    opt = ref_opt(['--embedding_dim', '16', '--gconv_dim', '16', '--gconv_hidden_dim', '32',
                   '--image_size', '32,32'])
    m = Acts2LayoutModel(opt)
    load_det(m, 51)
    batch = synthetic_batch(B=2, F=4, image_size=32, seed=77, with_images=False)
    obj_vecs, boxes_pred, extra = m(batch['objs'], batch['triplets'], batch['actions'], batch['boxes'])
    c1, c2 = det_tensor('a2l.c1', obj_vecs.shape, 7), det_tensor('a2l.c2', boxes_pred.shape, 7)
    ((obj_vecs * c1).sum() + (boxes_pred * c2).sum()).backward()
    save('acts2layout.pt', dict(over=dict(embedding_dim=16, gconv_dim=16, gconv_hidden_dim=32),
                                seed=51, batch_seed=77, batch={k: v for k, v in batch.items() if v is not None},
                                obj_vecs=obj_vecs.detach(), boxes_pred=boxes_pred.detach(),
                                temporal_triplets=extra[1], rel_t=extra[2], c1=c1, c2=c2,
                                dparams=grads_of(m)))


def generator_case():
    """BASELINE config 1: generator fwd+bwd, 64x64, batch 2, frames_per_action 4."""
    opt = ref_opt(['--image_size', '64,64', '--batch_size', '2'])
    m = AG2VideoModel(opt, torch.device('cpu'))
    load_det(m, 61)
    m.train()
    batch = synthetic_batch(B=2, F=4, image_size=64, seed=99)
    out = m(batch['imgs'], batch['objs'], batch['triplets'], batch['actions'],
            boxes_gt=batch['boxes'], test_mode=False, use_gt=True)
    imgs_pred, boxes_pred, flows, conf, _ = out
    loss = (imgs_pred - batch['imgs']).abs().mean() + (boxes_pred - batch['boxes'])[:, 1:].abs().mean()
    loss.backward()
    grads = grads_of(m)
    picks = ['layout_to_video.netG.up_3.norm_1.mlp_gamma.weight',
             'layout_to_video.netG.head_0.norm_0.mlp_shared.0.weight',
             'layout_to_video.netG.fc.bias',
             'acts_to_objs.gconvs.0.net1.0.weight', 'acts_to_objs.gconvs.2.net2.2.weight',
             'acts_to_boxes.box_net.2.weight', 'layout_to_video.attribute_embedding.att_emb_1.weight']
    save('generator64.pt', dict(seed=61, batch_seed=99, imgs_pred=imgs_pred.detach(),
                                boxes_pred=boxes_pred.detach(), flows=flows.detach(), conf=conf.detach(),
                                loss=loss.detach(),
                                grad_norms={k: float(v.norm()) for k, v in grads.items()},
                                grad_picks={k: grads[k].flatten()[:4096].clone() for k in picks}))


def losses_case():
    """One training iteration's three losses at BASELINE config 1 (64x64, batch 2, 4 frames):
    the reference's MultiscaleActionDiscriminator + LossModel (--no_vgg_loss) on the reference
    generator's output, in the order of scripts/train.py:446-493."""
    from models.spade_models.networks.discriminator import MultiscaleActionDiscriminator
    from models.spade_models.loss_model import LossModel
    opt = ref_opt(['--image_size', '64,64', '--batch_size', '2'])
    m = AG2VideoModel(opt, torch.device('cpu'))
    load_det(m, 61)
    m.train()
    netD = MultiscaleActionDiscriminator(opt)
    load_det(netD, 71)
    netD.train()
    holder = types.SimpleNamespace(img_discriminator=netD)
    lm = LossModel(opt, holder)
    b = synthetic_batch(B=2, F=4, image_size=64, seed=99)
    batch = (b['imgs'], b['objs'], b['boxes'], b['triplets'], b['actions'], None)
    out = m(b['imgs'], b['objs'], b['triplets'], b['actions'], boxes_gt=b['boxes'], test_mode=False, use_gt=True)
    # discriminator forward on the real frames: outputs of both scales, every level
    n = opt.n_frames_G - 1
    with torch.no_grad():
        d_real = netD(b['imgs'][:, n:], b['objs'], b['boxes'][:, n:], [a[:, n:] for a in out[4]])
    G = lm(batch, out, mode='compute_generator_loss')
    G = {k: v.mean() for k, v in G.items()}
    m.zero_grad(); netD.zero_grad()
    G['total_loss'].backward()
    g_grads = grads_of(m)
    netD.zero_grad()
    D = lm(batch, out, mode='compute_discriminator_loss')
    D = {k: v.mean() for k, v in D.items()}
    D['total_img_loss'].backward()
    d_grads = {k: p.grad.clone() for k, p in netD.named_parameters() if p.grad is not None}
    bg = synthetic_batch(B=2, F=16, image_size=64, seed=101, with_images=False)
    boxes_pred = m(None, bg['objs'], bg['triplets'], bg['actions'], boxes_gt=bg['boxes'], test_mode=False, graph_only=True)
    GG = lm((None, bg['objs'], bg['boxes'], bg['triplets'], bg['actions'], None), boxes_pred, mode='compute_graph_loss')
    m.zero_grad()
    GG['total_loss'].backward()
    gg_grads = grads_of(m)
    picks = ['layout_to_video.netG.up_3.norm_1.mlp_gamma.weight', 'layout_to_video.netG.fc.bias',
             'acts_to_objs.gconvs.0.net1.0.weight', 'layout_to_video.flows_network.conv_flow.0.weight']
    save('losses64.pt', dict(
        seed_g=61, seed_d=71, batch_seed=99, graph_batch_seed=101,
        d_real_last=[scale[-1].clone() for scale in d_real],
        d_real_norms=[[float(o.norm()) for o in scale] for scale in d_real],
        d_real_picks=[[o.flatten()[:1024].clone() for o in scale] for scale in d_real],
        G={k: v.detach().clone() for k, v in G.items()}, D={k: v.detach().clone() for k, v in D.items()},
        graph={k: v.detach().clone() for k, v in GG.items()},
        g_grad_norms={k: float(v.norm()) for k, v in g_grads.items()},
        g_grad_picks={k: g_grads[k].flatten()[:4096].clone() for k in picks if k in g_grads},
        d_grad_norms={k: float(v.norm()) for k, v in d_grads.items()},
        d_grad_picks={k: d_grads[k].flatten()[:2048].clone() for k in
                      ['discriminator_0.model0.0.weight', 'discriminator_1.model3.0.0.weight_orig',
                       'gconvs.0.net1.0.weight', 'fc_objs_vecs.weight', 'acts_embeddings.weight']},
        graph_grad_norms={k: float(v.norm()) for k, v in gg_grads.items() if k.startswith('acts_to_boxes')}))


def _iteration_golden(size, B, F, seed_batch, seed_graph, name_gen, name_loss, picks_g, full_grads=()):
    """Generator fwd+bwd (surrogate L1 loss) and one iteration's three losses of the REFERENCE at
    ``size`` x ``size`` on B clips x F frames (graph batch: B clips x 16 frames)."""
    from models.spade_models.networks.discriminator import MultiscaleActionDiscriminator
    from models.spade_models.loss_model import LossModel
    opt = ref_opt(['--image_size', '%d,%d' % (size, size), '--batch_size', str(B)])
    m = AG2VideoModel(opt, torch.device('cpu'))
    load_det(m, 61)
    m.train()
    b = synthetic_batch(B=B, F=F, image_size=size, seed=seed_batch)
    state0 = {k: v.clone() for k, v in m.state_dict().items()}
    out = m(b['imgs'], b['objs'], b['triplets'], b['actions'], boxes_gt=b['boxes'], test_mode=False, use_gt=True)
    imgs_pred, boxes_pred, flows, conf, _ = out
    loss = (imgs_pred - b['imgs']).abs().mean() + (boxes_pred - b['boxes'])[:, 1:].abs().mean()
    loss.backward()
    grads = grads_of(m)
    save(name_gen, dict(seed=61, batch_seed=seed_batch, size=size, B=B, F=F, imgs_pred=imgs_pred.detach(),
                        boxes_pred=boxes_pred.detach(), flows=flows.detach(), conf=conf.detach(), loss=loss.detach(),
                        grad_norms={k: float(v.norm()) for k, v in grads.items()},
                        grad_picks={k: grads[k].flatten()[:4096].clone() for k in picks_g},
                        grad_full={k: grads[k].clone() for k in full_grads}))
    # the three losses of one iteration, on a FRESH copy of the weights / running statistics
    m.load_state_dict(state0, strict=True)
    m.zero_grad()
    netD = MultiscaleActionDiscriminator(opt)
    load_det(netD, 71)
    netD.train()
    lm = LossModel(opt, types.SimpleNamespace(img_discriminator=netD))
    batch = (b['imgs'], b['objs'], b['boxes'], b['triplets'], b['actions'], None)
    out = m(b['imgs'], b['objs'], b['triplets'], b['actions'], boxes_gt=b['boxes'], test_mode=False, use_gt=True)
    G = {k: v.mean() for k, v in lm(batch, out, mode='compute_generator_loss').items()}
    netD.zero_grad()
    G['total_loss'].backward()
    g_grads = grads_of(m)
    netD.zero_grad()
    D = {k: v.mean() for k, v in lm(batch, out, mode='compute_discriminator_loss').items()}
    D['total_img_loss'].backward()
    d_grads = {k: p.grad.clone() for k, p in netD.named_parameters() if p.grad is not None}
    bg = synthetic_batch(B=B, F=16, image_size=size, seed=seed_graph, with_images=False)
    bp = m(None, bg['objs'], bg['triplets'], bg['actions'], boxes_gt=bg['boxes'], test_mode=False, graph_only=True)
    GG = lm((None, bg['objs'], bg['boxes'], bg['triplets'], bg['actions'], None), bp, mode='compute_graph_loss')
    m.zero_grad()
    GG['total_loss'].backward()
    gg_grads = grads_of(m)
    save(name_loss, dict(
        seed_g=61, seed_d=71, batch_seed=seed_batch, graph_batch_seed=seed_graph, size=size, B=B, F=F,
        G={k: v.detach().clone() for k, v in G.items()}, D={k: v.detach().clone() for k, v in D.items()},
        graph={k: v.detach().clone() for k, v in GG.items()}, graph_boxes_pred=bp.detach().clone(),
        g_grad_norms={k: float(v.norm()) for k, v in g_grads.items()},
        d_grad_norms={k: float(v.norm()) for k, v in d_grads.items()},
        graph_grad_norms={k: float(v.norm()) for k, v in gg_grads.items() if k.startswith('acts_to_boxes')},
        graph_grad_full={k: v.clone() for k, v in gg_grads.items()
                         if k in ('acts_to_boxes.box_net.2.weight', 'acts_to_boxes.gconvs.1.net2.2.weight',
                                  'acts_to_boxes.obj_vecs_net.0.weight', 'acts_to_boxes.pred_embeddings.weight',
                                  'acts_to_boxes.acts_embeddings.weight',
                                  'acts_to_boxes.attribute_embedding.attribute_fc_gen.bias')}))


def k2b_c4_case():
    """masks_to_layout / crop_bbox_batch at BASELINE config 4 sizes (10 objects, 256x256; masks 16 and 256
    wide; crops 256 -> 32, B=2, N=4).  Inputs are regenerated from seeds by the tests; D is 2 here to keep
    the fixture small (every channel runs the same code; the full D=512 case is checked against the oracle)."""
    vocab = cater_vocab()
    cases = {}
    batch = synthetic_batch(B=2, F=4, image_size=256, seed=404, n_objects=10)
    boxes = batch['boxes'][0, 0, :10].clone()
    for name, M, test_mode in [('m16_train', 16, False), ('m256_train', 256, False), ('m256_test', 256, True)]:
        masks = c4_masks(name, M)
        vecs = det_tensor('k2bc4.%s' % name, (10, 2), 8).requires_grad_()
        out = masks_to_layout(vecs, boxes, masks, 256, test_mode=test_mode)
        cot = det_tensor('k2bc4.cot.%s' % name, out.shape, 8)
        (out * cot).sum().backward()
        cases[name] = dict(M=M, test_mode=test_mode, boxes=boxes, out=out.detach(), dvecs=vecs.grad.clone())
    imgs = batch['imgs'].clone().requires_grad_()
    crops, objs_flat = crop_bbox_batch(imgs, batch['objs'], batch['boxes'], 32, vocab=vocab)
    cots = [det_tensor('k2bc4.crop.cot.%d' % i, c.shape, 9) for i, c in enumerate(crops)]
    sum((c * k).sum() for c, k in zip(crops, cots)).backward()
    dimgs = imgs.grad
    cases['crop'] = dict(batch_seed=404, crops=[c.detach() for c in crops], objs_flat=objs_flat,
                         dimgs_norm=float(dimgs.norm()), dimgs_pick=dimgs[:, :, :, ::8, ::8].clone())
    save('k2b_c4.pt', cases)


def generator256_case():
    """BASELINE config 3 shapes (256x256) on a bounded sample: 1 clip x 2 frames (one generated frame),
    the reference's generator fwd+bwd and the three losses of one iteration (generator256.pt, losses256.pt)."""
    picks = ['layout_to_video.netG.up_3.norm_1.mlp_gamma.weight',
             'layout_to_video.netG.head_0.norm_0.mlp_shared.0.weight',
             'layout_to_video.netG.up_1.conv_0.weight_orig',
             'acts_to_objs.gconvs.0.net1.0.weight', 'acts_to_objs.gconvs.2.net2.2.weight',
             'acts_to_boxes.box_net.2.weight', 'layout_to_video.attribute_embedding.att_emb_1.weight']
    _iteration_golden(256, 1, 2, 199, 201, 'generator256.pt', 'losses256.pt', picks)


def cater_case():
    """The reference's CATERDataset + collate_fn (data/cater.py, data/dataset_params.py) on the tiny on-disk fixture of
    tests/_cater_fixture.py: test-time window (deterministic) and a seeded training-time window.  scikit-video, which
    the reference imports to decode .avi files, is stubbed: the fixture ships the frame cache the reference reads."""
    import tempfile
    import numpy as np
    sk, skio = types.ModuleType('skvideo'), types.ModuleType('skvideo.io')
    skio.FFmpegReader = object
    sk.io = skio
    sys.modules['skvideo'], sys.modules['skvideo.io'] = sk, skio
    from data.cater import CATERDataset
    from data.dataset_params import collate_fn
    import _cater_fixture
    out = {}
    with tempfile.TemporaryDirectory() as root:
        labels, data_root = _cater_fixture.write(root)
        for mode, is_test in (('test', True), ('train', False)):
            ds = CATERDataset(labels, data_root, is_test=is_test, image_size=(16, 16), frames_per_action=4,
                              initial_frames_per_sample=48)
            np.random.seed(7)
            items = [ds[i] for i in range(len(ds))]
            imgs, objs, boxes, triplets, actions, ids = collate_fn(ds.vocab, items)
            out[mode] = dict(n=len(ds), names=list(ds.vid_names), imgs=imgs, objs=objs, boxes=boxes, triplets=triplets,
                             actions=actions, ids=ids)
    save('cater.pt', out)


def inputgrads_case():
    """The gradients the reference's autograd produces for inputs that are data on the training path - boxes of
    boxes_to_layout, masks and boxes of masks_to_layout - and crop_bbox with the 'jj' sampling backend
    (bilinear.py:127-128).  All from the reference's own functions."""
    from models.bilinear import crop_bbox
    out = {}
    boxes_a = torch.tensor([[0.10, 0.20, 0.30, 0.25], [0.00, 0.00, 0.00, 0.00], [0.55, 0.40, 0.219, 0.292],
                            [0.85, 0.80, 0.30, 0.30], [0.50, 0.50, 0.02, 0.02], [-0.10, 0.30, 0.40, 0.20]])
    for name, D, H, W, pooling in [('sum32', 8, 32, None, 'sum'), ('rect', 5, 24, 40, 'sum'), ('avg64', 16, 64, None, 'avg')]:
        vecs = det_tensor('ig.layout.%s' % name, (boxes_a.shape[0], D), 2).requires_grad_()
        boxes = boxes_a.clone().requires_grad_()
        res = boxes_to_layout(vecs, boxes, H, W, pooling=pooling)
        cot = det_tensor('ig.layout.cot.%s' % name, res.shape, 2)
        (res * cot).sum().backward()
        out['layout_' + name] = dict(vecs=vecs.detach(), boxes=boxes.detach(), H=H, W=W, pooling=pooling, cot=cot,
                                     dvecs=vecs.grad.clone(), dboxes=boxes.grad.clone())
    g = torch.Generator().manual_seed(19)
    boxes_m = torch.tensor([[0.10, 0.20, 0.30, 0.25], [0.25, 0.30, 0.40, 0.40], [0.55, 0.10, 0.219, 0.292],
                            [0.05, 0.60, 0.50, 0.30]])
    for name, M, H in [('m5', 5, 32), ('m16', 16, 48)]:
        masks = (torch.rand(4, M, M, generator=g) * (torch.rand(4, M, M, generator=g) > 0.3).float()).requires_grad_()
        vecs = det_tensor('ig.masks.%s' % name, (4, 6), 3).requires_grad_()
        boxes = boxes_m.clone().requires_grad_()
        res = masks_to_layout(vecs, boxes, masks, H)
        cot = det_tensor('ig.masks.cot.%s' % name, res.shape, 3)
        (res * cot).sum().backward()
        out['masks_' + name] = dict(vecs=vecs.detach(), boxes=boxes.detach(), masks=masks.detach(), H=H, cot=cot,
                                    dvecs=vecs.grad.clone(), dboxes=boxes.grad.clone(), dmasks=masks.grad.clone())
    feats = det_tensor('ig.jj.feats', (5, 3, 20, 28), 4).requires_grad_()
    bbox = torch.tensor([[0.10, 0.20, 0.30, 0.25], [0.0, 0.0, 1.0, 1.0], [0.55, 0.40, 0.219, 0.292],
                         [0.85, 0.80, 0.30, 0.30], [-0.10, 0.30, 0.40, 0.20]])       # two boxes leave the image: clamped taps
    crops = crop_bbox(feats, bbox, 8, 6, backend='jj')
    cot = det_tensor('ig.jj.cot', crops.shape, 4)
    (crops * cot).sum().backward()
    out['crop_jj'] = dict(feats=feats.detach(), bbox=bbox, HH=8, WW=6, crops=crops.detach(), cot=cot, dfeats=feats.grad.clone())
    save('input_grads.pt', out)


if __name__ == '__main__':
    which = sys.argv[1:] or ['gconv', 'gconv_edge', 'gconv_net', 'layout', 'masks', 'crop', 'spade', 'block',
                             'acts2layout', 'generator', 'losses']
    for w in which:
        globals()['%s_case' % w]()

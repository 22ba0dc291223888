"""K1 parity on the GPU: CUDA GraphTripleConv (through the C ABI) vs the golden
vectors from the reference and vs the CPU oracle.  Tolerance: 1e-3 relative
(north_star) is the bar; the 3xTF32 products land near 1e-6, asserted at 2e-5."""
import pytest
import torch

from _util import golden, load_det, max_rel
from ag2video_b200.config import microbench_graph
from oracle import ops as oops

pytestmark = pytest.mark.gpu
TOL = 2e-5


def _to(module_cls, kwargs, state, dev):
    m = module_cls(**kwargs) if isinstance(kwargs, dict) else module_cls(kwargs)
    m.load_state_dict(state, strict=True)
    return m.to(dev)


def _grads(m):
    return {k: p.grad for k, p in m.named_parameters()}


def test_gconv_golden():
    from ag2video_b200.graph import GraphTripleConv
    g = golden('gconv.pt')
    m = _to(GraphTripleConv, g['dims'], g['state'], 'cuda')
    obj = g['obj'].cuda().requires_grad_()
    pred = g['pred'].cuda().requires_grad_()
    new_obj, new_p = m(obj, pred, g['edges'].cuda(), g['ind'].cuda())
    assert max_rel(new_obj, g['new_obj']) <= TOL and max_rel(new_p, g['new_p']) <= TOL
    ((new_obj * g['c1'].cuda()).sum() + (new_p * g['c2'].cuda()).sum()).backward()
    assert max_rel(obj.grad, g['dobj']) <= TOL and max_rel(pred.grad, g['dpred']) <= TOL
    for k, v in g['dparams'].items():
        assert max_rel(_grads(m)[k], v) <= TOL, k


def test_gconv_net_golden():
    from ag2video_b200.graph import GraphTripleConvNet
    g = golden('gconv_net.pt')
    m = _to(GraphTripleConvNet, g['layers'], g['state'], 'cuda')
    obj = g['obj'].cuda().requires_grad_()
    pred = g['pred'].cuda().requires_grad_()
    o, p = m(obj, pred, g['edges'].cuda(), g['ind'].cuda())
    assert max_rel(o, g['new_obj']) <= TOL and max_rel(p, g['new_p']) <= TOL
    ((o * g['c1'].cuda()).sum() + (p * g['c2'].cuda()).sum()).backward()
    assert max_rel(obj.grad, g['dobj']) <= TOL and max_rel(pred.grad, g['dpred']) <= TOL
    for k, v in g['dparams'].items():
        assert max_rel(_grads(m)[k], v) <= TOL, k


@pytest.mark.parametrize('name', ['all_masked', 'single_node', 'one_live_edge'])
def test_gconv_edge_cases_golden(name):
    """Edge cases pinned on the reference itself (tests/golden/gconv_edge.pt)."""
    from ag2video_b200.graph import GraphTripleConv
    g = golden('gconv_edge.pt')
    c = g['cases'][name]
    m = _to(GraphTripleConv, g['dims'], g['state'], 'cuda')
    obj, pred = c['obj'].cuda().requires_grad_(), c['pred'].cuda().requires_grad_()
    o, p = m(obj, pred, c['edges'].cuda(), c['ind'].cuda())
    assert max_rel(o, c['new_obj']) <= TOL and max_rel(p, c['new_p']) <= TOL
    ((o * c['c1'].cuda()).sum() + (p * c['c2'].cuda()).sum()).backward()
    assert max_rel(obj.grad, c['dobj']) <= TOL and max_rel(pred.grad, c['dpred']) <= TOL
    got = _grads(m)
    for k, v in c['dparams'].items():
        assert max_rel(got[k], v) <= TOL, k


@pytest.mark.parametrize('case', ['layer0_c3', 'layer1_c3', 'layer0_c4', 'ragged', 'return_pred',
                                  'all_masked', 'disc_e6', 'single_obj', 'batch8'])
def test_gconv_vs_oracle_full_size(case):
    """BASELINE configs 3 and 4 shapes: Din 512/128, hidden 512, E = 16 or 40."""
    from ag2video_b200.graph import GraphTripleConv
    torch.manual_seed(0)
    Din = 128 if case == 'layer1_c3' else 512
    kw = dict(obj_input_dim=Din, object_output_dim=128, predicate_input_dim=128,
              predicate_output_dim=128, hidden_dim=512, num_attributes=4)
    if case == 'return_pred':
        kw['return_new_p_vecs'] = False
    if case == 'layer0_c4':
        B, O = 2, 11
        edges, ind = microbench_graph(B=B, O=10)
        E = edges.shape[1]
    else:
        # ragged: most edges masked; all_masked: no live edge at all (every node keeps a zero pooled vector);
        # disc_e6: the discriminator's action-only graph (discriminator.py:296-309); single_obj: one node,
        # self loops only; batch8: BASELINE config 2's batch (128 edge rows)
        B, O, E = {'ragged': (3, 7, 13), 'disc_e6': (2, 11, 6), 'single_obj': (2, 1, 3), 'batch8': (8, 11, 16)}.get(case, (2, 11, 16))
        g = torch.Generator().manual_seed(3)
        edges = torch.randint(0, O, (B, E, 2), generator=g)
        ind = torch.rand(B, E, generator=g) > {'ragged': 0.6, 'all_masked': 2.0}.get(case, 0.2)
    ref = load_det(oops.GraphTripleConv(**kw), 5)
    m = GraphTripleConv(**kw)
    m.load_state_dict(ref.state_dict(), strict=True)
    m.cuda()
    g = torch.Generator().manual_seed(4)
    obj_c = torch.randn(B, O, Din, generator=g).requires_grad_()
    pred_c = torch.randn(B, E, 128, generator=g).requires_grad_()
    c1, c2 = torch.randn(B, O, 128, generator=g), torch.randn(B, E, 128, generator=g)
    ro, rp = ref(obj_c, pred_c, edges, ind)
    ((ro * c1).sum() + (rp * c2).sum()).backward()
    obj = obj_c.detach().cuda().requires_grad_()
    pred = pred_c.detach().cuda().requires_grad_()
    go, gp = m(obj, pred, edges.cuda(), ind.cuda())
    ((go * c1.cuda()).sum() + (gp * c2.cuda()).sum()).backward()
    assert max_rel(go, ro) <= TOL and max_rel(gp, rp) <= TOL
    assert max_rel(obj.grad, obj_c.grad) <= TOL and max_rel(pred.grad, pred_c.grad) <= TOL
    rg = _grads(ref)
    for k, v in _grads(m).items():
        assert max_rel(v, rg[k]) <= TOL, k


def test_gconv_is_deterministic():
    """The reference's CUDA scatter_add uses atomics; K1 is specified deterministic."""
    from ag2video_b200.graph import GraphTripleConv
    kw = dict(obj_input_dim=512, object_output_dim=128, predicate_input_dim=128,
              predicate_output_dim=128, hidden_dim=512, num_attributes=4)
    m = load_det(GraphTripleConv(**kw), 2).cuda()
    edges, ind = microbench_graph(B=2, O=10)
    g = torch.Generator().manual_seed(1)
    obj0, pred0 = torch.randn(2, 11, 512, generator=g).cuda(), torch.randn(2, 40, 128, generator=g).cuda()
    outs = []
    for _ in range(2):
        obj, pred = obj0.clone().requires_grad_(), pred0.clone().requires_grad_()
        o, p = m(obj, pred, edges.cuda(), ind.cuda())
        m.zero_grad()
        (o.sum() + p.sum()).backward()
        outs.append([o, p, obj.grad, pred.grad] + [q.grad.clone() for q in m.parameters()])
    for a, b in zip(*outs):
        assert torch.equal(a, b)


def test_gconv_empty_graph_is_an_error():
    """E = 0 never happens on the path (E = objects + actions >= 1); the kernel refuses it loudly."""
    from ag2video_b200.graph import GraphTripleConv
    m = GraphTripleConv(8, 8, 8, 8, 8).cuda()
    with pytest.raises(RuntimeError, match='empty graph'):
        m(torch.zeros(1, 2, 8).cuda(), torch.zeros(1, 0, 8).cuda(), torch.zeros(1, 0, 2, dtype=torch.long).cuda(),
          torch.ones(1, 0, dtype=torch.bool).cuda())


def test_gconv_rejects_cpu_tensors():
    from ag2video_b200.graph import GraphTripleConv
    m = GraphTripleConv(8, 8, 8, 8, 8)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 2, 8), torch.zeros(1, 1, 8), torch.zeros(1, 1, 2, dtype=torch.long),
          torch.ones(1, 1, dtype=torch.bool))

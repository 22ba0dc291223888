import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch
from ag2video_b200 import _lib as L, recurrence
from ag2video_b200.config import make_opt, synthetic_batch
from ag2video_b200.networks import Acts2LayoutModel
from _util import det_state, max_rel
from oracle import networks as onet
opt = make_opt(64)
m = Acts2LayoutModel(opt); m.load_state_dict(det_state(m.state_dict(), 5)); m = m.cuda()
bc = synthetic_batch(B=2, F=16, image_size=8, seed=56, with_images=False)
b = {k: v.cuda() for k, v in bc.items() if v is not None}
ref = onet.Acts2LayoutModel(opt); ref.load_state_dict(det_state(ref.state_dict(), 5))
ro, rb, _ = ref(bc['objs'], bc['triplets'], bc['actions'], bc['boxes'])
(ro.sum() + rb.sum()).backward()
gr = {k: p.grad for k, p in ref.named_parameters() if p.grad is not None}
for core in (0, 1):
    L.lib().ag2v_recur_set_core(core)
    m.zero_grad(set_to_none=True)
    ov, bx, _ = m(b['objs'], b['triplets'], b['actions'], b['boxes'])
    (ov.sum() + bx.sum()).backward()
    worst = max(max_rel(p.grad, gr[k]) for k, p in m.named_parameters() if p.grad is not None and float(gr[k].abs().max()) > 0)
    print('core %d: outputs vs oracle %.2e %.2e, worst grad %.2e' % (core, max_rel(ov, ro), max_rel(bx, rb), worst))
    for _ in range(3): m(b['objs'], b['triplets'], b['actions'], b['boxes'])
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(10):
        with torch.no_grad(): m(b['objs'], b['triplets'], b['actions'], b['boxes'])
    e.record(); torch.cuda.synchronize()
    print('core %d: forward call %.1f us (T=16, B=2, incl. host glue)' % (core, s.elapsed_time(e) * 100))

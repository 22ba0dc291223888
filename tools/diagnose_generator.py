"""Where does the end-to-end difference come from?  Runs on the GPU box:
 (1) the oracle (plain torch) moved to the GPU vs the CPU golden  -> torch GPU-vs-CPU floor
 (2) our generator in every conv mode vs the oracle-on-GPU, stage by stage (forward hooks)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from _util import det_state, golden, max_rel  # noqa: E402
from ag2video_b200.config import make_opt, synthetic_batch  # noqa: E402
from ag2video_b200.networks import AG2VideoModel  # noqa: E402
import ag2video_b200.spade as sp  # noqa: E402
from oracle import networks as onet  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
c = golden('generator64.pt')
opt = make_opt(64, batch_size=2)
b = synthetic_batch(B=2, F=4, image_size=64, seed=c['batch_seed'], device='cuda')


def run(model, tag, store):
    hooks = []
    for name, mod in model.named_modules():
        if name.startswith('layout_to_video.netG.') and name.count('.') == 2 or name in (
                'layout_to_video.conv_dim_in', 'layout_to_video.flows_network', 'acts_to_objs', 'layout_to_video.netG'):
            def hook(m, i, o, name=name):
                o = o[0] if isinstance(o, (tuple, list)) else o
                store.setdefault(name, []).append(o.detach().float().contiguous().clone())
            hooks.append(mod.register_forward_hook(hook))
    with torch.no_grad():
        out = model(b['imgs'], b['objs'], b['triplets'], b['actions'], boxes_gt=b['boxes'], use_gt=True)
    for h in hooks:
        h.remove()
    print('%-28s imgs vs CPU golden %.2e' % (tag, max_rel(out[0], c['imgs_pred'])))
    return out


ref = onet.AG2VideoModel(opt)
ref.load_state_dict(det_state(ref.state_dict(), c['seed']), strict=True)
ref = ref.cuda().train()
ref_store = {}
run(ref, 'oracle on GPU (torch eager)', ref_store)

for impl, tag in [(3, 'ours 3xTF32 validation'), (1, 'ours mma.sync TF32'), (0, 'ours auto (tcgen05)')]:
    sp.CONV_IMPL = impl
    m = AG2VideoModel(opt)
    m.load_state_dict(det_state(m.state_dict(), c['seed']), strict=True)
    m = m.cuda().to(memory_format=torch.channels_last).train()
    store = {}
    run(m, tag, store)
    for name in ref_store:
        if name in store:
            errs = ['%.1e' % max_rel(a, r) for a, r in zip(store[name], ref_store[name])]
            print('    %-40s %s' % (name, ' '.join(errs)))

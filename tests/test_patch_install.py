"""patch.install() against the real reference tree (build container only: skipped where
/root/reference does not exist, e.g. on the GPU box).  Runs in a subprocess because importing the
reference needs process-wide shims (tests/golden/make_golden.py explains them)."""
import os
import subprocess
import sys
import textwrap

import pytest

from _util import ROOT

REF = os.environ.get('AG2V_REFERENCE', '/root/reference')

SCRIPT = textwrap.dedent('''
    import sys, types, torch
    shim = types.ModuleType('torch.tensor'); shim.Tensor = torch.Tensor; sys.modules['torch.tensor'] = shim
    sys.dont_write_bytecode = True
    sys.path.insert(0, %(root)r); sys.path.insert(0, %(ref)r)
    torch.Tensor.cuda = lambda self, *a, **k: self
    from ag2video_b200.config import cater_vocab
    from data.args import parser, init_args
    opt = parser.parse_args(['--image_size', '64,64', '--gpu_ids', '-1', '--use_cuda', '0', '--no_vgg_loss'])
    opt.vocab = cater_vocab(); init_args(opt)
    from models.meta_models import AG2VideoModel
    from models.spade_models.networks.discriminator import MultiscaleActionDiscriminator
    keys_g = list(AG2VideoModel(opt, torch.device('cpu')).state_dict().keys())
    keys_d = list(MultiscaleActionDiscriminator(opt).state_dict().keys())

    import ag2video_b200.patch as ag2v
    done = ag2v.install()
    import ag2video_b200.graph as g, ag2video_b200.spade as s, ag2video_b200.layout as l
    import models.graph_models.model as rm, models.spade_models.networks.generator as rg
    import models.spade_models.networks.discriminator as rd, models.spade_models.networks.spade_generator as rsg
    import models.spade_models.networks.architecture as ra
    assert rm.GraphTripleConv is g.GraphTripleConv and rd.GraphTripleConv is g.GraphTripleConv
    assert rg.boxes_to_layout is l.boxes_to_layout and rd.boxes_to_layout is l.boxes_to_layout
    assert rsg.SPADEResnetBlock is s.SPADEResnetBlock and ra.SPADE is s.SPADE
    m = AG2VideoModel(opt, torch.device('cpu'))          # the reference's own constructors, patched operators inside
    d = MultiscaleActionDiscriminator(opt)
    assert type(m.acts_to_objs.module.gconvs[0]) is g.GraphTripleConv
    assert type(m.layout_to_video.module.netG.up_3) is s.SPADEResnetBlock
    assert type(m.layout_to_video.module.netG.up_3.norm_0) is s.SPADE
    assert type(d.gconvs[1]) is g.GraphTripleConv
    assert list(m.state_dict().keys()) == keys_g, 'generator state-dict keys changed under the patch'
    assert list(d.state_dict().keys()) == keys_d, 'discriminator state-dict keys changed under the patch'
    print('OK', len(done), len(keys_g), len(keys_d))
''')


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'models')), reason='reference checkout not present')
def test_install_rebinds_the_reference_and_keeps_its_state_dict():
    r = subprocess.run([sys.executable, '-c', SCRIPT % dict(root=ROOT, ref=REF)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    assert r.stdout.strip().splitlines()[-1].startswith('OK')

"""CPU-side checks: the C-ABI library loads and exports every symbol that
include/ag2v.h declares (no compute calls without a GPU), the host modules keep
the reference's state-dict layout, and the ops refuse CPU tensors loudly."""
import ctypes
import os
import re

import pytest
import torch

from _util import ROOT, golden
from ag2video_b200.config import make_opt, synthetic_batch


def _declared():
    text = open(os.path.join(ROOT, 'include', 'ag2v.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(ag2v_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    from ag2video_b200 import _lib as L
    import ag2video_b200.spade  # noqa: F401  (registers the K3 signatures)
    import ag2video_b200.bilinear  # noqa: F401
    names = _declared()
    assert len(names) >= 20
    handle = ctypes.CDLL(L.LIB_PATH)
    missing = [n for n in names if not hasattr(handle, n)]
    assert not missing, missing
    undeclared = [n for n in L._SIGNATURES if n not in names]
    assert not undeclared, 'bound in Python but not declared in include/ag2v.h: %s' % undeclared
    lib = L.lib()
    assert lib.ag2v_version() >= 100
    assert isinstance(lib.ag2v_last_error_string(), bytes)


def test_size_queries_need_no_gpu():
    from ag2video_b200 import _lib as L
    lib = L.lib()
    # B=2, O=11, E=16, H=512, Dpo=128: h1 + h2 + pooled + g1 + cnt
    assert lib.ag2v_gcn_layer_saved_floats(2, 11, 16, 512, 128) == 32 * 512 + 32 * 1152 + 2 * 22 * 512 + 22
    assert lib.ag2v_boxes_to_layout_workspace_bytes(8, 11, 256, 256) >= 8 * 11 * 512 * 4


def test_generator_state_dict_matches_reference_layout():
    from ag2video_b200.networks import AG2VideoModel, load_reference_state
    c = golden('generator64.pt')
    m = AG2VideoModel(make_opt(64))
    have, want = {k for k, _ in m.named_parameters()}, set(c['grad_norms'].keys())
    # the golden lists parameters that received a gradient; the occlusion head (conv_w) does not
    assert want <= have and all('.conv_w.' in k for k in have - want)
    fake = {}
    for k, v in m.state_dict().items():
        for sub in ('acts_to_boxes', 'acts_to_objs', 'layout_to_video'):
            if k.startswith(sub + '.'):
                k = k.replace(sub + '.', sub + '.module.', 1)
        fake[k] = v
    load_reference_state(m, fake)


def test_ops_refuse_cpu_tensors():
    from ag2video_b200.layout import boxes_to_layout
    from ag2video_b200.spade import SPADE
    with pytest.raises(RuntimeError):
        boxes_to_layout(torch.zeros(2, 4), torch.ones(2, 4), 8)
    with pytest.raises(RuntimeError):
        SPADE('spadesyncbatch3x3', 8, 4)(torch.zeros(1, 8, 4, 4), torch.zeros(1, 4, 8, 8))


def test_synthetic_batch_contract():
    b = synthetic_batch(B=3, F=4, image_size=16, seed=1)
    B, F, O = b['boxes'].shape[:3]
    assert b['imgs'].shape == (3, 4, 3, 16, 16) and b['objs'].shape == (3, O, 4)
    assert b['triplets'].shape[:2] == (3, 4) and b['actions'].shape[2] == 7
    for i in range(B):
        n = int((b['objs'][i, :, 0] != 0).sum())
        assert torch.equal(b['boxes'][i, 0, n], torch.tensor([0., 0., 1., 1.]))      # the __image__ dummy
        assert (b['boxes'][i, :, n + 1:] == -1).all()                                 # padding rows
        assert (b['triplets'][i, 0, :n, 2] == n).all()


def test_discriminator_state_dict_matches_reference_layout():
    """Same keys and shapes as the reference's MultiscaleActionDiscriminator (the oracle's tree loads
    the reference's state names, tests/test_oracle_golden.py::test_losses64_match_reference)."""
    from ag2video_b200.discriminator import MultiscaleActionDiscriminator
    from oracle import losses as oloss
    opt = make_opt(64)
    ours, ref = MultiscaleActionDiscriminator(opt).state_dict(), oloss.MultiscaleActionDiscriminator(opt).state_dict()
    assert list(ours.keys()) == list(ref.keys())
    assert all(ours[k].shape == ref[k].shape for k in ref)


def test_trainer_partitions_parameters_like_the_reference():
    """optimizer_graph owns acts_to_boxes, optimizer_generator everything else (train.py:365-368)."""
    from ag2video_b200.discriminator import MetaDiscriminatorModel
    from ag2video_b200.losses import LossModel
    from ag2video_b200.networks import AG2VideoModel
    from ag2video_b200.trainer import Trainer
    opt = make_opt(64)
    m, meta = AG2VideoModel(opt), MetaDiscriminatorModel(opt, fused=False)
    tr = Trainer(opt, m, meta, LossModel(opt, meta), fused=False)
    n_graph = sum(p.numel() for p in m.acts_to_boxes.parameters())
    assert sum(p.numel() for p in tr.graph_params) == n_graph
    assert sum(p.numel() for p in tr.gen_params) == sum(p.numel() for p in m.parameters()) - n_graph
    assert len({id(p) for p in tr.gen_params} & {id(p) for p in tr.graph_params}) == 0


def test_trainer_gradient_buckets_follow_the_optimisers():
    """With more than one rank every backward is followed by the all-reduce of exactly the gradients its
    optimiser consumes: three disjoint bucket sets covering generator, discriminator and graph parameters."""
    from ag2video_b200.discriminator import MetaDiscriminatorModel
    from ag2video_b200.losses import LossModel
    from ag2video_b200.networks import AG2VideoModel
    from ag2video_b200.trainer import Trainer
    opt = make_opt(64)
    m, meta = AG2VideoModel(opt), MetaDiscriminatorModel(opt, fused=False)
    tr = Trainer(opt, m, meta, LossModel(opt, meta), world=2, fused=False)
    assert set(tr.buckets) == {'gen', 'd', 'graph'}
    ids = {k: {id(p) for bucket in b.buckets for p in bucket} for k, b in tr.buckets.items()}
    assert ids['gen'] == {id(p) for p in tr.gen_params if p.requires_grad}
    assert ids['graph'] == {id(p) for p in tr.graph_params}
    assert ids['d'] == {id(p) for p in meta.img_discriminator.parameters()}
    assert not (ids['gen'] & ids['d']) and not (ids['gen'] & ids['graph']) and not (ids['d'] & ids['graph'])
    for opt_, key in ((tr.optimizer_generator, 'gen'), (tr.optimizer_graph, 'graph'), (tr.optimizer_d_img, 'd')):
        owned = {id(p) for grp in opt_.param_groups for p in grp['params'] if p.requires_grad}
        assert owned == ids[key], key


def test_loss_model_modes_and_errors():
    """Same dispatch strings as the reference's LossModel.forward (loss_model.py:140-149); the CPU has no
    kernels, so only the graph loss (pure torch) is evaluated here - against the oracle."""
    from ag2video_b200.discriminator import MetaDiscriminatorModel
    from ag2video_b200.losses import LossModel
    from oracle import losses as oloss
    opt = make_opt(64)
    meta = MetaDiscriminatorModel(opt, fused=False)
    lm = LossModel(opt, meta)
    with pytest.raises(ValueError):
        lm({}, None, mode='compute_everything')
    b = synthetic_batch(B=2, F=16, image_size=64, seed=3, with_images=False)
    boxes_pred = b['boxes'] + 0.05 * torch.randn(b['boxes'].shape, generator=torch.Generator().manual_seed(1))
    got = lm(b, boxes_pred, mode='compute_graph_loss')
    want = oloss.LossModel(opt, None).compute_graph_loss(b, boxes_pred)
    assert set(got) == {'bbox_pred', 'total_loss'}
    assert abs(float(got['total_loss']) - float(want['total_loss'])) <= 1e-6 * abs(float(want['total_loss']))
    tup = (None, b['objs'], b['boxes'], b['triplets'], b['actions'], None)       # the reference's batch tuple
    assert torch.equal(lm(tup, boxes_pred, mode='compute_graph_loss')['total_loss'], got['total_loss'])


def test_missing_library_fails_loudly(monkeypatch):
    """Without the built .so every operator raises (no CPU / PyTorch fallback)."""
    from ag2video_b200 import _lib as L
    monkeypatch.setattr(L, '_lib', None)
    monkeypatch.setattr(L, 'LIB_PATH', os.path.join(ROOT, 'ag2video_b200', 'no_such_library.so'))
    with pytest.raises(RuntimeError, match='There is no CPU/PyTorch fallback'):
        L.lib()
    with pytest.raises(RuntimeError):
        L.launch_count()


def test_forced_build_recompiles_every_source():
    """__graft_entry__.build(force=True): the object cache is discarded, every .cu is recompiled for
    sm_100a and the library is relinked (the default build reuses mtime-cached objects)."""
    import time
    import __graft_entry__ as entry
    from ag2video_b200 import build as B
    t0 = time.time() - 1.0
    path = entry.build(force=True)
    objs = [os.path.join(B.OBJ, s[:-3] + '.o') for s in B._sources()]
    assert objs and all(os.path.getmtime(o) >= t0 for o in objs)
    assert os.path.getmtime(path) >= t0
    logs = [open(o + '.ptxas.log').read() for o in objs]
    assert any("Compiling entry function" in l and "sm_100a" in l for l in logs)


def test_binary_is_sm100a_only_and_carries_tcgen05_tma_code():
    """Binary evidence without a GPU (cuobjdump): the shipped library holds sm_100a code only, the tcgen05 convolution
    and weight-gradient kernels contain UTC*MMA (tcgen05.mma), LDTM (tcgen05.ld) and UTMALDG (TMA loads), the persistent
    recurrence kernel contains bulk copies (UBLKCP) and mbarrier SYNCS - the mnemonics B200_PROFILING.md lists."""
    import shutil
    import subprocess
    from ag2video_b200 import _lib as L
    if shutil.which('cuobjdump') is None:
        import pytest
        pytest.skip('cuobjdump not on PATH')
    elfs = subprocess.run(['cuobjdump', '--list-elf', L.LIB_PATH], capture_output=True, text=True, check=True).stdout
    archs = set(re.findall(r'\.(sm_\w+)\.cubin', elfs))
    assert archs == {'sm_100a'}, archs
    sass = subprocess.run(['cuobjdump', '-sass', L.LIB_PATH], capture_output=True, text=True, check=True).stdout
    blocks = dict((m.group(1), m.group(2)) for m in re.finditer(r'Function : (\S+)(.*?)(?=\n\s*Function : |\Z)', sass, flags=re.S))

    def ops(fragment):
        body = ''.join(v for k, v in blocks.items() if fragment in k)
        assert body, 'no kernel named *%s*' % fragment
        return body

    conv, wgrad, recur = ops('conv3x3_tc_kernel'), ops('wgrad3x3_tc_kernel'), ops('recur_fwd_kernel')
    for body, name in ((conv, 'conv'), (wgrad, 'wgrad')):
        assert re.search(r'UTC\w*MMA', body), name + ': no tcgen05.mma'
        assert 'LDTM' in body, name + ': no tcgen05.ld'
        assert 'UTMALDG' in body, name + ': no TMA load'
    assert 'UBLKCP' in recur and 'SYNCS' in recur
    assert 'exchange_kernel' in ''.join(blocks) and 'ce_reduce_kernel' in ''.join(blocks)

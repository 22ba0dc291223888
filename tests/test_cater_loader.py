"""The CATER loader (ag2video_b200/data.py) against the REFERENCE's own CATERDataset +
collate_fn run on the same tiny on-disk fixture (tests/golden/cater.pt, made by `make_golden.py cater`), and against the
tensor contract the training loop and `synthetic_batch` use."""
import numpy as np
import torch

from _util import golden
import _cater_fixture
from ag2video_b200.config import synthetic_batch


def _load(tmp_path, is_test):
    from ag2video_b200.data import CATERDataset, collate_fn
    labels, root = _cater_fixture.write(str(tmp_path))
    ds = CATERDataset(labels, root, is_test=is_test, image_size=(16, 16), frames_per_action=4, initial_frames_per_sample=48)
    np.random.seed(7)
    items = [ds[i] for i in range(len(ds))]
    return ds, collate_fn(ds.vocab, items)


def test_loader_reproduces_the_reference_on_the_fixture(tmp_path):
    g = golden('cater.pt')
    for mode, is_test in (('test', True), ('train', False)):
        ds, (imgs, objs, boxes, triplets, actions, ids) = _load(tmp_path / mode, is_test)
        c = g[mode]
        assert len(ds) == c['n'] and ds.vid_names == c['names']          # skip list and missing videos handled alike
        assert ids == c['ids']
        assert torch.equal(objs, c['objs']) and torch.equal(triplets, c['triplets'])
        assert torch.equal(boxes, c['boxes'])
        assert torch.equal(actions, c['actions'])
        assert imgs.shape == c['imgs'].shape and (imgs - c['imgs']).abs().max() <= 1e-6


def test_loader_contract_matches_synthetic_batch(tmp_path):
    """Same keys, dtypes, ranks and padding conventions as config.synthetic_batch."""
    from ag2video_b200.data import as_batch
    ds, collated = _load(tmp_path, True)
    real = as_batch(collated)
    syn = synthetic_batch(B=2, F=4, image_size=16, seed=1)
    assert set(real) == set(syn)
    for k in syn:
        assert real[k].dtype == syn[k].dtype and real[k].dim() == syn[k].dim(), k
    B, F = real['imgs'].shape[:2]
    assert real['boxes'].shape[:2] == (B, F) and real['triplets'].shape[:2] == (B, F) and real['actions'].shape[2] == 7
    O = real['objs'].shape[1]
    assert real['boxes'].shape[2] == O and real['objs'].shape[2] == 4
    for b in range(B):
        n = int((real['objs'][b].sum(-1) != 0).sum())                          # real objects; then the dummy; then padding
        assert torch.equal(real['boxes'][b, :, n], torch.tensor([0., 0., 1., 1.]).expand(F, 4))
        assert bool((real['boxes'][b, :, n + 1:] == -1).all()) and bool((real['objs'][b, n:] == 0).all())
        assert bool((real['triplets'][b, :, :n, 2] == n).all()) and bool((real['triplets'][b, :, n:, 1] == 7).all())
    pad = real['actions'][:, :, 1] == 6
    assert bool((real['actions'][pad][:, [0, 2, 3, 4, 5, 6]] == 0).all())


def test_missing_frame_cache_is_an_error(tmp_path):
    import shutil
    import pytest
    from ag2video_b200.data import CATERDataset
    labels, root = _cater_fixture.write(str(tmp_path), n_frames=301)
    ds = CATERDataset(labels, root, is_test=True, image_size=(16, 16), frames_per_action=4)
    shutil.rmtree(tmp_path / 'videos' / 'CATER_new_000001')
    (tmp_path / 'videos' / 'CATER_new_000001.avi').write_bytes(b'')
    ds2 = CATERDataset(labels, root, is_test=True, image_size=(16, 16), frames_per_action=4)
    assert ds2.vid_names == ds.vid_names
    with pytest.raises(RuntimeError, match='frame cache'):
        ds2[0]

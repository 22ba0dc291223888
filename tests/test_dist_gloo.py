"""Host-side data-parallel logic on CPU: two gloo ranks, sharding by clip and the
bucketed gradient all-reduce (the NCCL path on the GPU box uses the same code)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ag2video_b200.dist import GradBuckets, shard_clips


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.manual_seed(0)                                # identical replicas
        net = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3), torch.nn.Flatten(), torch.nn.Linear(8 * 36, 5))
        net[0].to(memory_format=torch.channels_last)
        frozen = torch.nn.Parameter(torch.zeros(3), requires_grad=False)
        params = list(net.parameters()) + [frozen]
        # every rank owns different clips
        clips = shard_clips(6, rank, world)
        g = torch.Generator().manual_seed(100)
        data = torch.randn(6, 3, 8, 8, generator=g)
        loss = net(data[clips]).pow(2).mean()
        loss.backward()
        local = [p.grad.clone() for p in net.parameters()]
        buckets = GradBuckets(params, bucket_mb=0)          # one parameter per bucket: exercises the loop
        buckets.allreduce()
        gathered = [torch.zeros_like(torch.cat([l.flatten() for l in local])) for _ in range(world)]
        dist.all_gather(gathered, torch.cat([l.flatten() for l in local]))
        want = sum(gathered) / world
        got = torch.cat([p.grad.flatten() for p in net.parameters()])
        out[rank] = (float((got - want).abs().max()), clips, len(buckets.buckets))
    finally:
        dist.destroy_process_group()


def test_grad_allreduce_and_sharding_two_ranks():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    assert set(res) == {0, 1}
    assert all(err <= 1e-6 for err, _, _ in res.values())
    assert sorted(res[0][1] + res[1][1]) == list(range(6)) and not set(res[0][1]) & set(res[1][1])
    assert res[0][2] >= 4


def test_shard_clips_equal_shards():
    """Every rank gets the same number of clips (ragged shards bias the 1/world gradient average
    and the SyncBN count, an empty shard hangs the collectives)."""
    import pytest
    parts = [shard_clips(7, r, 4) for r in range(4)]                       # drop_last: 7 -> 4 x 1
    assert [len(p) for p in parts] == [1, 1, 1, 1] and len(set(sum(parts, []))) == 4
    parts = [shard_clips(7, r, 4, drop_last=False) for r in range(4)]      # padded by wrapping: 4 x 2
    assert [len(p) for p in parts] == [2, 2, 2, 2] and set(sum(parts, [])) == set(range(7))
    assert [shard_clips(8, r, 2) for r in range(2)] == [[0, 1, 2, 3], [4, 5, 6, 7]]
    with pytest.raises(ValueError):
        shard_clips(2, 3, 4)
    assert shard_clips(2, 3, 4, drop_last=False) == [1]


# ---- Trainer.iteration over two gloo ranks (host logic of the N>1 path, CPU stand-in networks) ----------
class _TinyModel(torch.nn.Module):
    """Same interface as AG2VideoModel as far as Trainer uses it (acts_to_boxes + the rest; graph_only)."""

    def __init__(self):
        super().__init__()
        self.acts_to_boxes = torch.nn.Linear(4, 4)
        self.acts_to_objs = torch.nn.Linear(4, 4)
        self.layout_to_video = torch.nn.Linear(4, 3)

    def forward(self, imgs, objs, triplets, actions, boxes_gt=None, test_mode=False, use_gt=False, graph_only=False):
        boxes_pred = boxes_gt + self.acts_to_boxes(boxes_gt)
        if graph_only:
            return boxes_pred
        imgs_pred = self.layout_to_video(self.acts_to_objs(boxes_gt)).mean(dim=2)        # [B, F, 3]
        return imgs_pred, boxes_pred, None, None, None


class _TinyD(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.img_discriminator = torch.nn.Linear(3, 1)
        self.optimizer_d_img = torch.optim.Adam(self.img_discriminator.parameters(), lr=1e-2, betas=(0.5, 0.999))


class _TinyLosses(torch.nn.Module):
    def __init__(self, d):
        super().__init__()
        self.d = d

    def forward(self, batch, out, mode):
        D = self.d.img_discriminator
        if mode == 'compute_generator_loss':
            return {'total_loss': -D(out[0]).mean() + (out[0] - batch['imgs']).abs().mean()}
        if mode == 'compute_discriminator_loss':
            return {'total_img_loss': torch.relu(1 + D(out[0].detach())).mean() + torch.relu(1 - D(batch['imgs'])).mean()}
        return {'total_loss': torch.nn.functional.smooth_l1_loss(out, batch['boxes'])}


def _tiny_batches(seed, B):
    g = torch.Generator().manual_seed(seed)
    clip = dict(imgs=torch.randn(B, 4, 3, generator=g), objs=None, triplets=None, actions=None, boxes=torch.rand(B, 4, 5, 4, generator=g))
    graph = dict(objs=None, triplets=None, actions=None, boxes=torch.rand(B, 16, 5, 4, generator=g))
    return clip, graph


def _tiny_trainer(world):
    from types import SimpleNamespace
    from ag2video_b200.trainer import Trainer
    torch.manual_seed(0)
    model, d = _TinyModel(), _TinyD()
    opt = SimpleNamespace(learning_rate=1e-2, beta1=0.5)
    return Trainer(opt, model, d, _TinyLosses(d), world=world, fused=False), model, d


def _trainer_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        tr, model, d = _tiny_trainer(world)
        clip, graph = _tiny_batches(7, 4)                      # 4 clips; every rank owns 2 of them
        mine = shard_clips(4, rank, world)
        take = lambda b: {k: (v[mine] if v is not None else None) for k, v in b.items()}
        for _ in range(2):
            tr.iteration(take(clip), take(graph))
        out[rank] = torch.cat([p.detach().flatten() for p in list(model.parameters()) + list(d.parameters())])
    finally:
        dist.destroy_process_group()


def test_trainer_iteration_two_ranks_equals_one_rank_on_all_clips():
    """Clips sharded over two ranks + the three per-optimiser gradient all-reduces give the same parameters
    (on every rank) as one process stepping on all clips: losses are batch means, shards are equal-sized."""
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_trainer_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    tr, model, d = _tiny_trainer(1)
    clip, graph = _tiny_batches(7, 4)
    for _ in range(2):
        tr.iteration(clip, graph)
    want = torch.cat([p.detach().flatten() for p in list(model.parameters()) + list(d.parameters())])
    assert torch.equal(res[0], res[1])
    assert (res[0] - want).abs().max() <= 1e-5 * want.abs().max()


def _collective_choice_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      AG2V_GRAD_ALLREDUCE='ce')
    os.environ.pop('AG2V_PEER_SYNCBN', None)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import warnings
        from ag2video_b200 import peer
        with warnings.catch_warnings(record=True) as seen:
            warnings.simplefilter('always')
            ex = peer.get(None)
        buckets = GradBuckets([torch.nn.Parameter(torch.zeros(4))])
        out[rank] = (ex is None, peer.status(), len(seen), buckets.ce is None)
    finally:
        dist.destroy_process_group()


def test_peer_windows_are_not_used_outside_nccl_groups_and_say_so():
    """The NVLink peer windows (SyncBN sums, copy-engine gradient exchange) exist for NCCL groups with one GPU per rank.
    On a gloo group of CPU ranks both stay off, the process group's own all-reduce is used, and the choice is
    reported (status + one warning on rank 0) instead of being silent."""
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_collective_choice_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    for rank in range(world):
        is_none, status, warned, no_ce = res[rank]
        assert is_none and no_ce
        assert status['collective'] == 'group' and 'gloo' in status['why']
        assert warned == (1 if rank == 0 else 0)

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


def test_shard_clips_ragged():
    parts = [shard_clips(7, r, 4) for r in range(4)]
    assert sorted(sum(parts, [])) == list(range(7))
    assert shard_clips(2, 3, 4) == []

"""Data parallelism: one process per GPU, clips sharded across ranks.

The reference is single-process ``nn.DataParallel`` with a thread-based SyncBN
(models/meta_models.py:16-27, sync_batchnorm/); ``torch.distributed`` is only
consulted to pick a sampler (scripts/train.py:128-133).  Here every rank owns
its clips (no data-path collective); the only exchanges are

* the gradient all-reduce(avg), bucketed so each NCCL call is large enough to
  run at NVLink/NVSwitch bandwidth and launched while later buckets are still
  being flattened, and
* SPADE's batch-norm statistics (``ag2video_b200.spade.set_sync_bn``), 2*C
  doubles per layer, which reproduce the reference's global-batch statistics.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Rendezvous from RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device('cuda', local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def shard_clips(n_clips, rank, world, drop_last=True):
    """Disjoint clip indices for ``rank`` with the SAME count on every rank (DistributedSampler semantics):
    the gradient average is an unweighted 1/world and SyncBN scales the local pixel count by ``world``
    (spade.py), and a rank without work would skip its collectives and hang the others - so shards are
    never ragged.  ``drop_last`` drops the ``n_clips % world`` tail clips; otherwise the tail is padded by
    wrapping around to the first clips (those are then seen twice in the epoch)."""
    if n_clips < world and drop_last:
        raise ValueError('shard_clips: %d clips cannot be sharded over %d ranks without padding' % (n_clips, world))
    if drop_last:
        per = n_clips // world
        return list(range(rank * per, (rank + 1) * per))
    if n_clips <= 0:
        raise ValueError('shard_clips: no clips')
    per = (n_clips + world - 1) // world
    return [i % n_clips for i in range(rank * per, (rank + 1) * per)]


class GradBuckets:
    """Flat gradient buckets in reverse parameter order (the order backward fills
    them).  ``allreduce()`` averages all gradients across ranks."""

    def __init__(self, params, bucket_mb=48, group=None):
        self.group = group
        self.params = [p for p in params if p.requires_grad]
        self.buckets = []
        cur, size, cap = [], 0, bucket_mb * 1024 * 1024 // 4
        for p in reversed(self.params):
            cur.append(p)
            size += p.numel()
            if size >= cap:
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self.flat = [None] * len(self.buckets)

    def allreduce(self):
        if not (dist.is_available() and dist.is_initialized()):
            return
        world = dist.get_world_size(self.group)
        if world == 1:
            return
        works = []
        for i, bucket in enumerate(self.buckets):
            grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in bucket]
            n = sum(g.numel() for g in grads)
            if self.flat[i] is None or self.flat[i].numel() != n:
                self.flat[i] = torch.empty(n, device=grads[0].device, dtype=grads[0].dtype)
            views, off = [], 0
            for g in grads:
                views.append(self.flat[i][off:off + g.numel()].view(g.shape))
                off += g.numel()
            torch._foreach_copy_(views, [g.contiguous() for g in grads])
            works.append((dist.all_reduce(self.flat[i], group=self.group, async_op=True), bucket, views))
        for work, bucket, views in works:
            work.wait()
            for p, v in zip(bucket, views):
                if p.grad is None:
                    p.grad = torch.empty_like(p)
            torch._foreach_copy_([p.grad for p in bucket], views)
            torch._foreach_div_([p.grad for p in bucket], float(world))

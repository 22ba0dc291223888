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
            kw = {}
            if os.environ.get('AG2V_NCCL_HIGH_PRIORITY', '0') == '1':
                # Ablation switch.  The compute stream is full of persistent kernels that hold every SM until they
                # finish, so NCCL's CTAs only get an SM at a kernel boundary; a high-priority NCCL stream was tried to
                # place them first - no measurable change on 2 B200s (52.9 vs 52.8 ms per step), so it stays off.
                opts = dist.ProcessGroupNCCL.Options()
                opts.is_high_priority_stream = True
                kw['pg_options'] = opts
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device('cuda', local), **kw)
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
    """Gradient all-reduce(avg) in flat buckets, overlapped with the backward pass.

    Parameters are grouped in reverse order (the order backward produces their gradients) into buckets of
    ``bucket_mb``.  Every bucket owns a persistent flat buffer; a post-accumulate hook on each parameter counts the
    bucket's gradients as they appear, and the moment the last one lands the bucket is copied into its buffer (one
    multi-tensor copy), ``p.grad`` is re-pointed at the buffer's views (same strides as the parameter, so the fused
    optimiser reads the reduced values in place: no copy back) and an asynchronous all-reduce with the AVG reduction
    starts (no separate divide) - while the rest of the backward keeps running.  ``allreduce()`` after the backward
    launches whatever did not complete through the hooks (parameters without a gradient contribute zeros), then makes
    the current stream wait for all of them.  Everything is stream-ordered, so it can be captured in a CUDA graph.
    """

    def __init__(self, params, bucket_mb=48, group=None, overlap=True):
        bucket_mb = int(os.environ.get('AG2V_BUCKET_MB', bucket_mb))              # ablation switches
        overlap = overlap and os.environ.get('AG2V_GRAD_OVERLAP', '1') != '0'
        self.launched_by_hook = self.launched_at_end = 0
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
        # AG2V_GRAD_ALLREDUCE=ce: the buckets live in a CUDA-IPC arena and are all-reduced by copy engines + a reduction
        # kernel (peer.GradientExchange, csrc/k8_peer.cu) instead of NCCL; default: NCCL
        self.ce = None
        mode = os.environ.get('AG2V_GRAD_ALLREDUCE', 'nccl')
        if mode in ('ce', 'ce-force') and self._active() and self.params and self.params[0].is_cuda and \
                (dist.get_backend(group) == 'nccl' or mode == 'ce-force'):
            from . import peer
            self.ce = peer.GradientExchange([sum(p.numel() for p in b) for b in self.buckets], group)
        self.flat = [None] * len(self.buckets)
        self.views = [None] * len(self.buckets)
        self._ready = [0] * len(self.buckets)
        self._launched = [False] * len(self.buckets)
        self._works = []
        self._hooks = []
        if overlap:
            for i, bucket in enumerate(self.buckets):
                for p in bucket:
                    self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(i)))

    def _active(self):
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _make_hook(self, i):
        def hook(param):
            if not self._active():
                return
            self._ready[i] += 1
            if self._ready[i] == len(self.buckets[i]) and not self._launched[i]:
                self.launched_by_hook += 1
                self._launch(i)
        return hook

    def _views(self, i):
        if self.views[i] is None:
            bucket = self.buckets[i]
            n = sum(p.numel() for p in bucket)
            if self.ce is not None:
                if bucket[0].dtype != torch.float32:
                    raise RuntimeError('copy-engine gradient exchange: float32 parameters only')
                self.flat[i] = self.ce.bucket(i)                       # padded to whole slices; the pad stays zero
            else:
                self.flat[i] = torch.zeros(n, device=bucket[0].device, dtype=bucket[0].dtype)
            views, off = [], 0
            for p in bucket:
                dense = p.is_contiguous() or p.is_contiguous(memory_format=torch.channels_last)
                v = self.flat[i][off:off + p.numel()]
                views.append(v.as_strided(p.shape, p.stride()) if dense else v.view(p.shape))
                off += p.numel()
            self.views[i] = views
        return self.views[i]

    def _launch(self, i):
        bucket, views = self.buckets[i], self._views(i)
        have = [(v, p.grad) for v, p in zip(views, bucket) if p.grad is not None and p.grad.data_ptr() != v.data_ptr()]
        missing = [v for v, p in zip(views, bucket) if p.grad is None]
        if self.ce is not None:
            self.ce.before_fill(i)             # nobody reads the previous contents of this bucket any more
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        if missing:
            torch._foreach_zero_(missing)
        for v, p in zip(views, bucket):
            p.grad = v
        world = dist.get_world_size(self.group)
        if self.ce is not None:
            self.ce.exchange(i)                # on the exchange stream, after everything enqueued so far
            self._launched[i] = True
            return
        if dist.get_backend(self.group) == 'nccl':
            work = dist.all_reduce(self.flat[i], op=dist.ReduceOp.AVG, group=self.group, async_op=True)
            self._works.append((work, None))
        else:                                              # gloo (CPU tests) has no AVG
            work = dist.all_reduce(self.flat[i], group=self.group, async_op=True)
            self._works.append((work, (self.flat[i], float(world))))
        self._launched[i] = True

    def begin(self):
        """Call right before the backward whose gradients this object reduces (after ``zero_grad``): forgets launches
        triggered by someone else's backward through these parameters (their collectives are waited for first, so every
        rank keeps issuing the same sequence)."""
        for work, _ in self._works:
            work.wait()
        if self.ce is not None:
            self.ce.finish()
        self._works = []
        self._ready = [0] * len(self.buckets)
        self._launched = [False] * len(self.buckets)

    def allreduce(self):
        if not self._active():
            return
        for i in range(len(self.buckets)):
            if not self._launched[i]:
                self.launched_at_end += 1
                self._launch(i)
        for work, post in self._works:
            work.wait()
            if post is not None:
                post[0].div_(post[1])
        if self.ce is not None:
            self.ce.finish()
        self._works = []
        self._ready = [0] * len(self.buckets)
        self._launched = [False] * len(self.buckets)

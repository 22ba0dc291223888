"""SyncBN statistics exchange over NVLink peer memory (csrc/k8_peer.cu) - the all-reduce of the per-channel sums of
sync_batchnorm/batchnorm.py:74-83,105-145, one kernel per rank instead of an NCCL launch.

``PeerExchange`` is the plumbing: every rank allocates its window, the 64-byte CUDA IPC handles travel through
``torch.distributed`` (any backend), every rank maps the others' windows.  ``allreduce`` is then ONE launch of this
library on the caller's stream.  All ranks must live on one node (one process per GPU; two processes sharing a GPU also
work and are what the single-GPU test uses) and must issue the same calls in the same order.

``get(group)`` returns the exchange of a process group, creating it on first use, or None where it cannot exist (not
NCCL, several nodes, IPC refused by the platform) - the callers then use the group's own all-reduce; which one runs is
reported by ``status()`` and printed once, never silent.
"""
import ctypes
import os
import socket
import warnings

import torch

from . import _lib as L

L.register('ag2v_peer_window_bytes', L.c_sz, [L.c_i, L.c_i])
L.register('ag2v_peer_window_alloc', L.c_i, [L.c_i, L.c_i, ctypes.POINTER(ctypes.c_void_p)])
L.register('ag2v_peer_window_free', L.c_i, [L.c_p])
L.register('ag2v_peer_window_export', L.c_i, [L.c_p, ctypes.c_char_p])
L.register('ag2v_peer_window_import', L.c_i, [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)])
L.register('ag2v_peer_window_close', L.c_i, [L.c_p])
L.register('ag2v_peer_allreduce_f64', L.c_i, [L.c_p, L.c_i, ctypes.POINTER(ctypes.c_void_p), L.c_i, L.c_i, L.c_i, L.c_p])

L.register('ag2v_peer_alloc', L.c_i, [L.c_sz, ctypes.POINTER(ctypes.c_void_p)])
L.register('ag2v_ce_flag_bytes', L.c_sz, [])
L.register('ag2v_peer_memcpy', L.c_i, [L.c_p, L.c_p, L.c_sz, L.c_p])
L.register('ag2v_ce_sync', L.c_i, [ctypes.POINTER(ctypes.c_void_p)] + [L.c_i] * 8 + [L.c_p])
L.register('ag2v_ce_reduce', L.c_i, [L.c_p, L.c_p, L.c_i, ctypes.c_longlong, L.c_f, L.c_i, L.c_p])

CAP = 16384            # doubles per call: 5 * C * groups of the widest SPADE layer (C = 1024) with room to spare
CHANNELS = 2           # 0: calls on the caller's stream; 1: calls overlapped with other work on a side stream


class _Pending:
    def __init__(self, stream):
        self.stream = stream

    def wait(self):
        torch.cuda.current_stream().wait_stream(self.stream)


class PeerExchange:
    def __init__(self, group=None, cap=CAP):
        import torch.distributed as dist
        self.group, self.cap = group, int(cap)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        lib = L.lib()
        self.own, self.windows = [], []
        self._imported = []
        handles = []
        for _ in range(CHANNELS):
            ptr = ctypes.c_void_p()
            L.check(lib.ag2v_peer_window_alloc(self.world, self.cap, ctypes.byref(ptr)))
            self.own.append(ptr)
            buf = ctypes.create_string_buffer(64)
            L.check(lib.ag2v_peer_window_export(ptr, buf))
            handles.append(buf.raw)
        mine = dict(host=socket.gethostname(), pid=os.getpid(), handles=handles)
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        error = None
        if len({e['host'] for e in everyone}) != 1:
            error = 'ranks on several hosts'
        else:
            try:
                for c in range(CHANNELS):
                    arr = (ctypes.c_void_p * self.world)()
                    for r, e in enumerate(everyone):
                        if r == self.rank:
                            arr[r] = self.own[c]
                            continue
                        ptr = ctypes.c_void_p()
                        L.check(lib.ag2v_peer_window_import(e['handles'][c], ctypes.byref(ptr)))
                        self._imported.append(ptr)
                        arr[r] = ptr
                    self.windows.append(arr)
            except RuntimeError as exc:            # e.g. IPC refused between these two processes
                error = str(exc)
        errors = [None] * self.world
        dist.all_gather_object(errors, error, group=group)     # every rank takes the same decision
        errors = [e for e in errors if e]
        if errors:
            self.close()
            raise RuntimeError('peer windows unavailable: ' + errors[0])
        self.side = torch.cuda.Stream()
        dist.barrier(group=group)                              # every window is zeroed and mapped before the first call

    def allreduce(self, vec, channel=0):
        """In-place sum over the ranks of a contiguous float64 CUDA tensor, on the current stream."""
        if vec.dtype != torch.float64 or not vec.is_cuda or not vec.is_contiguous():
            raise ValueError('peer all-reduce takes a contiguous float64 CUDA tensor')
        n = vec.numel()
        if n > self.cap:
            raise ValueError('peer all-reduce: %d values exceed the window capacity %d' % (n, self.cap))
        # the call number lives in the window (device side): the launch is CUDA-graph capturable
        L.check(L.lib().ag2v_peer_allreduce_f64(L.ptr(vec), n, self.windows[channel], self.rank, self.world, self.cap,
                                                L.stream()))
        return vec

    def allreduce_async(self, vec):
        """The same exchange on a side stream, so that kernels enqueued next on the current stream run while the sums
        travel; ``.wait()`` on the result orders the current stream after it (the host does not block)."""
        self.side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.side):
            self.allreduce(vec, channel=1)
        return _Pending(self.side)

    def close(self):
        lib = L.lib()
        for p in self._imported:
            lib.ag2v_peer_window_close(p)
        for p in self.own:
            lib.ag2v_peer_window_free(p)
        self._imported, self.own, self.windows = [], [], []


_EXCHANGES = {}
_STATUS = {'collective': 'none', 'why': 'single rank'}


def status():
    """Which collective carries the SyncBN sums: {'collective': 'peer' | 'group' | 'none', 'why': ...}."""
    return dict(_STATUS)


def get(group=None):
    """The PeerExchange of `group`, or None when the group's own all-reduce has to carry the sums."""
    import torch.distributed as dist
    key = id(group)
    if key in _EXCHANGES:
        return _EXCHANGES[key]
    ex, why = None, None
    mode = os.environ.get('AG2V_PEER_SYNCBN', '1')          # '0': never; 'force': also for non-NCCL groups (tests)
    if mode == '0':
        why = 'AG2V_PEER_SYNCBN=0'
    elif dist.get_backend(group) != 'nccl' and mode != 'force':
        why = 'backend %s (peer windows are set up for NCCL groups: one process per GPU)' % dist.get_backend(group)
    else:
        try:
            ex = PeerExchange(group)
        except RuntimeError as exc:
            why = str(exc)
    _EXCHANGES[key] = ex
    _STATUS.update(collective='peer' if ex is not None else 'group', why=why or 'CUDA IPC windows over NVLink')
    if ex is None and dist.get_rank(group) == 0:
        warnings.warn('SyncBN statistics travel by the process group\'s all-reduce, not by peer memory: %s' % why)
    return ex


class _External:
    """Device memory of this library seen as a 1-D float32 array (``torch.as_tensor`` maps it without a copy)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {'shape': (int(n),), 'typestr': '<f4', 'data': (int(ptr), False), 'version': 2}


class GradientExchange:
    """All-reduce(avg) of gradient buckets with the copy engines doing the transport (csrc/k8_peer.cu, K8b).

    Every rank owns one exportable arena: the flat buckets (what ``p.grad`` points into), a staging area of
    (world - 1) slices per bucket, and a flag block.  ``bucket(i)`` is the torch view of flat bucket i;
    ``before_fill(i)`` goes on the compute stream right before the bucket is overwritten; ``exchange(i)`` enqueues the
    exchange of bucket i on the exchange stream (after everything enqueued so far on the current stream)."""

    def __init__(self, sizes, group=None, reduce_ctas=64):
        import torch.distributed as dist
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if len(sizes) > 64:
            raise RuntimeError('gradient exchange: at most 64 buckets (got %d)' % len(sizes))
        quantum = 4 * self.world                                  # slices of whole float4s
        self.sizes = [(n + quantum - 1) // quantum * quantum for n in sizes]
        self.slice = [n // self.world for n in self.sizes]
        self.reduce_ctas = int(reduce_ctas)
        self.flat_off, self.stage_off, off = [], [], 0
        for n in self.sizes:
            self.flat_off.append(off)
            off += (n + 63) // 64 * 64
        for sl in self.slice:
            self.stage_off.append(off)
            off += ((self.world - 1) * sl + 63) // 64 * 64
        self.flag_off_bytes = off * 4
        lib = L.lib()
        total = self.flag_off_bytes + lib.ag2v_ce_flag_bytes()
        ptr = ctypes.c_void_p()
        L.check(lib.ag2v_peer_alloc(total, ctypes.byref(ptr)))
        self.own = ptr
        buf = ctypes.create_string_buffer(64)
        L.check(lib.ag2v_peer_window_export(ptr, buf))
        mine = dict(host=socket.gethostname(), handle=buf.raw, sizes=self.sizes)
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        error, self._imported, self.base = None, [], [None] * self.world
        if len({e['host'] for e in everyone}) != 1:
            error = 'ranks on several hosts'
        elif any(e['sizes'] != self.sizes for e in everyone):
            error = 'ranks disagree on the bucket sizes'
        else:
            try:
                for r, e in enumerate(everyone):
                    if r == self.rank:
                        self.base[r] = ptr.value
                        continue
                    q = ctypes.c_void_p()
                    L.check(lib.ag2v_peer_window_import(e['handle'], ctypes.byref(q)))
                    self._imported.append(q)
                    self.base[r] = q.value
            except RuntimeError as exc:
                error = str(exc)
        errors = [None] * self.world
        dist.all_gather_object(errors, error, group=group)
        errors = [e for e in errors if e]
        if errors:
            self.close()
            raise RuntimeError('gradient exchange windows unavailable: ' + errors[0])
        self.flags = (ctypes.c_void_p * self.world)(*[b + self.flag_off_bytes for b in self.base])
        self.stream = torch.cuda.Stream()
        self._pending = False
        self._keep = [_External(self.base[self.rank] + 4 * o, n) for o, n in zip(self.flat_off, self.sizes)]
        self._buckets = [torch.as_tensor(e, device=torch.device('cuda', torch.cuda.current_device())) for e in self._keep]
        dist.barrier(group=group)

    def bucket(self, i):
        return self._buckets[i]

    def _sync(self, i, phase, bump, post, wait, back):
        L.check(L.lib().ag2v_ce_sync(self.flags, self.rank, self.world, i, phase, bump, post, wait, back, L.stream()))

    def before_fill(self, i):
        """Count the exchange that starts with this refill and wait until no rank reads the bucket's previous
        contents any more (phase C of the previous exchange).  On the current (compute) stream."""
        self._sync(i, 2, 1, 0, 1, 1)

    def exchange(self, i):
        lib, me, W = L.lib(), self.rank, self.world
        sl, fo, so = self.slice[i], self.flat_off[i], self.stage_off[i]
        peers = [r for r in range(W) if r != me]
        self.stream.wait_stream(torch.cuda.current_stream())
        self._pending = True
        with torch.cuda.stream(self.stream):
            st = L.stream()
            self._sync(i, 0, 0, 1, 1, 0)                                     # A: every rank's bucket is complete
            for k, r in enumerate(peers):                                    # my slice of every peer's bucket -> staging
                L.check(lib.ag2v_peer_memcpy(self.base[me] + 4 * (so + k * sl), self.base[r] + 4 * (fo + me * sl), 4 * sl, st))
            L.check(lib.ag2v_ce_reduce(self.base[me] + 4 * (fo + me * sl), self.base[me] + 4 * so, W - 1, sl, 1.0 / W,
                                       self.reduce_ctas, st))
            self._sync(i, 1, 0, 1, 1, 0)                                     # B: every rank's slice is reduced
            for r in peers:                                                  # the reduced slices of the others
                L.check(lib.ag2v_peer_memcpy(self.base[me] + 4 * (fo + r * sl), self.base[r] + 4 * (fo + r * sl), 4 * sl, st))
            self._sync(i, 2, 0, 1, 0, 0)                                     # C: I no longer read the peers' buckets

    def finish(self):
        """Order the current stream after every exchange enqueued so far.  (Only when there is one: inside a CUDA-graph
        capture a wait on the idle exchange stream would be a dependency from outside the capture.)"""
        if self._pending:
            torch.cuda.current_stream().wait_stream(self.stream)
            self._pending = False

    def close(self):
        lib = L.lib()
        for q in getattr(self, '_imported', []):
            lib.ag2v_peer_window_close(q)
        if getattr(self, 'own', None) is not None:
            lib.ag2v_peer_window_free(self.own)
        self._imported, self.own = [], None

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

"""ctypes binding of libag2v_sm100a.so (the C ABI declared in include/ag2v.h).

There is no CPU or PyTorch fallback: if the shared library is missing, or an op
is handed a non-CUDA tensor, the call raises.  C entry points return 0 or a
negative code; ``check`` turns the code into a RuntimeError carrying
``ag2v_last_error_string()`` (errors are never swallowed, unlike the reference's
training loop, scripts/train.py:466-468).
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libag2v_sm100a.so')
_lib = None

c_p = ctypes.c_void_p
c_i = ctypes.c_int
c_f = ctypes.c_float
c_sz = ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/ag2v.h one to one
_SIGNATURES = {
    'ag2v_last_error_string': (ctypes.c_char_p, []),
    'ag2v_version': (c_i, []),
    'ag2v_check_device': (c_i, []),
    'ag2v_sm_count': (c_i, []),
    'ag2v_launch_count': (ctypes.c_ulonglong, []),
    'ag2v_gcn_layer_saved_floats': (c_sz, [c_i] * 5),
    'ag2v_gcn_layer_bwd_workspace_floats': (c_sz, [c_i] * 7),
    'ag2v_gcn_layer_fwd': (c_i, [c_p] * 12 + [c_i] * 8 + [c_p] * 3 + [c_p]),
    'ag2v_gcn_layer_bwd': (c_i, [c_p] * 12 + [c_i] * 8 + [c_p] * 11 + [c_p]),
    'ag2v_boxes_to_layout_workspace_bytes': (c_sz, [c_i] * 4),
    'ag2v_boxes_to_layout_fwd': (c_i, [c_p] * 5 + [c_i] * 6 + [c_p] * 2 + [c_p]),
    'ag2v_boxes_to_layout_bwd': (c_i, [c_p] * 5 + [c_i] * 7 + [c_p] * 2 + [c_p]),
}


def register(name, restype, argtypes):
    _SIGNATURES[name] = (restype, argtypes)
    if _lib is not None:
        fn = getattr(_lib, name)
        fn.restype, fn.argtypes = restype, argtypes


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                'ag2video_b200: %s is missing. Build it with `python -m ag2video_b200.build` '
                '(nvcc, sm_100a). There is no CPU/PyTorch fallback.' % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in _SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError if the .so lacks a declared symbol
            fn.restype, fn.argtypes = restype, argtypes
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().ag2v_last_error_string()
        raise RuntimeError('libag2v_sm100a error %d: %s' % (rc, msg.decode() if msg else '?'))


def ptr(t):
    return None if t is None else c_p(t.data_ptr())


def stream():
    return c_p(torch.cuda.current_stream().cuda_stream)


def launch_count():
    return int(lib().ag2v_launch_count())


def need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError('ag2video_b200 operators run on CUDA tensors only (sm_100a kernels, '
                               'no CPU fallback); got a %s tensor' % t.device)


def f32c(t):
    """fp32, contiguous (a no-op for the layouts the callers already produce)."""
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


# ---- per-launch timing (bench.py / tools): CUDA events around a launch on the current stream ----------
PROFILE = None          # set to a list to record (kind, work, start_event, end_event, tag) for every timed launch


class timed:
    """``with timed(kind, work, tag):`` around one kernel launch (or one operator call).  ``work`` =
    algorithmic FLOPs (GEMM-shaped kernels) or algorithmic bytes (HBM-bound kernels) of the launch.
    A no-op unless ``PROFILE`` is a list (events cannot be recorded inside a replayed CUDA graph, so the
    bench measures them in a separate eager pass over the same steps)."""

    def __init__(self, kind, work, tag=None):
        self.kind, self.work, self.tag = kind, work, tag

    def __enter__(self):
        if PROFILE is not None:
            self.s = torch.cuda.Event(enable_timing=True)
            self.e = torch.cuda.Event(enable_timing=True)
            self.s.record()

    def __exit__(self, *a):
        if PROFILE is not None:
            self.e.record()
            PROFILE.append((self.kind, self.work, self.s, self.e, self.tag))

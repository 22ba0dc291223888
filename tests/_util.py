"""Shared test helpers: deterministic weights, error metrics, fixture paths."""
import os
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def _rng(name, seed):
    return np.random.RandomState((zlib.crc32(name.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)


def det_tensor(name, shape, seed=0, scale=1.0):
    """numpy's frozen legacy generator: identical on every machine."""
    return torch.from_numpy((_rng(name, seed).standard_normal(tuple(shape)) * scale).astype(np.float32))


def det_state(state, seed=0):
    """Deterministic values for a state_dict, keyed by the parameter name with any
    DataParallel '.module.' infix removed, so the reference tree and ours agree."""
    out = {}
    for key, ref in state.items():
        name = key.replace('.module.', '.')
        if not torch.is_floating_point(ref):
            out[key] = ref.clone()
            continue
        shape = tuple(ref.shape)
        leaf = name.rsplit('.', 1)[-1]
        if leaf == 'running_var':
            val = det_tensor(name, shape, seed).abs() + 0.5
        elif leaf == 'running_mean':
            val = det_tensor(name, shape, seed, 0.1)
        elif leaf in ('weight_u', 'weight_v'):
            val = det_tensor(name, shape, seed)
            val = val / val.norm().clamp_min(1e-12)
        elif leaf == 'bias':
            val = det_tensor(name, shape, seed, 0.05)
        elif len(shape) == 1:                       # affine norm weight
            val = 1.0 + det_tensor(name, shape, seed, 0.1)
        else:
            fan_in = int(np.prod(shape[1:]))
            val = det_tensor(name, shape, seed, (2.0 / max(fan_in, 1)) ** 0.5)
        out[key] = val
    return out


def load_det(module, seed=0):
    module.load_state_dict(det_state(module.state_dict(), seed), strict=True)
    return module


def max_rel(a, b):
    """max|a-b| / max|b| — the per-tensor metric of SURVEY.md section 8(d)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    den = b.abs().max().item()
    return (a - b).abs().max().item() / (den if den > 0 else 1.0)


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    den = b.norm().item()
    return (a - b).norm().item() / (den if den > 0 else 1.0)


def golden(name):
    return torch.load(os.path.join(GOLDEN, name), map_location='cpu', weights_only=False)


# ---- "dyadic" data: every TF32 product is exact --------------------------------
# ReLU / LeakyReLU make the gradients of SPADE discontinuous: an activation whose
# pre-activation lies within TF32 rounding of zero takes the other branch and its
# gradient changes by O(1), not O(eps).  To check the backward strictly, tests use
# segmaps and modulation weights made of small dyadic rationals (few mantissa
# bits), for which TF32 operands and fp32 accumulation are EXACT, so the forward
# pre-activations - and therefore every gate - agree with the fp32 reference.

def dy_tensor(name, shape, seed, step, levels, density=1.0):
    """Values k*step with integer |k| <= levels, a `density` fraction non-zero."""
    r = _rng(name, seed)
    k = r.randint(-levels, levels + 1, size=tuple(shape)).astype(np.float32)
    if density < 1.0:
        k = k * (r.random_sample(tuple(shape)) < density)
    return torch.from_numpy((k * step).astype(np.float32))


def dyadic_spade_state(state, seed=0, prefix_filter=None):
    """Overwrite the SPADE modulation parameters inside a state_dict with dyadic values."""
    out = dict(state)
    for key, ref in state.items():
        name = key.replace('.module.', '.')
        shape = tuple(ref.shape)
        if name.endswith('mlp_shared.0.weight'):
            out[key] = dy_tensor(name, shape, seed, 0.125, 1, density=0.5)
        elif name.endswith('mlp_shared.0.bias'):
            out[key] = dy_tensor(name, shape, seed, 0.25, 1)
        elif name.endswith('mlp_gamma.weight') or name.endswith('mlp_beta.weight'):
            out[key] = dy_tensor(name, shape, seed, 0.0625, 1, density=0.5)
        elif name.endswith('mlp_gamma.bias') or name.endswith('mlp_beta.bias'):
            out[key] = dy_tensor(name, shape, seed, 0.125, 4)
    return out


def dyadic_seg(name, shape, seed, density):
    return dy_tensor(name, shape, seed, 1.0, 1, density=density)


def c4_masks(name, M, O=10):
    """Soft object masks for the BASELINE config-4 K2b cases, regenerated from the name (not stored)."""
    keep = det_tensor('k2bc4.mask.keep.%s' % name, (O, M, M), 8) > -0.5
    return keep.float() * det_tensor('k2bc4.mask.val.%s' % name, (O, M, M), 8).abs().clamp(max=1.0)

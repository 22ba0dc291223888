"""Options namespace, CATER vocabulary and synthetic CATER-shaped clips.

The reference passes one argparse namespace (``opt``) into every module
constructor (data/args.py:22-207) and attaches the dataset vocabulary to it at
run time (scripts/train.py:337).  ``make_opt`` builds an equivalent namespace
with the reference's defaults for the fields the hot path reads; any object with
the same attributes (e.g. the reference's own parsed args) works as well.
"""
from types import SimpleNamespace

import torch


def cater_vocab():
    """The CATER vocabulary (data/cater.py:93-122): 8 spatial predicates, 7
    actions, 4 attribute tables; index 0 of every table is the __image__ dummy."""
    v = {}
    v['pred_name_to_idx'] = {'__in_image__': 0, 'right': 1, 'above': 2, 'below': 3, 'left': 4,
                             'surrounding': 5, 'inside': 6, '__padding__': 7}
    v['pred_idx_to_name'] = {i: n for n, i in v['pred_name_to_idx'].items()}
    v['action_name_to_idx'] = {'__in_image__': 0, '_no_op': 1, '_slide': 2, '_contain': 3,
                               '_rotate': 4, '_pick_place': 5, '__padding__': 6}
    v['action_idx_to_name'] = {i: n for n, i in v['action_name_to_idx'].items()}
    v['attributes'] = {
        'shape': {'__image__': 0, 'cube': 1, 'sphere': 2, 'cylinder': 3, 'spl': 4, 'cone': 5},
        'color': {'__image__': 0, 'gray': 1, 'red': 2, 'blue': 3, 'green': 4, 'brown': 5,
                  'purple': 6, 'cyan': 7, 'yellow': 8, 'gold': 9},
        'material': {'__image__': 0, 'rubber': 1, 'metal': 2},
        'size': {'__image__': 0, 'small': 1, 'large': 2, 'medium': 3},
    }
    v['reverse_attributes'] = {a: {i: n for n, i in t.items()} for a, t in v['attributes'].items()}
    names, ind = {}, 0
    for table in v['attributes'].values():
        for label in table:
            names[label if ind == 0 else '%s_%d' % (label, ind)] = ind
            ind += 1
    v['object_name_to_idx'] = names
    return v


def make_opt(image_size=256, batch_size=2, **overrides):
    """Reference defaults (data/args.py:27-181) for everything the path reads."""
    size = (image_size, image_size) if isinstance(image_size, int) else tuple(image_size)
    opt = SimpleNamespace(
        image_size=size, batch_size=batch_size, vocab=cater_vocab(),
        embedding_dim=128, gconv_dim=128, gconv_hidden_dim=512, gconv_pooling='avg',
        gconv_num_layers=3, mlp_normalization='none', mask_size=0, only_temporal=0,
        num_upsampling_layers='normal', ngf=64, aspect_ratio=1.0,
        norm_G='spectralspadesyncbatch3x3', norm_F='spectralsyncbatch',
        n_blocks_F=6, nff=32, n_downsample_F=3, flow_deconv=False, flow_multiplier=20,
        frames_per_action=4, n_frames_G=2, bp_prev=0, learning_rate=1e-4, beta1=0.5,
        crop_size=32, gpu_ids=[],
        # discriminator / losses (args.py:61-63,105,152-169)
        num_D=2, n_layers_D=4, ndf=64, norm_D='spectralinstance', use_actions_loss=1, gan_mode='hinge',
        no_ganFeat_loss=False, no_vgg_loss=True, lambda_feat=10.0, lambda_F_warp=10.0,
        discriminator_img_loss_weight=1.0, bbox_pred_loss_weight=10, frames_per_action_graph=4,
    )
    for k, v in overrides.items():
        setattr(opt, k, v)
    opt.semantic_nc = len(opt.vocab['attributes']) * opt.embedding_dim   # args.py:207
    return opt


_W = (0.094, 0.156, 0.219)     # CATER object widths / 320   (data/cater.py:260-325)
_H = (0.125, 0.208, 0.292)     # CATER object heights / 240


def synthetic_batch(B=2, F=4, image_size=256, seed=1234, n_objects=None, n_actions=None,
                    device='cpu', with_images=True, pad_to=None):
    """A seeded CATER-shaped batch with the tensor contract of the reference's
    collate_fn (data/dataset_params.py:8-104): ``imgs [B,F,3,H,W]``, ``objs
    [B,Omax,4]`` (last real row = __image__ dummy = zeros, padding zeros), ``boxes
    [B,F,Omax,4]`` xywh (dummy [0,0,1,1], padding -1), ``triplets [B,F,Tmax,3]``
    ([i, __in_image__, O]; padding [0,7,0]), ``actions [B,Amax,7]``
    ([s, a, o, t1, t2, x_end, y_end]; padding [0,6,0,0,0,0,0])."""
    g = torch.Generator().manual_seed(seed)
    ri = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=g))
    ru = lambda lo, hi, n=1: torch.rand(n, generator=g) * (hi - lo) + lo
    n_obj = [n_objects if n_objects is not None else ri(5, 10) for _ in range(B)]
    n_act = [n_actions if n_actions is not None else ri(2, 6) for _ in range(B)]
    Omax, Tmax, Amax = max(n_obj) + 1, max(n_obj), max(n_act)
    if pad_to is not None:          # fixed shapes (CUDA-graph replay): (objects incl. dummy, actions)
        Omax, Tmax, Amax = max(Omax, pad_to[0]), max(Tmax, pad_to[0] - 1), max(Amax, pad_to[1])
    objs = torch.zeros(B, Omax, 4, dtype=torch.long)
    boxes = torch.full((B, F, Omax, 4), -1.0)
    triplets = torch.zeros(B, F, Tmax, 3, dtype=torch.long)
    triplets[..., 1] = 7
    actions = torch.zeros(B, Amax, 7)
    actions[..., 1] = 6
    hi = (5, 9, 2, 3)
    for b in range(B):
        n = n_obj[b]
        for k in range(4):
            objs[b, :n, k] = torch.randint(1, hi[k] + 1, (n,), generator=g)
        w = torch.tensor([_W[ri(0, 2)] for _ in range(n)])
        h = torch.tensor([_H[ri(0, 2)] for _ in range(n)])
        x0 = ru(0.0, 1.0, n) * (1 - w)
        y0 = ru(0.0, 1.0, n) * (1 - h)
        dx, dy = ru(-0.02, 0.02, n), ru(-0.02, 0.02, n)
        for f in range(F):
            boxes[b, f, :n, 0] = (x0 + f * dx).clamp(0, 1)
            boxes[b, f, :n, 1] = (y0 + f * dy).clamp(0, 1)
            boxes[b, f, :n, 2] = w
            boxes[b, f, :n, 3] = h
            boxes[b, f, n] = torch.tensor([0.0, 0.0, 1.0, 1.0])
            for i in range(n):
                triplets[b, f, i] = torch.tensor([i, 0, n])
        for j in range(n_act[b]):
            s, a = ri(0, n - 1), ri(2, 5)
            xe, ye = (float(ru(0, 0.8)), float(ru(0, 0.8))) if a in (2, 5) else (0.0, 0.0)
            actions[b, j] = torch.tensor([s, a, s, float(ru(-0.5, 0.3)), float(ru(0.7, 1.5)), xe, ye])
    H = image_size
    imgs = torch.randn(B, F, 3, H, H, generator=g) if with_images else None
    batch = dict(imgs=imgs, objs=objs, boxes=boxes, triplets=triplets, actions=actions)
    if device != 'cpu':
        batch = {k: (v.to(device) if v is not None else None) for k, v in batch.items()}
    return batch


def microbench_graph(B=2, O=10, E_spatial=10, E_action=30, seed=7):
    """BASELINE config 4 (graph half): O real objects + dummy, 40 edges, all live."""
    g = torch.Generator().manual_seed(seed)
    E = E_spatial + E_action
    edges = torch.zeros(B, E, 2, dtype=torch.long)
    for b in range(B):
        edges[b, :E_spatial, 0] = torch.arange(E_spatial) % O
        edges[b, :E_spatial, 1] = O
        s = torch.randint(0, O, (E_action,), generator=g)
        edges[b, E_spatial:, 0] = s
        edges[b, E_spatial:, 1] = s
    ind = torch.ones(B, E, dtype=torch.bool)
    return edges, ind

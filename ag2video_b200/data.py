"""CATER clips for the training loop (reference: data/cater.py:180-330, 356-444 and the
collate function data/dataset_params.py:8-104).

On-disk layout (the reference's): ``<data_root>/videos/<video_id>/%05d.png`` - the 301-frame cache the reference
writes on first use (cater.py:421-444; decoding the .avi needs scikit-video, which this image does not have, so a
missing cache is an error here) -, ``<data_root>/scenes/<video_id>.json`` with ``objects`` (``instance``, ``shape``,
``color``, ``material``, ``size``, ``locations`` = {frame: xyz}) and ``movements`` ({instance: [[action, other,
first_frame, last_frame], ...]}), and a label file with one ``<video_id>.avi <labels>`` line per clip.

``CATERDataset[i]`` returns the reference's tuple ``(vids [F,3,H,W], objs {attribute: LongTensor[O+1]},
boxes [F,O+1,4] xywh in [0,1], triplets [F,O,3], actions [A,7], video_id)``; ``collate_fn`` pads to the batch maxima
exactly like the reference (objects with zeros, boxes with -1, triplets with ``[0, __padding__, 0]``, actions with
``[0, __padding__, 0, 0, 0, 0, 0]``) and ``as_batch`` names the tensors the way ``Trainer.iteration`` takes them -
the same contract as ``config.synthetic_batch``.
"""
import json
import os
from glob import glob

import numpy as np
import torch
from torch.utils.data import Dataset

from .config import cater_vocab

# half extents in pixels of the 320x240 render: (half width, extent above the centre, extent below), cater.py:260-325
_FULL = {'large': (35, 35, 35), 'medium': (25, 25, 25), 'small': (15, 15, 15)}
_EXTENT = {
    'spl': _FULL, 'cylinder': _FULL, 'cube': _FULL,
    'cone': {'large': (35, 25, 40), 'medium': (25, 15, 30), 'small': (20, 20, 20)},
    'sphere': {'large': (35, 25, 40), 'medium': (25, 25, 25), 'small': (15, 15, 15)},
}
_RENDER_W, _RENDER_H = 320, 240
# fixed camera of CATER's renderer (cater.py:339-343): rows give x, y, (unused), w of the projected point
_CAMERA = np.array([(1.4503, 1.6376, 0.0000, -0.0251), (-1.0346, 0.9163, 2.5685, 0.0095),
                    (-0.6606, 0.5850, -0.4748, 10.5666), (-0.6592, 0.5839, -0.4738, 10.7452)])
_SKIP = {'CATER_new_004798', 'CATER_new_006532', 'CATER_new_001175', 'CATER_new_000434', 'CATER_new_000346'}   # cater.py:83-85
IMG_MEAN, IMG_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def project_points(xyz):
    """[N,3] world points -> [N,2] image coordinates in [-1, 1], y pointing down (cater.py:331-353)."""
    homo = np.hstack((xyz, np.ones((xyz.shape[0], 1))))
    p = homo @ _CAMERA.T
    return np.stack([p[:, 0] / p[:, 3], -p[:, 1] / p[:, 3]], axis=1)


def scene_boxes(scene):
    """[frames, O+1, 4] xywh boxes in [0,1] for every frame of a scene, last row = the whole image (cater.py:247-329)."""
    per_object = []
    n_frames = None
    for obj in scene['objects']:
        centre = project_points(np.array(list(obj['locations'].values())))
        cx = (centre[:, 0] + 1) * _RENDER_W / 2
        cy = (centre[:, 1] + 1) * _RENDER_H / 2
        half_w, up, down = _EXTENT[obj['shape']][obj['size']]
        x0, y0, x1, y1 = cx - half_w, cy - up, cx + half_w, cy + down
        per_object.append(np.stack([x0 / _RENDER_W, y0 / _RENDER_H, (x1 - x0) / _RENDER_W, (y1 - y0) / _RENDER_H], axis=1))
        n_frames = cx.size
    per_object.append(np.tile([[0., 0., 1., 1.]], [n_frames, 1]))
    return torch.FloatTensor(np.transpose(np.array(per_object), (1, 0, 2)))


class CATERDataset(Dataset):
    def __init__(self, image_dir, data_root, is_test=False, is_val=False, debug=False, nframes=301, frames_mapping=None,
                 image_size=(64, 64), fps=24, frames_per_action=16, initial_frames_per_sample=48, max_samples=None,
                 include_relationships=True, resize_or_crop='resize', fine_size=64, load_size=64, aspect_ratio=1,
                 no_flip=True):
        super().__init__()
        self.data_dir, self.data_root = image_dir, data_root
        self.videos_path = os.path.join(data_root, 'videos')
        self.scenes_path = os.path.join(data_root, 'scenes')
        self.fps, self.nframes = fps, nframes
        self.initial_frames_per_sample, self.frames_per_action = initial_frames_per_sample, frames_per_action
        self.is_val, self.is_test, self.max_samples = is_val, is_test, max_samples
        self.include_relationships = include_relationships
        self.image_size = tuple(image_size)
        self.img_mean, self.img_std = list(IMG_MEAN), list(IMG_STD)
        self.vocab = cater_vocab()
        present = {name.split('.')[0] for name in os.listdir(self.videos_path)}
        self.vid_labels, self.vid_names = {}, []
        with open(image_dir) as f:
            for line in f:
                fields = line.replace('\n', '').split(' ')
                name = fields[0].split('.')[0]
                if name in present and name not in _SKIP:
                    self.vid_labels[name] = [int(n) for n in fields[1].split(',')]
                    self.vid_names.append(name)
        self.json_data = {}
        for fn in os.listdir(self.scenes_path):
            name = fn.split('.')[0]
            if name in self.vid_labels:
                with open(os.path.join(self.scenes_path, fn)) as f:
                    self.json_data[name] = json.load(f)

    def __len__(self):
        return len(self.vid_labels) if self.max_samples is None else min(len(self.vid_labels), self.max_samples)

    # ---- graph side ------------------------------------------------------------------------------------------
    def extract_objs(self, sg):
        """{attribute: LongTensor[O+1]}: one id per object and the __image__ dummy last (cater.py:163-169)."""
        return {attr: torch.LongTensor([table[obj[attr]] for obj in sg['objects']] + [table['__image__']])
                for attr, table in self.vocab['attributes'].items()}

    def extract_triplets(self, boxes):
        """[F, O, 3]: every object is __in_image__ of the dummy (cater.py:171-185)."""
        F, O = boxes.size(0), boxes.size(1) - 1
        in_image = self.vocab['pred_name_to_idx']['__in_image__']
        return torch.LongTensor([[[i, in_image, O] for i in range(O)] for _ in range(F)])

    def _actions(self, sg):
        index = {obj['instance']: i for i, obj in enumerate(sg['objects'])}
        out = []
        for subject, moves in sg['movements'].items():
            for action, other, first, last in moves:
                if last - first < 12:                       # cater.py:201-203
                    continue
                out.append([index[subject], self.vocab['action_name_to_idx'][action],
                            index[other] if other is not None else index[subject], first, last])
        return out

    def extract_actions(self, sg):
        return torch.LongTensor(self._actions(sg))

    def extract_actions_split(self, sg, max_frame, is_test):
        """Actions overlapping a window of ``initial_frames_per_sample`` frames: the first action's start at test time,
        a uniformly random start (numpy's global generator, like the reference) otherwise (cater.py:187-218)."""
        actions = self._actions(sg)
        starts, ends = [a[3] for a in actions], [a[4] for a in actions]
        if is_test:
            start = min(starts)
            end = min(max(ends), start + self.initial_frames_per_sample)
        else:
            start = np.random.randint(0, min(max(ends), max_frame) - self.initial_frames_per_sample + 1)
            end = start + self.initial_frames_per_sample
        return torch.LongTensor([a for a in actions if not (a[3] > end or a[4] < start)]), [start, end]

    extract_bounding_boxes = staticmethod(scene_boxes)

    def normalized_actions(self, actions, boxes, s_frame, e_frame):
        """[A,5] (subject, action, object, first, last) -> [A',7] (subject, action, object, t1, t2, x_end, y_end): the
        window in units of the action's duration, actions outside it dropped, and the subject's final position for
        slide / pick-place (cater.py:446-467)."""
        f1, f2 = actions[:, 3].float(), actions[:, 4].float()
        t1 = (s_frame - f1) / (f2 - f1 + 1)
        t2 = (e_frame - f1) / (f2 - f1 + 1)
        rows = torch.cat([actions[:, :3].float(), torch.stack([t1, t2], dim=-1)], dim=-1)
        keep = ~((t1 > 1) | (t2 < 0))
        rows = rows[keep]
        final = boxes[f2[keep].long(), rows[:, 0].long()][:, :2]
        names = self.vocab['action_name_to_idx']
        moves = (rows[:, 1] == names['_pick_place']) | (rows[:, 1] == names['_slide'])
        final[~moves] = 0.
        return torch.cat([rows, final], dim=1)

    # ---- pixels ------------------------------------------------------------------------------------------------
    def extract_frames(self, video_id):
        cache = os.path.join(self.videos_path, video_id)
        if not os.path.isdir(cache):
            raise RuntimeError('CATER frame cache %s is missing: extract the %d frames of %s.avi to %%05d.png first '
                               '(the reference does it with scikit-video, cater.py:421-444)' % (cache, self.nframes, video_id))
        frames = sorted(glob(os.path.join(cache, '*.png')))
        if len(frames) != self.nframes:
            print('Number of frames in %s is %d' % (video_id, len(frames)))
            return None
        return np.array(frames)

    def load_frames(self, paths):
        """[F,3,H,W]: bilinear resize, [0,1], ImageNet mean / std (cater.py:141-151)."""
        from PIL import Image
        H, W = self.image_size
        out = []
        for fn in paths:
            img = Image.open(fn).convert('RGB').resize((W, H), Image.BILINEAR)
            out.append(torch.from_numpy(np.asarray(img, dtype=np.uint8).copy()).permute(2, 0, 1).float().div(255))
        vids = torch.stack(out, 0)
        mean = torch.tensor(self.img_mean).view(1, 3, 1, 1)
        std = torch.tensor(self.img_std).view(1, 3, 1, 1)
        return (vids - mean) / std

    def __getitem__(self, index):
        video_id = self.vid_names[index]
        sg = self.json_data[video_id]
        paths = self.extract_frames(video_id)
        if paths is None:
            return None, None, None, None, None, None
        actions, (s_frame, e_frame) = self.extract_actions_split(sg, len(paths) - 1, self.is_test)
        frames = list(range(s_frame, e_frame))
        assert len(frames) == self.initial_frames_per_sample, 'different size'
        frames = frames[0:self.initial_frames_per_sample:self.initial_frames_per_sample // self.frames_per_action]
        vids = self.load_frames(paths[frames])
        all_boxes = scene_boxes(sg)
        boxes = all_boxes[frames]
        return (vids, self.extract_objs(sg), boxes, self.extract_triplets(boxes),
                self.normalized_actions(actions, all_boxes, s_frame, e_frame), '%s_%d-%d' % (video_id, s_frame, e_frame))


def collate_fn(vocab, batch):
    """(imgs [B,F,3,H,W], objs [B,Omax,n_attr], boxes [B,F,Omax,4], triplets [B,F,Tmax,3], actions [B,Amax,7], ids)
    padded to the batch maxima (data/dataset_params.py:8-104); clips that failed to load are dropped."""
    batch = [item for item in batch if item[0] is not None and not isinstance(item[0], bool)]
    if not batch:
        return None, None, None, None, None, None
    Omax = max(next(iter(objs.values())).size(0) for _, objs, _, _, _, _ in batch)
    Tmax = max(triplets.size(1) for _, _, _, triplets, _, _ in batch)
    Amax = max(actions.size(0) for _, _, _, _, actions, _ in batch)
    pad_pred, pad_act = vocab['pred_name_to_idx']['__padding__'], vocab['action_name_to_idx']['__padding__']
    vids, all_objs, all_boxes, all_triplets, all_actions, ids = [], [], [], [], [], []
    for vid, objs, boxes, triplets, actions, video_id in batch:
        O, F = boxes.size(1), boxes.size(0)
        table = torch.zeros(Omax, len(objs), dtype=torch.long)              # attribute columns in the vocabulary's order
        for k, values in enumerate(objs.values()):
            table[:O, k] = values
        boxes = torch.cat([boxes, boxes.new_full((F, Omax - O, 4), -1.0)], dim=1)
        fill = torch.tensor([0, pad_pred, 0], dtype=torch.long).repeat(F, Tmax - triplets.size(1), 1)
        triplets = torch.cat([triplets, fill], dim=1)
        fill = torch.tensor([0, pad_act, 0, 0, 0, 0, 0], dtype=torch.float32).repeat(Amax - actions.size(0), 1)
        actions = torch.cat([actions, fill], dim=0)
        vids.append(vid); all_objs.append(table); all_boxes.append(boxes); all_triplets.append(triplets)
        all_actions.append(actions); ids.append(video_id)
    return (torch.stack(vids), torch.stack(all_objs), torch.stack(all_boxes), torch.stack(all_triplets),
            torch.stack(all_actions), ids)


def as_batch(collated, device=None):
    """The collated tuple as the dict ``Trainer.iteration`` takes (the contract of ``config.synthetic_batch``)."""
    imgs, objs, boxes, triplets, actions, _ = collated
    out = dict(imgs=imgs, objs=objs, boxes=boxes, triplets=triplets, actions=actions)
    if device is not None:
        out = {k: (v.to(device, non_blocking=True) if v is not None else None) for k, v in out.items()}
    return out


def build_loader(image_dir, data_root, batch_size, image_size=(256, 256), frames_per_action=4, initial_frames_per_sample=48,
                 is_test=False, num_workers=0, rank=0, world=1, seed=0, **kw):
    """DataLoader over CATER clips; with ``world`` > 1 every rank draws a disjoint, equally sized shard
    (DistributedSampler, drop_last) - the sharding ``dist.shard_clips`` describes."""
    from functools import partial
    from torch.utils.data import DataLoader
    from torch.utils.data.distributed import DistributedSampler
    ds = CATERDataset(image_dir, data_root, is_test=is_test, image_size=image_size, frames_per_action=frames_per_action,
                      initial_frames_per_sample=initial_frames_per_sample, **kw)
    sampler = DistributedSampler(ds, num_replicas=world, rank=rank, shuffle=not is_test, seed=seed, drop_last=True) if world > 1 else None
    return DataLoader(ds, batch_size=batch_size, shuffle=(sampler is None and not is_test), sampler=sampler,
                      num_workers=num_workers, collate_fn=partial(collate_fn, ds.vocab), drop_last=True, pin_memory=True)

"""A tiny deterministic dataset in CATER's on-disk layout (see ag2video_b200/data.py), written by the golden
generator (tests/golden/make_golden.py cater) and by tests/test_cater_loader.py - never stored."""
import json
import os

import numpy as np


def write(root, n_frames=301, size=(20, 15)):
    from PIL import Image
    rng = np.random.RandomState(1234)
    os.makedirs(os.path.join(root, 'videos'), exist_ok=True)
    os.makedirs(os.path.join(root, 'scenes'), exist_ok=True)
    shapes, colors = ['cube', 'sphere', 'cylinder', 'spl', 'cone'], ['gray', 'red', 'blue', 'green', 'brown', 'purple', 'cyan', 'yellow', 'gold']
    materials, sizes = ['rubber', 'metal'], ['small', 'large', 'medium']
    names = ['CATER_new_000001', 'CATER_new_000002', 'CATER_new_000346']          # the last one is on the reference's skip list
    labels = []
    for v, name in enumerate(names):
        n_obj = 3 + v
        objects = []
        for i in range(n_obj):
            start = rng.uniform(-2.5, 2.5, size=3) * np.array([1, 1, 0]) + np.array([0, 0, 0.35])
            drift = rng.uniform(-0.004, 0.004, size=3) * np.array([1, 1, 0])
            objects.append({'instance': 'obj_%d' % i, 'shape': shapes[(i + v) % 5], 'color': colors[(2 * i + v) % 9],
                            'material': materials[i % 2], 'size': sizes[(i + 2 * v) % 3],
                            'locations': {str(f): (start + f * drift).tolist() for f in range(n_frames)}})
        movements = {'obj_0': [['_slide', None, 5, 40], ['_rotate', None, 50, 58], ['_pick_place', None, 70, 110]],
                     'obj_1': [['_contain', 'obj_2', 20, 60], ['_no_op', None, 100, 160]],
                     'obj_2': [['_rotate', None, 0, 30], ['_slide', None, 200, 260]]}
        with open(os.path.join(root, 'scenes', name + '.json'), 'w') as f:
            json.dump({'objects': objects, 'movements': movements}, f)
        vdir = os.path.join(root, 'videos', name)
        os.makedirs(vdir, exist_ok=True)
        base = rng.randint(0, 256, size=(size[1], size[0], 3)).astype(np.int32)
        for f in range(n_frames):
            frame = ((base + 3 * f + rng.randint(0, 8, size=base.shape)) % 256).astype(np.uint8)
            Image.fromarray(frame).save(os.path.join(vdir, '%05d.png' % f))
        labels.append('%s.avi %d,%d' % (name, v, v + 1))
    labels.append('CATER_new_009999.avi 1')                                        # labelled, but no video on disk
    with open(os.path.join(root, 'labels.txt'), 'w') as f:
        f.write('\n'.join(labels) + '\n')
    return os.path.join(root, 'labels.txt'), root

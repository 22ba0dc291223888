"""Where one stage of the recurrence kernel spends its time: globaltimer stamps of CTA 0 (k1r_recur.cu: stamp()).
    python tools/recur_timeline.py [T] [B]   (on the GPU box)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ag2video_b200 import _lib as L  # noqa: E402
from ag2video_b200 import recurrence  # noqa: E402
from ag2video_b200.config import make_opt, synthetic_batch  # noqa: E402
from ag2video_b200.networks import Acts2LayoutModel  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 16
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
if len(sys.argv) > 3:
    L.lib().ag2v_recur_set_core(int(sys.argv[3]))
m = Acts2LayoutModel(make_opt(64)).cuda()
b = {k: v.cuda() for k, v in synthetic_batch(B=B, F=T, image_size=8, seed=1, with_images=False, pad_to=(11, 6)).items() if v is not None}
for _ in range(3):
    m(b['objs'], b['triplets'], b['actions'], b['boxes'])
buf = torch.zeros(1 + 2 * 4000, dtype=torch.int64, device='cuda')
L.lib().ag2v_recur_set_profile(L.ptr(buf))
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
m(b['objs'], b['triplets'], b['actions'], b['boxes'])
e.record()
torch.cuda.synchronize()
L.lib().ag2v_recur_set_profile(None)
ev = buf.cpu().tolist()
n = ev[0]
rows = [(ev[1 + 2 * i], ev[2 + 2 * i]) for i in range(n)]
print('launch %.1f us for T=%d, B=%d; %d stamps' % (s.elapsed_time(e) * 1e3, T, B, n))
names = {1: 'A load issued', 2: 'A in smem', 3: 'MMA loop done', 4: 'reduce+epilogue done', 5: 'cluster handshake done',
         6: 'chunk wait', 7: 'chunk ready', 8: 'chunk consumed (warp 0)'}
acc = {}
for (w0, t0), (w1, t1) in zip(rows[:-1], rows[1:]):
    key = '%s -> %s' % (names[w0], names[w1])
    a = acc.setdefault(key, [0, 0])
    a[0] += t1 - t0
    a[1] += 1
tot = rows[-1][1] - rows[0][1]
for k, (ns, c) in sorted(acc.items(), key=lambda kv: -kv[1][0]):
    print('%-50s %8.1f us total  %6.2f us avg x %d  (%.0f%%)' % (k, ns / 1e3, ns / 1e3 / c, c, 100.0 * ns / tot))
print('first 120 stamps (us since the first):')
for w, t in rows[:120]:
    print('  %8.2f  %s' % ((t - rows[0][1]) / 1e3, names[w]))

"""Worker of tests/test_gpu_syncbn.py: one data-parallel rank (of 2) sharing cuda:0, gloo rendezvous.
Runs SPADE and the affine bn_act stage on ITS sample of a 2-sample batch with SyncBN statistics
(ag2video_b200.spade.set_sync_bn) and saves outputs, running statistics and every gradient."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from _util import det_tensor, load_det  # noqa: E402


def case_tensors(C, L, r, Hs):
    return (det_tensor('sync.x', (2, C, r, r), 4).mul(1.4).add(0.3), det_tensor('sync.seg', (2, L, Hs, Hs), 4),
            det_tensor('sync.cot', (2, C, r, r), 4))


def run(rank_slice, C=16, L=8, r=8, Hs=16):
    """rank_slice: slice of the batch this process owns (slice(None) = the single-process run)."""
    import ag2video_b200.spade as sp
    from ag2video_b200.networks import _BN2d
    torch.backends.cudnn.allow_tf32 = False
    sp.CONV_IMPL = 3                               # fp32-class products: the comparison isolates the statistics
    x0, seg0, cot = case_tensors(C, L, r, Hs)
    res = {}
    m = sp.SPADE('spadesyncbatch3x3', C, L)
    load_det(m, 33)
    m.fused_slope = 0.2
    m = m.cuda().train()
    x, seg = x0[rank_slice].cuda().requires_grad_(), seg0[rank_slice].cuda().requires_grad_()
    out = m(x, seg)
    (out * cot[rank_slice].cuda()).sum().backward()
    res['spade'] = dict(out=out.detach().cpu(), dx=x.grad.cpu(), dseg=seg.grad.cpu(),
                        grads={k: p.grad.cpu() for k, p in m.named_parameters()},
                        buffers={k: v.cpu() for k, v in m.named_buffers()})
    bn = _BN2d(C)
    load_det(bn, 34)
    bn = bn.cuda().train()
    x = x0[rank_slice].cuda().requires_grad_()
    y = bn(x, 1, 0.2)
    (y * cot[rank_slice].cuda()).sum().backward()
    res['bn_act'] = dict(out=y.detach().cpu(), dx=x.grad.cpu(), grads={k: p.grad.cpu() for k, p in bn.named_parameters()},
                         buffers={k: v.cpu() for k, v in bn.named_buffers()})
    return res


if __name__ == '__main__':
    rank, world, port, outdir = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%s' % port, rank=rank, world_size=world)
    import ag2video_b200.spade as sp
    sp.set_sync_bn(True)
    torch.save(run(slice(rank, rank + 1)), os.path.join(outdir, 'rank%d.pt' % rank))
    from ag2video_b200 import peer
    if os.environ.get('AG2V_EXPECT_PEER') == '1':         # tests/test_gpu_peer.py: the sums must have travelled by peer memory
        assert peer.status()['collective'] == 'peer', peer.status()
    else:
        assert peer.status()['collective'] == 'group', peer.status()
    dist.barrier()
    dist.destroy_process_group()

"""Worker of tests/test_gpu_peer.py: one rank of a `world`-rank node, all ranks on cuda:0 (time-sliced), gloo carries
the CUDA IPC handles.  Runs a fixed list of peer all-reduces on both channels and saves the results."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SIZES = [1, 2, 7, 511, 512, 513, 2048, 5 * 1024 + 3, 16384, 6, 6, 6, 6, 6, 6]      # the tail cycles through all slots


def contribution(rank, call, n):
    """What `rank` contributes to call number `call`: exact in double, different on every rank and call."""
    i = torch.arange(n, dtype=torch.float64)
    return (i % 97 + 1) * (rank + 1) + call * 1e-3 * (rank + 2) + (i % 5 == rank).double() * 1e9


def main():
    rank, world, port, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.cuda.set_device(0)
    from ag2video_b200 import peer
    ex = peer.PeerExchange(None)
    res = []
    for call, n in enumerate(SIZES):
        v = contribution(rank, call, n).cuda()
        if call % 3 == 2:                          # every third call overlapped on the side stream (channel 1)
            pending = ex.allreduce_async(v)
            filler = torch.ones(1 << 20, device='cuda').sum()      # unrelated work on the current stream
            pending.wait()
            del filler
        else:
            ex.allreduce(v)
        res.append(v.cpu())
    torch.cuda.synchronize()
    torch.save(res, os.path.join(out, 'peer%d.pt' % rank))
    dist.barrier()
    ex.close()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()

"""Worker of tests/test_gpu_peer.py: one rank of a `world`-rank node, all ranks on cuda:0 (time-sliced), gloo carries
the CUDA IPC handles.  Runs a fixed list of peer all-reduces on both channels and saves the results."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SIZES = [1, 2, 7, 511, 512, 513, 2048, 5 * 1024 + 3, 16384, 6, 6, 6, 6, 6, 6]      # the tail cycles through all slots
GRAPH_SIZES = [2048, 640, 10240]
GRAPH_REPLAYS = 6


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
    # the same launches captured in a CUDA graph and replayed (how bench.py runs the training step): the call number is
    # counted on the device, so every replay is a new exchange
    bufs = [torch.zeros(n, dtype=torch.float64, device='cuda') for n in GRAPH_SIZES]
    stream = torch.cuda.Stream()
    stream.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(stream):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            ex.allreduce(bufs[0])
            pending = ex.allreduce_async(bufs[1])
            ex.allreduce(bufs[2])
            pending.wait()
        replays = []
        for rep in range(GRAPH_REPLAYS):
            for k, n in enumerate(GRAPH_SIZES):
                bufs[k].copy_(contribution(rank, 100 + 10 * rep + k, n))
            g.replay()
            replays.append([b.cpu() for b in bufs])
    torch.cuda.synchronize()
    torch.save(dict(eager=res, graph=replays), os.path.join(out, 'peer%d.pt' % rank))
    dist.barrier()
    ex.close()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()

"""Worker of tests/test_gpu_peer.py::test_copy_engine_gradient_exchange...: one rank of `world` ranks sharing cuda:0
(gloo carries the IPC handles).  GradBuckets with AG2V_GRAD_ALLREDUCE=ce-force over a handful of parameters: hooked
launches during a real backward, three eager steps, then the same step captured in a CUDA graph and replayed."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SHAPES = [(300, 1000), (64, 32, 3, 3), (7,), (513, 511), (128, 64, 3, 3), (1000, 300), (5, 5)]


def coefficient(rank, step, k, shape):
    g = torch.Generator().manual_seed(1000 * step + 10 * k + rank)
    return torch.randn(shape, generator=g)


def main():
    rank, world, port, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=port, AG2V_GRAD_ALLREDUCE='ce-force', AG2V_BUCKET_MB='1')
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.cuda.set_device(0)
    torch.cuda.set_stream(torch.cuda.Stream())
    from ag2video_b200.dist import GradBuckets
    params = [torch.nn.Parameter(torch.zeros(s, device='cuda')) for s in SHAPES]
    params[4].data = params[4].data.contiguous(memory_format=torch.channels_last)
    buckets = GradBuckets(params)
    assert buckets.ce is not None and len(buckets.buckets) >= 3, (buckets.ce, len(buckets.buckets))
    coefs = [torch.zeros(s, device='cuda') for s in SHAPES]          # static inputs of the step

    def step():
        for p in params:
            p.grad = None
        buckets.begin()
        loss = sum((p * c).sum() for p, c in zip(params[:-1], coefs[:-1]))      # the last parameter gets no gradient
        loss.backward()
        buckets.allreduce()

    def load(s):
        for k, c in enumerate(coefs):
            c.copy_(coefficient(rank, s, k, SHAPES[k]))

    results = []
    for s in range(3):
        load(s)
        step()
        torch.cuda.synchronize()
        results.append([p.grad.detach().cpu().clone() for p in params])
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=torch.cuda.current_stream()):
        step()
    for s in range(3, 6):
        load(s)
        g.replay()
        torch.cuda.synchronize()
        results.append([p.grad.detach().cpu().clone() for p in params])
    torch.save(dict(results=results, hook=buckets.launched_by_hook, end=buckets.launched_at_end), os.path.join(out, 'gradex%d.pt' % rank))
    dist.barrier()
    dist.destroy_process_group()
    os._exit(0)


if __name__ == '__main__':
    main()

"""All-reduce bandwidth of the gradient buckets alone (no compute next to them): N ranks, float32, AVG.
torchrun --nproc-per-node N tools/nccl_bw.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    for mb in (1, 8, 53, 477):
        t = torch.randn(mb * 2**20 // 4, device='cuda')
        for _ in range(5):
            dist.all_reduce(t, op=dist.ReduceOp.AVG)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            dist.all_reduce(t, op=dist.ReduceOp.AVG)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        if rank == 0:
            print('all_reduce %4d MB x %d ranks: %.3f ms  algbw %.0f GB/s  busbw %.0f GB/s'
                  % (mb, world, ms, mb * 2**20 / ms / 1e6, mb * 2**20 / ms / 1e6 * 2 * (world - 1) / world), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()

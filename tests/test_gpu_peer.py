"""K8 - the SyncBN statistics exchange over peer memory (csrc/k8_peer.cu) through the C ABI.  Ranks are processes
sharing cuda:0 (CUDA IPC between processes, the GPU time-slices their kernels), so the test runs on a one-GPU box;
with one GPU per rank the same kernel crosses NVLink.  Integer-exact: the totals are sums of doubles in rank order,
compared bit for bit with the same sum on the host, and all ranks hold identical bits."""
import os
import socket
import subprocess
import sys

import pytest
import torch

from _util import max_rel

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _spawn(script, world, tmp_path, env=None):
    port = str(_free_port())
    e = dict(os.environ, **(env or {}))
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, script), str(r), str(world), port, str(tmp_path)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=e) for r in range(world)]
    for p in procs:
        out, _ = p.communicate(timeout=600)
        assert p.returncode == 0, out[-4000:]


@pytest.mark.parametrize('world', [2, 3])
def test_peer_allreduce_bit_exact_and_identical_on_all_ranks(world, tmp_path):
    import _peer_worker as w
    _spawn('_peer_worker.py', world, tmp_path)
    got = [torch.load(os.path.join(str(tmp_path), 'peer%d.pt' % r), weights_only=False) for r in range(world)]
    def total(call, n):
        want = torch.zeros(n, dtype=torch.float64)
        for r in range(world):                       # rank order, as the kernel sums
            want = want + w.contribution(r, call, n)
        return want

    for call, n in enumerate(w.SIZES):
        for r in range(world):
            assert got[r]['eager'][call].shape == (n,)
            assert torch.equal(got[r]['eager'][call], total(call, n)), 'call %d (n=%d) rank %d' % (call, n, r)
    # CUDA-graph replays: every replay is a fresh exchange of that replay's inputs
    for rep in range(w.GRAPH_REPLAYS):
        for k, n in enumerate(w.GRAPH_SIZES):
            for r in range(world):
                assert torch.equal(got[r]['graph'][rep][k], total(100 + 10 * rep + k, n)), 'replay %d vector %d rank %d' % (rep, k, r)


def test_syncbn_over_peer_memory_equals_one_process_tol1e5(tmp_path):
    """The SyncBN case of tests/test_gpu_syncbn.py with the sums carried by the peer exchange instead of the process
    group (AG2V_PEER_SYNCBN=force: the group is gloo here): SPADE and bn_act, forward and backward."""
    _spawn('_syncbn_worker.py', 2, tmp_path, env={'AG2V_PEER_SYNCBN': 'force', 'AG2V_EXPECT_PEER': '1'})
    ranks = [torch.load(os.path.join(str(tmp_path), 'rank%d.pt' % r), weights_only=False) for r in range(2)]
    import _syncbn_worker as w
    import ag2video_b200.spade as sp
    old = sp.CONV_IMPL
    try:
        sp.set_sync_bn(False)
        one = w.run(slice(None))
    finally:
        sp.CONV_IMPL = old
    worst = 0.0
    for op in ('spade', 'bn_act'):
        for r in range(2):
            worst = max(worst, max_rel(ranks[r][op]['out'], one[op]['out'][r:r + 1]), max_rel(ranks[r][op]['dx'], one[op]['dx'][r:r + 1]))
            for k, v in one[op]['buffers'].items():
                if v.is_floating_point():
                    worst = max(worst, max_rel(ranks[r][op]['buffers'][k], v))
        for k, v in one[op]['grads'].items():
            worst = max(worst, max_rel(ranks[0][op]['grads'][k] + ranks[1][op]['grads'][k], v))
    print('SyncBN over peer memory, 2 ranks x 1 sample vs 1 process x 2 samples: %.2e' % worst)
    assert worst <= 1e-5


@pytest.mark.parametrize('world', [2, 3])
def test_copy_engine_gradient_exchange_equals_the_mean_over_ranks(world, tmp_path):
    """GradBuckets with the copy-engine exchange (K8b: cudaMemcpyAsync between IPC windows + one reduction kernel per
    bucket) on a real backward with hooked bucket launches: every rank ends with the mean of the ranks' gradients
    (1e-6: float sums in rank order, then * 1/world), identical bits on all ranks, a parameter without gradient gets
    zeros; three eager steps and three replays of the step captured in a CUDA graph."""
    import _gradex_worker as w
    _spawn('_gradex_worker.py', world, tmp_path)
    got = [torch.load(os.path.join(str(tmp_path), 'gradex%d.pt' % r), weights_only=False) for r in range(world)]
    assert got[0]['hook'] > 0                      # buckets were exchanged from the backward hooks, not only at the end
    for s in range(6):
        for k, shape in enumerate(w.SHAPES):
            want = sum(w.coefficient(r, s, k, shape) for r in range(world)) / world
            if k == len(w.SHAPES) - 1:
                want = torch.zeros(shape)
            for r in range(world):
                g = got[r]['results'][s][k]
                assert torch.equal(g, got[0]['results'][s][k]), 'step %d parameter %d: rank %d differs from rank 0' % (s, k, r)
                assert max_rel(g, want) <= 1e-6 or float(want.abs().max()) == 0 and float(g.abs().max()) == 0, (s, k, r)

"""SyncBN semantics of the data-parallel path (reference:
sync_batchnorm/batchnorm.py:74-83,105-145): two ranks with one sample each must reproduce ONE process
with both samples - outputs, running statistics and all gradients to 1e-5.  The two ranks are two
processes sharing cuda:0 with a gloo rendezvous on 127.0.0.1 (the all-reduce of the 2C / 5C double
sums is the only exchange); the same code path runs over NCCL with one GPU per rank."""
import os
import socket
import subprocess
import sys

import pytest
import torch

from _util import max_rel

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
TOL = 1e-5


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def test_two_ranks_equal_one_process_tol1e5(tmp_path):
    port = str(_free_port())
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, '_syncbn_worker.py'), str(r), '2', port, str(tmp_path)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    for p in procs:
        out, _ = p.communicate(timeout=600)
        assert p.returncode == 0, out[-4000:]
    ranks = [torch.load(os.path.join(str(tmp_path), 'rank%d.pt' % r), weights_only=False) for r in range(2)]
    import _syncbn_worker as w
    import ag2video_b200.spade as sp
    old = sp.CONV_IMPL
    try:
        sp.set_sync_bn(False)
        one = w.run(slice(None))
    finally:
        sp.CONV_IMPL = old
    worst = {}
    for op in ('spade', 'bn_act'):
        ref = one[op]
        for r in range(2):
            got = ranks[r][op]
            worst['%s.out' % op] = max(worst.get('%s.out' % op, 0), max_rel(got['out'], ref['out'][r:r + 1]))
            worst['%s.dx' % op] = max(worst.get('%s.dx' % op, 0), max_rel(got['dx'], ref['dx'][r:r + 1]))
            if 'dseg' in ref:
                worst['%s.dseg' % op] = max(worst.get('%s.dseg' % op, 0), max_rel(got['dseg'], ref['dseg'][r:r + 1]))
            for k, v in ref['buffers'].items():          # every rank ends with the global-batch running statistics
                if v.is_floating_point():
                    worst['%s.%s' % (op, k)] = max(worst.get('%s.%s' % (op, k), 0), max_rel(got['buffers'][k], v))
                else:
                    assert torch.equal(got['buffers'][k], v)
        for k, v in ref['grads'].items():                # data-parallel gradients add up to the single-process gradient
            worst['%s.d%s' % (op, k)] = max_rel(ranks[0][op]['grads'][k] + ranks[1][op]['grads'][k], v)
    print('SyncBN 2 ranks x 1 sample vs 1 process x 2 samples: ' + ', '.join('%s %.1e' % kv for kv in worst.items()))
    assert max(worst.values()) <= TOL, {k: v for k, v in worst.items() if v > TOL}

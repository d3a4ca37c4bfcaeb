"""the data-parallel training step on 2 GPUs (NCCL) against the same step on one GPU.
Needs two CUDA devices (`gpurun --gpus 2`); skipped otherwise."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_dp2_step_matches_single_gpu():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
           '127.0.0.1', '--master-port', str(port), os.path.join(ROOT, 'tests', 'dp_worker.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    if r.returncode != 0 or 'dp ok' not in r.stdout:
        # the interleaved two-rank stderr is long: keep rank 0's traceback and every error / assertion line
        lines = r.stderr.splitlines()
        digest = [ln for ln in lines if '[rank0]' in ln or 'Error' in ln or 'assert' in ln]
        raise AssertionError('data-parallel worker failed:\n' + '\n'.join(digest[-60:]) + '\nstdout: ' + r.stdout[-500:])

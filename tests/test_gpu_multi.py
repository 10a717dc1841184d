"""Multi-GPU parity (block rows over ranks, NCCL exchanges): needs >= 2 GPUs on the box.
Launches tests/mgpu_worker.py under torchrun and reads its verdict line."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("variant,extra", [('deltaU_to_deltaP', []), ('deltaU_to_deltaP', ['--comm', 'nccl']),
                                           ('deltaU_to_deltaP', ['--near-wall', '0.05', '--halo', 'grid']),
                                           ('U_to_gradP', []), ('U_to_gradP', ['--halo', 'grid'])])
def test_two_rank_shards_match_oracle_and_single_gpu(variant, extra):
    if _n_gpus() < 2:
        pytest.skip('needs 2 GPUs (run with gpurun --gpus 2)')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', '29533', os.path.join(HERE, 'mgpu_worker.py'), '--variant', variant] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout[-3000:])
    sys.stderr.write(r.stderr[-3000:])
    assert r.returncode == 0 and 'MGPU_PARITY_OK' in r.stdout

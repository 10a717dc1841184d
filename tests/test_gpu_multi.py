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


def _run(world, variant, extra):
    if _n_gpus() < world:
        pytest.skip('needs %d GPUs (run with gpurun --gpus %d)' % (world, world))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world), '--master-addr', '127.0.0.1',
           '--master-port', '29533', os.path.join(HERE, 'mgpu_worker.py'), '--variant', variant] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout[-3000:])
    sys.stderr.write(r.stderr[-3000:])
    assert r.returncode == 0 and 'MGPU_PARITY_OK' in r.stdout


@pytest.mark.parametrize("variant,extra", [('deltaU_to_deltaP', []), ('deltaU_to_deltaP', ['--comm', 'nccl']),
                                           ('deltaU_to_deltaP', ['--near-wall', '0.05', '--halo', 'grid']),
                                           ('deltaU_to_deltaP', ['--builder', 'band']), ('deltaU_to_deltaP', ['--legacy']),
                                           ('U_to_gradP', []), ('U_to_gradP', ['--halo', 'grid']), ('U_to_gradP', ['--builder', 'band'])])
def test_two_rank_shards_match_oracle_and_single_gpu(variant, extra):
    _run(2, variant, extra)


@pytest.mark.parametrize("world", [4, 8])
@pytest.mark.parametrize("variant,extra", [('deltaU_to_deltaP', ['--builder', 'band']), ('deltaU_to_deltaP', []),
                                           ('U_to_gradP', ['--builder', 'band'])])
def test_four_and_eight_rank_shards(world, variant, extra):
    """World 4 and 8 on a tall mesh (>= 1 block row per rank): interior ranks exchange with BOTH neighbours, the offset
    recurrence crosses every rank boundary, and the band-local table builder (the one bench.py uses) is checked against the
    oracle and a single-GPU handle built from global tables."""
    _run(world, variant, extra)

"""The only collective of the system under pytest: two NCCL ranks (one per GPU) run three PPO updates - the minibatch
step replayed as a CUDA graph that contains the flat-gradient all-reduce and the fused clip + Adam - and must end
every update with bit-identical parameters although their env shards differ (agents/ppo.py:112-169 data-parallel).
Needs two GPUs: run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_ddp.py -m gpu`; skipped on one."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_nccl_update_keeps_replicas_bit_identical():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (NCCL refuses two ranks on one device)")
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "ddp_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, out
        assert "rank %d ok" % rank in out
    assert "graph identical: True" in outs[0]

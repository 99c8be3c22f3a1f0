"""Host-side multi-GPU plumbing on CPU: world_size-2 gloo processes exercise the env sharding rule and
the max-over-ranks / sum-over-ranks reductions bench.py and compute_gae use (no GPU, no collectives on
the rollout path itself - envs are independent)."""
import os
import socket
import subprocess
import sys
import textwrap

import pytest

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_env_shard_partitions_exactly():
    import uavenv_b200  # noqa: F401  (registers the package)
    from target_allocation_ppo_transformer_b200.parallel import env_shard
    for total in (1, 7, 64, 65536, 65537):
        for world in (1, 2, 3, 8):
            spans = [env_shard(total, r, world) for r in range(world)]
            assert spans[0][0] == 0
            for (b0, c0), (b1, _) in zip(spans, spans[1:]):
                assert b0 + c0 == b1
            assert spans[-1][0] + spans[-1][1] == total
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        env_shard(8, 2, 2)


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import torch, torch.distributed as dist
    import uavenv_b200
    from target_allocation_ppo_transformer_b200 import parallel
    rank, local_rank, world = parallel.init("gloo")
    assert world == 2 and dist.get_world_size() == 2
    base, count = parallel.env_shard(65537, rank, world)
    # every global env id is owned exactly once
    owned = parallel.reduce_scalar(count, "sum")
    assert owned == 65537, owned
    assert (base, count) == ((0, 32769) if rank == 0 else (32769, 32768))
    # job time = max over ranks; job throughput = sum of env-steps / that time
    t = parallel.reduce_scalar(1.0 + rank, "max")
    assert t == 2.0
    assert parallel.reduce_scalar(5.0 - rank, "min") == 4.0
    # advantage-normalisation statistics {count, sum, sumsq} are summed over ranks (SURVEY 8e)
    stats = torch.tensor([10.0, 3.0 + rank, 7.0], dtype=torch.float64)
    dist.all_reduce(stats)
    assert stats.tolist() == [20.0, 7.0, 14.0]
    parallel.barrier()
    dist.destroy_process_group()
    print("rank", rank, "ok")
""")


def test_two_rank_gloo_reductions(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, out
        assert "rank %d ok" % rank in out


def test_bench_reference_arm_runs_on_rank0_only():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "3", "--warmup", "3"], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == ""           # other ranks exit 0 without work
    env = dict(os.environ, RANK="0", LOCAL_RANK="0", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "3", "--warmup", "3", "--workload", "c2"], env=env, capture_output=True, text=True,
                         timeout=300)
    import json
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["metric"] == "env_steps_per_sec"

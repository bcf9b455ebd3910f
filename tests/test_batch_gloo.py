"""N > 1 host logic on CPU: two gloo ranks shard a batch of images (image i -> rank i mod N), carve their shards
(through the oracle here -- the GPU ranks do the same through liblqr-1.so) and gather the results on rank 0."""
import importlib
import json
import os
import subprocess
import sys
import textwrap

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
batch = importlib.import_module("gimp-lqr-plugin_b200.batch")


def test_shards_partition_the_batch():
    for n in (1, 5, 8, 256):
        for world in (1, 2, 4, 8):
            shards = [batch.shard_indices(n, world, r) for r in range(world)]
            flat = sorted(i for s in shards for i in s)
            assert flat == list(range(n))
            assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
    with pytest.raises(ValueError):
        batch.shard_indices(4, 2, 2)


WORKER = textwrap.dedent("""
    import importlib, json, os, sys
    sys.path.insert(0, {repo!r})
    import torch.distributed as dist
    pkg = importlib.import_module("gimp-lqr-plugin_b200")
    batch = importlib.import_module("gimp-lqr-plugin_b200.batch")
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    lib = pkg.load_oracle()
    local = batch.carve_shard(lib, batch.shard_indices(6, world, rank), 48, 36, 40, 38)
    merged = batch.gather_results(local, dist)
    if rank == 0:
        print("RESULT " + json.dumps({{str(k): v for k, v in sorted(merged.items())}}))
    dist.barrier()
    dist.destroy_process_group()
""")


def test_two_rank_gloo_batch_equals_serial(tmp_path, oracle):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(repo=REPO))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29571", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = next(l for l in out.stdout.splitlines() if l.startswith("RESULT "))
    got = json.loads(line[len("RESULT "):])
    serial = batch.carve_shard(oracle, range(6), 48, 36, 40, 38)
    assert {str(k): [list(v[0]), v[1]] for k, v in serial.items()} == {k: [list(v[0]), v[1]] for k, v in got.items()}

"""CPU test of the N>1 path: world_size-2 gloo ranks shard a batch with partition_by_bytes, each rank
processes only its slice (with the oracle standing in for the device), and the max-over-ranks timing
reduction used by bench.py works."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import numpy as np, torch, torch.distributed as dist
    from aircompressor_b200.sharding import shard_for_rank
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    sizes = np.arange(1, 1001, dtype=np.int64)
    b, e = shard_for_rank(sizes, rank, world)
    local = torch.tensor([int(sizes[b:e].sum()), e - b], dtype=torch.int64)
    dist.all_reduce(local, op=dist.ReduceOp.SUM)
    assert int(local[0]) == int(sizes.sum()) and int(local[1]) == 1000, local
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert float(t[0]) == float(world)
    dist.barrier()
    if rank == 0: print("GLOO_OK")
""") % ROOT


def test_world_size_2_gloo_sharding(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "GLOO_OK" in out.stdout

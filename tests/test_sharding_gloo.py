"""The N>1 path on CPU: two gloo ranks shard a frame list r::world, "process" their frames and gather
the per-frame point clouds to rank 0 through sharding.gather_point_clouds (the same code bench.py
runs over NCCL)."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, %(root)r)
    from laser_3d_reconstruction_b200 import sharding
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    nframes = 11
    def cloud(f):  # deterministic stand-in for a frame's point cloud (variable length, may be empty)
        n = (f * 7) %% 5
        return np.arange(n * 3, dtype=np.float64).reshape(n, 3) + 1000.0 * f
    mine = sharding.shard_frames(nframes, rank, world)
    table = sharding.pack_clouds(mine, [cloud(f) for f in mine])
    out = sharding.gather_point_clouds(table, device="cpu")
    if rank == 0:
        back = sharding.unpack_clouds(out, nframes)
        assert len(back) == nframes
        for f in range(nframes):
            assert np.array_equal(back[f], cloud(f)), f
        print("GATHER_OK", sum(len(b) for b in back))
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()
""")


def test_two_rank_gloo_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % dict(root=ROOT))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "GATHER_OK 20" in res.stdout

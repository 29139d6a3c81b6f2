"""world_size=2 gloo run of the multi-GPU host logic (sharding + counter reduction) on CPU."""
import os
import socket
import subprocess
import sys
import textwrap

from conftest import ROOT

_WORKER = textwrap.dedent(
    """
    import os, sys
    sys.path.insert(0, %r)
    import torch, torch.distributed as dist
    from mujoco_panda_pnp_b200 import distributed as D

    rank, local, world = D.init_process_group(backend="gloo")
    assert world == 2 and dist.is_initialized()
    n = 1001
    b, e = D.shard_range(n, rank, world)
    # each rank "solves" its shard: counters = [n, converged, success, sum(iterations)]
    idx = torch.arange(b, e)
    conv = (idx %% 10 != 0)
    local_counters = torch.tensor([e - b, int(conv.sum()), int(conv.sum()), int((idx %% 7 + 1).sum())], dtype=torch.int64)
    total = D.reduce_counters(local_counters)
    full = torch.arange(n)
    want = [n, int((full %% 10 != 0).sum()), int((full %% 10 != 0).sum()), int((full %% 7 + 1).sum())]
    assert total.tolist() == want, (total.tolist(), want)
    assert local_counters.tolist() != want  # reduce_counters does not modify its input
    t = D.reduce_max(1.0 + rank)
    assert t == 2.0
    D.barrier()
    s = D.summarize_ik(total)
    assert s["n"] == n and abs(s["success_rate"] - want[1] / n) < 1e-12
    dist.destroy_process_group()
    print("rank", rank, "ok")
    """
)


def test_two_rank_gloo_shard_and_reduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % ROOT)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=120)[0] for p in procs]
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, out
        assert f"rank {rank} ok" in out

"""What can the box's host side deliver?  Raw pinned cudaMemcpyAsync bandwidth with NO kernels, all ranks at the same
time: D2H only, H2D only, and both directions at once (the traffic pattern of the host-buffer IK operator: 12 B up and
48 B down per query).  Run under torchrun with 1 / 2 / 4 / 8 ranks; rank 0 prints one JSON line per run.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/dev/dev_pcie_multi.py >> profiles/pcie_multi_r2.log
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mujoco_panda_pnp_b200 import distributed as D  # noqa: E402

sys.path.insert(0, ROOT)
import bench  # noqa: E402  (bind_to_gpu_numa_node: the same CPU binding as the bench's e2e leg)


def main():
    rank, local_rank, world = D.init_process_group("nccl")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bench.bind_to_gpu_numa_node(local_rank) if world > 1 else None
    n = 1 << 24
    up_bytes, dn_bytes = n * 12, n * 48
    h_up = torch.empty(up_bytes, dtype=torch.uint8).pin_memory()
    h_dn = torch.empty(dn_bytes, dtype=torch.uint8).pin_memory()
    h_up.fill_(1)
    h_dn.fill_(2)  # first touch after the affinity bind
    d_up = torch.empty(up_bytes, dtype=torch.uint8, device=dev)
    d_dn = torch.empty(dn_bytes, dtype=torch.uint8, device=dev)
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

    def run(up, dn, reps=8):
        def once():
            if up:
                with torch.cuda.stream(s_up):
                    d_up.copy_(h_up, non_blocking=True)
            if dn:
                with torch.cuda.stream(s_dn):
                    h_dn.copy_(d_dn, non_blocking=True)
            s_up.synchronize()
            s_dn.synchronize()

        once()
        D.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            once()
        dt = D.reduce_max(time.perf_counter() - t0, dev) / reps
        return dt

    t_dn = run(False, True)
    t_up = run(True, False)
    t_both = run(True, True)
    if rank == 0:
        print(json.dumps({
            "n_gpus": world, "numa_bound_cpus": numa, "bytes_up_per_rank": up_bytes, "bytes_down_per_rank": dn_bytes,
            "d2h_only_gbs_total": dn_bytes * world / t_dn / 1e9, "h2d_only_gbs_total": up_bytes * world / t_up / 1e9,
            "both_d2h_gbs_total": dn_bytes * world / t_both / 1e9, "both_h2d_gbs_total": up_bytes * world / t_both / 1e9,
            "both_total_gbs": (up_bytes + dn_bytes) * world / t_both / 1e9,
            "ik_queries_per_s_ceiling": n * world / t_both,
        }))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()

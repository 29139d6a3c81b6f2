#!/usr/bin/env python
"""BASELINE cfg5: IK scaling sweep, N in {2^20 .. 2^28} cold reachable targets, with the NCCL reduction of the
success / iteration counters.  Run plain (1 GPU) or under torchrun.  Default: N targets PER GPU (weak scaling);
--total: N targets in all, rank r solves its contiguous shard of N / world (the 5 x 4 table of BASELINE cfg5).

    python tools/sweep_cfg5.py [--max-log2 28] [--total] > profiles/sweep_cfg5_rN.json
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from mujoco_panda_pnp_b200 import KinematicTree, _lib, engine, synthetic  # noqa: E402
from mujoco_panda_pnp_b200 import distributed as D  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--min-log2", type=int, default=20)
    ap.add_argument("--max-log2", type=int, default=28)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--total", action="store_true", help="N is the total over all GPUs (strong scaling)")
    args = ap.parse_args()
    rank, local_rank, world = D.init_process_group("nccl")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = _lib.load()
    tree = KinematicTree.from_mjcf()
    engine.set_tree(tree)
    neutral = torch.tensor(synthetic.NEUTRAL_Q, dtype=torch.float32, device=dev)
    params = engine.ik_params()
    stream = torch.cuda.current_stream().cuda_stream
    rows = []
    for log2n in range(args.min_log2, args.max_log2 + 1, 2):
        n = (1 << log2n) // world if args.total else 1 << log2n
        # generate targets in slices so the float64 generator temporaries stay small
        targets = torch.empty((n, 3), dtype=torch.float32, device=dev)
        step = 1 << 24
        for off in range(0, n, step):
            m = min(step, n - off)
            q = synthetic.random_joint_configs(m, tree.lower, tree.upper, seed=1234 + rank + 1000 * (off // step), device=dev)
            targets[off:off + m] = engine.fk_jac(q, want_quat=False, want_jac=False)[0]
            del q
        q8 = torch.empty((n, 8), dtype=torch.float32, device=dev)
        aux = torch.empty((n, 4), dtype=torch.float32, device=dev)
        cnt = torch.zeros(4, dtype=torch.int64, device=dev)

        def step_fn(c=None):
            _lib.check(lib.pnp_ik_solve_packed_f32(targets.data_ptr(), neutral.data_ptr(), 0, n, ctypes.byref(params),
                                                   q8.data_ptr(), aux.data_ptr(), c.data_ptr() if c is not None else None,
                                                   stream), "ik")

        for _ in range(3):
            step_fn()
        step_fn(cnt)
        torch.cuda.synchronize()
        D.barrier()
        ts = []
        for _ in range(args.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); step_fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = D.reduce_max(float(np.median(ts)), dev)
        c = D.reduce_counters(cnt).cpu().numpy()
        rows.append({"log2_n_total" if args.total else "log2_n_per_gpu": log2n, "n_per_gpu": n, "n_gpus": world, "ms": ms, "converged_solves_per_s": float(c[1]) / (ms * 1e-3),
                     "success_rate": float(c[2]) / float(c[0]), "mean_iterations": float(c[3]) / float(c[0]),
                     "alg_tflops_per_gpu": (500.0 * float(c[3]) + 216.0 * float(c[0])) / world / (ms * 1e-3) / 1e12})
        del targets, q8, aux
        torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps({"workload": "cfg5 cold IK sweep, targets = FK(U(limits)), q_init = neutral, defaults", "rows": rows}, indent=1))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()

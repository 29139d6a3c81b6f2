#!/bin/bash
# round 2, multi-GPU call (run with `gpurun --gpus N -- bash tools/dev/gpu_call_multi.sh N`): what the host side of the
# box delivers to N ranks at once, the cfg5 strong-scaling sweep at N ranks, and the bench line at N ranks.
set -u
N=${1:-2}
O=gpurun_out
mkdir -p $O
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" = "1" ]; then RUN1="python"; else RUN1="$RUN --master-port 29511"; fi
timeout 300 $RUN1 tools/dev/dev_pcie_multi.py > $O/pcie_multi_r2_n$N.json 2> $O/pcie_multi_r2_n$N.err; echo "pcie rc=$?"; cat $O/pcie_multi_r2_n$N.json
if [ "$N" = "1" ]; then RUN2="python"; else RUN2="$RUN --master-port 29512"; fi
timeout 900 $RUN2 tools/sweep_cfg5.py --total --max-log2 28 > $O/sweep_cfg5_total_r2_n$N.json 2> $O/sweep_cfg5_total_r2_n$N.err; echo "sweep rc=$?"; tail -c 600 $O/sweep_cfg5_total_r2_n$N.json
if [ "$N" = "1" ]; then RUN3="python"; else RUN3="$RUN --master-port 29513"; fi
timeout 900 $RUN3 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_r2_n$N.json 2> $O/bench_r2_n$N.err; echo "bench rc=$?"; tail -c 900 $O/bench_r2_n$N.json

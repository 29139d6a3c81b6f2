#!/bin/bash
# Run on the GPU box (via gpurun): bench line, ncu launch list of the same command, and one
# `--set full` capture per hot kernel.  Outputs land in gpurun_out/ (copied to profiles/ here).
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
python bench.py > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err || { echo "bench failed"; tail -20 $OUT/bench_${TAG}.err; exit 1; }
cat $OUT/bench_${TAG}.json
$CMD > $OUT/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_${TAG}.csv $CMD > $OUT/ncu_launches_${TAG}.log 2>&1
echo "launch list rc=$?"
SMALL="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --log2-n-ik 22 --log2-n-reward 24"
$SMALL > $OUT/plain_small_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ik_solve_v_kernel -s 3 -c 1 -f -o $OUT/ik_${TAG} $SMALL > $OUT/ncu_ik_${TAG}.log 2>&1
echo "ik full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:reward_kernel -s 3 -c 1 -f -o $OUT/reward_${TAG} $SMALL > $OUT/ncu_reward_${TAG}.log 2>&1
echo "reward full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:her_relabel_kernel -s 3 -c 1 -f -o $OUT/her_${TAG} $SMALL > $OUT/ncu_her_${TAG}.log 2>&1
echo "her full rc=$?"
ls -la $OUT

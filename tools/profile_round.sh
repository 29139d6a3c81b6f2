#!/bin/bash
# Run on the GPU box (via gpurun): bench line, ncu launch list of the same command, and one
# `--set full` capture per hot kernel.  Outputs land in gpurun_out/ (copied to profiles/ here).
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
python bench.py > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err || { echo "bench failed"; tail -20 $OUT/bench_${TAG}.err; exit 1; }
cat $OUT/bench_${TAG}.json
$CMD > $OUT/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_${TAG}.csv $CMD > $OUT/ncu_launches_${TAG}.log 2>&1
echo "launch list rc=$?"
SMALL="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --log2-n-ik 22 --log2-n-reward 24"
# every --set full capture sits behind ONE successful plain run of the same command: ncu on a program that
# faulted leaves the GPU unusable until a reset (B200_PROFILING.md)
if $SMALL > $OUT/plain_small_${TAG}.log 2>&1; then
  for K in ik_solve_v_kernel ik_solve_small_kernel reward_kernel her_relabel_kernel move_ik_plan_v_kernel; do
    SKIP=3; [ $K = move_ik_plan_v_kernel ] && SKIP=1   # the bench launches the planner three times in all
    RE="regex:$K"; BASE=function
    # a big IK solve is two launches of the same template (pair kernel + resume launch): pick the pair kernel by its type
    [ $K = ik_solve_v_kernel ] && { RE="regex:ik_solve_v_kernel<pnp_spec::F2"; BASE=demangled; }
    ncu --set full --clock-control none --import-source on --kernel-name-base $BASE -k "$RE" -s $SKIP -c 1 -f -o $OUT/${K}_${TAG} $SMALL > $OUT/ncu_${K}_${TAG}.log 2>&1
    echo "$K full rc=$?"
  done
else
  echo "plain run failed: no ncu captures"; tail -5 $OUT/plain_small_${TAG}.log
fi
ls -la $OUT

#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout=900 > $O/r2h_tests.log 2>&1; echo "tests rc=$?"; tail -15 $O/r2h_tests.log
timeout 600 python tools/dev/dev_pose.py > $O/r2h_pose.json 2> $O/r2h_pose.err; echo "pose rc=$?"; cat $O/r2h_pose.json

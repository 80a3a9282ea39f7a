#!/bin/bash
# GPU lab round: correctness matrix, then stage times of the bench workload for the default build and the variants
# found under csrc/_build/v_*/ (usage: gpurun -- bash tools/q_lab_run.sh)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/q_lab.py check > gpurun_out/check.log 2>&1; echo "check rc=$?"; grep -v "^\[fb\]" gpurun_out/check.log | grep -v '"q_equals_gen1": true' | tail -15
echo "== default build"; timeout 300 python tools/q_lab.py time 2>&1 | grep -v "^\[fb\]" | tee gpurun_out/time_default.log
for v in fast-barnes-py_b200/csrc/_build/v_*/libfastbarnes_b200.so; do
  [ -f "$v" ] || continue
  echo "== $v"; FB_LIB_PATH=$PWD/$v timeout 300 python tools/q_lab.py time 2>&1 | grep -v "^\[fb\]" | tee gpurun_out/time_$(basename $(dirname $v)).log
done

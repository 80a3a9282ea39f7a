#!/bin/bash
# GPU lab round: correctness matrix, then stage times of the bench workload for the default build and the variants
# found under csrc/_build/v_*/ (usage: gpurun -- bash tools/q_lab_run.sh [nocheck])
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export QLAB_REF_OUT=/tmp/qlab_ref.npy
rm -f $QLAB_REF_OUT
if [ "$1" != "nocheck" ]; then
  timeout 300 python tools/q_lab.py check > gpurun_out/check.log 2>&1; echo "check rc=$?"; grep -v "^\[fb\]" gpurun_out/check.log | grep -v '"q_equals_gen1": true' | tail -15
fi
echo "== default build"; timeout 300 python tools/q_lab.py time 2>&1 | grep -v "^\[fb\]" | tee gpurun_out/time_default.log
DEF='[{"sweepq":1}]'
VS="${QLAB_VARIANT_SETTINGS:-$DEF}"
for v in fast-barnes-py_b200/csrc/_build/v_*/libfastbarnes_b200.so; do
  [ -f "$v" ] || continue
  echo "== $v"; QLAB_SETTINGS="$VS" FB_LIB_PATH=$PWD/$v timeout 300 python tools/q_lab.py time 2>&1 | grep -v "^\[fb\]" | tee gpurun_out/time_$(basename $(dirname $v)).log
done

#!/bin/bash
# compute-sanitizer over the smoke test and a few small GPU parity cases (memcheck, racecheck, synccheck); summaries go to
# gpurun_out/sanitizer_<tool>.txt.  Usage on the GPU box:  bash tools/sanitize.sh
# racecheck / synccheck watch shared memory and barriers; the TMA (async proxy) traffic of the q kernels is outside what
# racecheck tracks, the mbarrier waits are what orders it.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CASES='test_small_volumes_many_times or test_barnes_golden or test_line_kernel_golden or test_sweepq_wide_kernels_vs_oracle or test_sweepq_vs_first_generation or test_1d_exact_default_and_segmented_option or test_z_slab_decomposition_single_gpu or test_injection_lists_vs_segments'
for tool in ${TOOLS:-memcheck racecheck synccheck}; do
  out=gpurun_out/sanitizer_$tool.txt
  {
    echo "== compute-sanitizer --tool $tool: __graft_entry__.smoke()"
    timeout 900 compute-sanitizer --tool $tool --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|=========.*(Error|error|hazard|Invalid)|smoke" | head -20
    echo "== compute-sanitizer --tool $tool: pytest -m gpu -k '$CASES'"
    timeout 1500 compute-sanitizer --tool $tool --print-limit 5 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$CASES" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|=========.*(Error|error|hazard|Invalid)|passed|failed" | head -20
  } > $out 2>&1
  echo "--- $out"; cat $out
done

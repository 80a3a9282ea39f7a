#!/bin/bash
# SASS evidence for the kernels of the hot path: per kernel the count of tensor-memory (LDTM / STTM), TMA (UTMALDG / UTMASTG),
# mbarrier (SYNCS), cp.async (LDGSTS), fp64 (DADD / DMUL / DFMA) and shared-memory instructions, registers and spills.
# Usage: bash tools/sass_counts.sh > profiles/r2_sass_counts.txt   (needs only cuobjdump; no GPU)
cd "$(dirname "$0")/.."
LIB=fast-barnes-py_b200/csrc/_build/libfastbarnes_b200.so
LOG=$LIB.ptxas.log
echo "cuobjdump -sass $LIB (sm_100a), instruction counts per kernel (whole kernel: fast path + line-end path + prologue)"
for k in 'fb_sweepq_kernelILi4ELi1ELi1E' 'fb_sweepq_kernelILi4ELi1ELi2E' 'fb_sweepq_kernelILi4ELi1ELi0E' 'fb_sweepqs_kernelILi4ELi1E' \
         'fb_sweepp_kernelILi4ELi1E' 'fb_sweepp_kernelILi4ELi2E' 'fb_line1d_kernelILi4ELi2E' 'fb_sweep32_kernelILi4ELi2E' 'fb_sweep_kernelILi4ELi2ELi8E'; do
  name=$(cuobjdump -sass $LIB | grep -o "Function : [A-Za-z0-9_]*$k[A-Za-z0-9_]*" | head -1 | sed 's/Function : //')
  [ -z "$name" ] && continue
  echo; echo "== $(echo $name | c++filt)"
  grep -A3 "Compiling entry function '$name'" $LOG | grep -E "Used|spill" | sed 's/^ptxas info    : /   /'
  cuobjdump -sass -fun "$name" $LIB 2>/dev/null | grep -oE "^\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T]+\s+)?[A-Z0-9_]+(\.[A-Za-z0-9_]+)*" | awk '{print $NF}' \
    | grep -E "^(LDTM|STTM|UTMALDG|UTMASTG|UTMAPF|SYNCS|LDGSTS|DADD|DMUL|DFMA|LDS|STS|SHFL|BAR|ELECT|REDUX|MUFU\.RCP64H|UTCATOMSWS|USETMAXREG)" \
    | sort | uniq -c | awk '{printf "   %6d %s\n", $1, $2}'
done

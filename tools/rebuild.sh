#!/bin/bash
# rebuild the library in-tree and record the source digest (so that the GPU box does not rebuild it)
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
make -C "$ROOT/fast-barnes-py_b200/csrc" "$@" 2>&1 | tail -2
cd "$ROOT" && python -c "
import sys; sys.path.insert(0, 'fast-barnes-py_b200')
from fastbarnes import _lib; _lib.mark_built(); print('stale:', _lib.is_stale())"

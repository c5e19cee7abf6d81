#!/bin/bash
# Rebuilds the library; non-zero exit (and the compiler output) when nvcc fails.
cd "$(dirname "$0")/.." && python -c "
import sys; sys.path.insert(0, '.')
from graph_peak_caller_b200 import build
build.build_library(force=True, verbose='-v' in sys.argv)
print('built')" "$@"

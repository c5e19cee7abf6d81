#!/bin/sh
# TEST INFRASTRUCTURE (build container only): runs the reference's own unit
# tests for the callpeaks path, unmodified, on top of oracle/refshim.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
TMP=$(mktemp -d)
cat > "$TMP/conftest.py" <<PY
import sys
sys.path.insert(0, "$HERE/refshim")
import refcompat; refcompat.install()
sys.path.insert(0, "/root/reference/tests")
PY
for t in test_create_sample_pileup test_create_control_whole test_pvalues \
         test_sparseholescleaning test_whole_callpeaks test_call_peaks_from_qvalues \
         test_filter_duplicates test_sparsediffs test_linearpileup testeventsort \
         test_haplotyping test_linearfilter test_multigraph_shiftestimation; do
    cp /root/reference/tests/$t.py "$TMP/"
done
cp /root/reference/tests/util.py "$TMP/"
cd "$TMP" && python -m pytest -q -p no:cacheprovider *.py

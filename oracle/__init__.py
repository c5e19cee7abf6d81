"""ORACLE -- test infrastructure, not product code.

A CPU restatement (numpy / scipy / plain Python loops) of the reference's
callpeaks hot path (uio-bmi/graph_peak_caller v1.2.3), operating on the same
flat arrays the CUDA path consumes.  Every function cites the reference
file:line it follows.

Who may use it: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs -- as the checker or as the timed CPU baseline, never
as a fallback of the product.  graph_peak_caller_b200/ does not import it.

Parity status: PINNED.  The restatement is compared with the *unmodified*
reference (imported through oracle/refshim, a stand-in for the absent
third-party ``offsetbasedgraph``) on the reference's own unit-test inputs and
on randomised graphs by tests/test_oracle_vs_reference.py (runs wherever
/root/reference exists), and with vectors generated from that reference and
committed under tests/golden/ (tests/test_oracle_golden.py, runs everywhere).
Exceptions: oracle/ingest.py (vg JSON -> reads / graph) restates the absent
third-party package pyvg and is UNPINNED against it; its header lists the
anchors the reference tree offers instead (its chromosome splitter, on the other
hand, is pinned: byte-identical files with the live reference's).  oracle/shift.py
(fragment-length estimation) is pinned against the live PeakModel / HaploTyper /
LinearFilter and the reference's own known answers, except for ONE rule taken from the
absent offsetbasedgraph: the linear offset of a reverse-strand read start (half
pinned, see its header).  oracle/summits.py is pinned against the live
PeakCollection.cut_around_summit and the reference's known answer.
"""

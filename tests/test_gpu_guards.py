"""GPU: the whole path in GUARD MODE (GPC_GUARD=1), the stand-in for compute-sanitizer's
memcheck / initcheck on a pool where that tool is closed (profiles/README.md).

In a child process -- the mode is fixed when the library is loaded -- every scratch buffer
of the library and every output array is poisoned with 0xA5 and fenced by 256-byte canaries:
  * a kernel that writes outside its buffer trips a canary (counted, named on stderr);
  * a kernel that reads memory nobody wrote computes on poison, and the parity checks of the
    same run (integer tracks and peaks bit-exact against the oracle) fail.
Cases: random bubble graphs through every sample-pileup route, control with and without
MNP scales (both linear-track paths), whole passes with hole cleaning and max paths, a
wide fan-out (frontier spill pool), cyclic graphs, three chromosomes with a joined table."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import sys
import numpy as np
sys.path.insert(0, "tests")
from graph_peak_caller_b200 import _lib, device, pipeline, kernels
from graph_peak_caller_b200.control.linearmap import LinearMap
from graph_peak_caller_b200.device import DeviceGraph, DeviceReads
from graph_peak_caller_b200.synthetic import make_graph, make_reads
from oracle import control as ocontrol, pipeline as opipeline
import test_gpu_sample, test_cyclic
assert device.GUARD and _lib.load().gpc_guard_mode() == 1
checked = 0
def fence(what):
    global checked
    bad = device.check_guards()
    assert bad == 0, "%s: %d buffers written outside their bounds" % (what, bad)
    checked += 1
# sample pileup, every route (difference array, event list, two-call form, page-locked upload)
for trial in range(4):
    test_gpu_sample.test_random_graphs(trial)
    fence("sample %d" % trial)
test_gpu_sample.test_wide_fan_out_frontier(0)
fence("fan-out")
for case in ("self_loop", "two_node", "padded", 0, 1):
    test_cyclic.test_worklist_kernel_vs_oracle(case)
    fence("cyclic %s" % case)
# whole passes against the oracle
cpu = lambda t: t.cpu().numpy()
for trial, (mnp, has_control) in enumerate(((0.0, True), (0.3, True), (0.0, False), (0.2, True))):
    g = make_graph(20000 + 9000 * trial, 0.04, seed=300 + trial, mnp_frac=mnp, indel_mean=6.0)
    R, F = 12, 70
    s = make_reads(g, 3000, R, F, seed=310 + trial, peak_frac=0.6, peak_spacing=3000)
    c = make_reads(g, 3500, R, F, seed=320 + trial, peak_frac=0.0, dup_frac=0.0)
    lm = LinearMap.from_graph(g)
    dg = DeviceGraph(g, lm.as_arrays())
    st, qt, q, thr, cleaned, peaks = pipeline.callpeaks(dg, DeviceReads(s), DeviceReads(c), R, F,
                                                        has_control, None, True)
    fence("pass %d" % trial)
    ref = opipeline.callpeaks(g, ocontrol.linear_map_from_graph(g), opipeline.unique_reads(s),
                              opipeline.unique_reads(c), R, F, has_control=has_control)
    assert np.array_equal(cpu(st.sample[0]).view(np.uint32), ref["sample"][0])
    assert np.array_equal(cpu(st.sample[1]), ref["sample"][1])
    assert np.array_equal(cpu(cleaned[0]).view(np.uint32), ref["hole_cleaned"][0])
    assert np.array_equal(cpu(cleaned[1]).astype(bool), np.asarray(ref["hole_cleaned"][1]).astype(bool))
    got = sorted((p[0], p[1], tuple(p[2])) for p in peaks.final_peaks())
    want = sorted((p[0], p[1], tuple(p[2])) for p in ref["peaks"])
    assert got == want and len(want) > 0
print("GUARD OK fences=%d violations=%d" % (checked, int(_lib.load().gpc_guard_violations())))
'''


@pytest.mark.gpu
def test_whole_path_in_guard_mode():
    env = dict(os.environ, GPC_GUARD="1")
    out = subprocess.run([sys.executable, "-c", CHILD], capture_output=True, text=True, timeout=1500,
                         cwd=ROOT, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-6000:]
    assert "GUARD OK" in out.stdout and "violations=0" in out.stdout
    assert "GUARD:" not in out.stderr

"""Loader for tests/golden/*.npz (written by oracle/make_golden.py from the
unmodified reference)."""
import glob
import json
import os

import numpy as np

from graph_peak_caller_b200.graph import Graph
from graph_peak_caller_b200.reads import ReadBatch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def case_names():
    return sorted(os.path.basename(p)[:-4]
                  for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


class GoldenCase:
    def __init__(self, name):
        o = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        n = o["node_len"].size
        src = np.repeat(np.arange(n), np.diff(o["succ_ptr"]))
        self.graph = Graph.from_edges(o["node_len"], src, o["succ"],
                                      int(o["min_node"]))
        self.sample = ReadBatch(*[o["s_" + f] for f in (
            "start_node", "start_off", "end_node", "end_off", "path_ptr", "path")])
        self.control = ReadBatch(*[o["c_" + f] for f in (
            "start_node", "start_off", "end_node", "end_off", "path_ptr", "path")])
        self.linear_map = (o["lin_starts"], o["lin_ends"])
        self.meta = json.loads(str(o["meta"]))
        self.touched = o["touched"]
        self.p2q = (o["p2q_p"], o["p2q_q"])
        self._o = o

    def track(self, key):
        return self._o[key + "_idx"], self._o[key + "_val"]

    @property
    def peaks(self):
        return [(p[0], p[1], p[2], tuple(p[3]), p[4]) for p in self.meta["peaks"]]

    @property
    def all_max_paths(self):
        return [(p[0], p[1], p[2]) for p in self.meta["all_max_paths"]]

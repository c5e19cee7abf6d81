"""``Reporter``: the on-disk contract of the reference
(graph_peak_caller/reporter.py:7-65).  ``add(name, data)`` writes what the
reference writes under the same file names:

* tracks -> ``<base><name>_indexes.npy`` / ``<base><name>_values.npy``
  (``to_sparse_files``): qvalues, pvalues, hole_cleaned, fragment_pileup,
  background_track, direct_pileup;
* peak lists -> ``<base><name>.intervalcollection`` (one JSON object per line):
  all_max_paths, max_paths;
* touched_nodes -> ``<base>touched_nodes.npy``;
* sub_graphs -> ``<base>sub_graphs.graphs.npz`` / ``<base>sub_graphs.nodeids.npz``, one
  entry ``peak<i>`` per final peak (reporter.py:11-15): the matrices and node ids of
  ``postprocess/graphs.py``; the list is lazy, so a reporter that keeps things in memory
  costs the pass nothing.

Anything else is kept but not written -- "thresholded" among them: the
reference's writer for it is spelt ``thersholded`` and is never found by its
``add`` (reporter.py:41-42,52-57).  A reporter with
``base_name=None`` keeps everything in memory only, which is what the benchmark
uses; ``memory`` holds the last object reported under every name either way."""
import logging

import numpy as np

from .peakcollection import PeakCollection

TRACKS = frozenset(("qvalues", "pvalues", "hole_cleaned", "fragment_pileup",
                    "background_track", "direct_pileup"))
PEAK_LISTS = frozenset(("all_max_paths", "max_paths"))


class Reporter:
    def __init__(self, base_name):
        self._base_name = base_name
        self.memory = {}

    def _write(self, name, data):
        if name in TRACKS:
            data.to_sparse_files(self._base_name + name)
        elif name in PEAK_LISTS:
            PeakCollection(data).to_file(self._base_name + name + ".intervalcollection",
                                         text_file=True)
        elif name == "sub_graphs":
            np.savez(self._base_name + "sub_graphs.graphs",
                     **{"peak%s" % i: p._graph for i, p in enumerate(data)})
            np.savez(self._base_name + "sub_graphs.nodeids",
                     **{"peak%s" % i: p._node_ids for i, p in enumerate(data)})
        elif name == "touched_nodes":
            ids = data.ids() if hasattr(data, "ids") else list(data)      # TouchedNodes: no list
            np.save(self._base_name + "touched_nodes.npy", np.array(ids, dtype="int"))
        else:
            return False
        return True

    def add(self, name, data):
        self.memory[name] = data
        if self._base_name is not None and self._write(name, data):
            logging.info("Wrote %s to file", name)
        else:
            logging.info("Skipping reporting of %s", name)

    def get_sub_reporter(self, name):
        """Reporter of one chromosome: ``<base><name>_`` (the empty name adds nothing)."""
        if self._base_name is None:
            return self.__class__(None)
        return self.__class__(self._base_name + (name + "_" if name != "" else ""))

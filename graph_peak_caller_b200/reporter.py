"""``Reporter``: the on-disk contract of the reference
(graph_peak_caller/reporter.py:7-65).  Every intermediate track goes to
``<base>X_indexes.npy`` / ``<base>X_values.npy``; peaks to JSON lines.  A
reporter with ``base_name=None`` keeps everything in memory (nothing is
written), which is what the benchmark uses."""
import logging

import numpy as np

from .peakcollection import PeakCollection


class Reporter:
    def __init__(self, base_name):
        self._base_name = base_name
        self.memory = {}

    def qvalues(self, data):
        data.to_sparse_files(self._base_name + "qvalues")

    def pvalues(self, data):
        data.to_sparse_files(self._base_name + "pvalues")

    def all_max_paths(self, data):
        PeakCollection(data).to_file(self._base_name + "all_max_paths.intervalcollection",
                                     text_file=True)

    def max_paths(self, data):
        PeakCollection(data).to_file(self._base_name + "max_paths.intervalcollection",
                                     text_file=True)

    def hole_cleaned(self, data):
        data.to_sparse_files(self._base_name + "hole_cleaned")

    def fragment_pileup(self, data):
        data.to_sparse_files(self._base_name + "fragment_pileup")

    def background_track(self, data):
        data.to_sparse_files(self._base_name + "background_track")

    def thersholded(self, data):      # sic: the reference never writes "thresholded"
        data.to_sparse_files(self._base_name + "thresholded")

    def direct_pileup(self, data):
        data.to_sparse_files(self._base_name + "direct_pileup")

    def touched_nodes(self, data):
        np.save(self._base_name + "touched_nodes.npy", np.array(list(data), dtype="int"))

    def add(self, name, data):
        self.memory[name] = data
        if self._base_name is not None and hasattr(self, name) and name != "memory":
            getattr(self, name)(data)
            logging.info("Wrote %s to file", name)
        else:
            logging.info("Skipping reporting of %s", name)

    def get_sub_reporter(self, name):
        if name != "":
            name += "_"
        base = None if self._base_name is None else self._base_name + name
        return self.__class__(base)

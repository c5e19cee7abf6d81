"""``Configuration``, ``CallPeaks`` and ``CallPeaksFromQvalues`` with the
reference's signatures and public state (graph_peak_caller/callpeaks.py:14-218).
The bodies are the kernel chain of pipeline.py; nothing here computes on the
host."""
import logging

import numpy as np

from .control import get_background_track_from_control, get_background_track_from_input, \
    scale_tracks
from .graph import Graph
from .postprocess import HolesCleaner, SparseMaxPaths
from .sample import get_fragment_pileup
from .sparsediffs import SparseValues
from .sparsepvalues import PToQValuesMapper, PValuesFinder, QValuesFinder


class Configuration:
    def __init__(self):
        self.read_length = None
        self.fragment_length = None
        self.linear_map_name = None
        self.has_control = False
        self.q_values_threshold = 0.05
        self.global_min = None
        self.keep_duplicates = False

    def copy(self):
        o = Configuration()
        o.__dict__.update(self.__dict__)
        # the reference's copy() drops linear_map_name (typo, callpeaks.py:28);
        # every caller sets it again right after, so behave the same way
        o.linear_map_name = None
        return o


class CallPeaks(object):
    def __init__(self, graph, config, reporter, variant_maps=None):
        self.graph = Graph.from_obg(graph)
        self.config = config
        assert self.config.fragment_length > self.config.read_length, \
            "Read length %d is larger than fragment size" % (self.config.read_length)
        self.info = config
        self._reporter = reporter
        self.variant_maps = variant_maps
        self.direct_pileup = None

    def run_pre_callpeaks(self, input_reads, control_reads):
        if hasattr(control_reads, "prefetch") and control_reads is not input_reads:
            # sample upload first (it is needed first and gets the whole link); the
            # control copies queue behind it on the copy stream and overlap the
            # sample stage instead of competing with the sample upload
            if hasattr(input_reads, "_device_reads"):
                input_reads._device_reads()
            control_reads.prefetch()
        sample_pileup = get_fragment_pileup(self.graph, input_reads, self.config,
                                            self._reporter)
        self.direct_pileup = sample_pileup.direct_pileup
        background_func = get_background_track_from_control
        if not self.config.has_control:
            background_func = get_background_track_from_input
        control_pileup = background_func(self.graph, control_reads, self.config,
                                         sample_pileup.touched_nodes)
        scale_tracks(sample_pileup, control_pileup,
                     input_reads.n_reads / control_reads.n_reads)
        self._reporter.add("fragment_pileup", sample_pileup)
        self._reporter.add("touched_nodes", sample_pileup.touched_nodes)
        self._reporter.add("background_track", control_pileup)
        self.touched_nodes = sample_pileup.touched_nodes
        self.control_pileup = control_pileup
        self.sample_pileup = sample_pileup

    def get_p_values(self):
        assert self.sample_pileup is not None
        assert self.control_pileup is not None
        self.p_values_pileup = PValuesFinder(
            self.sample_pileup, self.control_pileup).get_p_values_pileup()
        self.p_values_pileup.track_size = self.graph.node_indexes[-1]
        self._reporter.add("pvalues", self.p_values_pileup)
        self.sample_pileup = None
        self.control_pileup = None

    def get_p_to_q_values_mapping(self):
        assert self.p_values_pileup is not None
        finder = PToQValuesMapper.from_p_values_pileup(self.p_values_pileup)
        self.p_to_q_values_mapping = finder.get_p_to_q_values()

    def get_q_values(self):
        assert self.p_values_pileup is not None
        assert self.p_to_q_values_mapping is not None
        finder = QValuesFinder(self.p_values_pileup, self.p_to_q_values_mapping)
        self.q_values_pileup = finder.get_q_values()
        self.q_values_pileup.track_size = self.p_values_pileup.track_size
        self._reporter.add("qvalues", self.q_values_pileup)

    def call_peaks_from_q_values(self, linear_path=None):
        assert self.q_values_pileup is not None
        caller = CallPeaksFromQvalues(
            self.graph, self.q_values_pileup, self.config, self._reporter,
            raw_pileup=self.direct_pileup, touched_nodes=self.touched_nodes,
            config=self.config, linear_path=linear_path, variant_maps=self.variant_maps)
        caller.callpeaks()
        self.max_path_peaks = caller.max_paths

    def run(self, input_intervals, control_intervals):
        self.run_to_p_values(input_intervals, control_intervals)
        self.get_p_to_q_values_mapping()
        self.get_q_values()
        self.call_peaks_from_q_values()

    def run_to_p_values(self, input_intervals, control_intervals):
        self.run_pre_callpeaks(input_intervals, control_intervals)
        self.get_p_values()


class CallPeaksFromQvalues:
    def __init__(self, graph, q_values_pileup, experiment_info, reporter, cutoff=0.1,
                 raw_pileup=None, touched_nodes=None, config=None, q_values_max_path=False,
                 linear_path=None, variant_maps=None):
        self.graph = Graph.from_obg(graph)
        self.q_values = q_values_pileup
        self.info = experiment_info
        self.cutoff = cutoff
        self.raw_pileup = raw_pileup
        self.touched_nodes = touched_nodes
        self.q_values_max_path = q_values_max_path
        self.linear_path = linear_path
        self.variant_maps = variant_maps
        if config is not None:
            self.cutoff = config.q_values_threshold
        self._reporter = reporter
        logging.info("Using q value cutoff %.4f" % self.cutoff)

    def _threshold(self):
        threshold = -np.log10(self.cutoff)
        logging.info("Thresholding peaks on q value %.4f" % threshold)
        self.pre_processed_peaks = self.q_values.threshold_copy(threshold)
        self._reporter.add("thresholded", self.pre_processed_peaks)

    def _postprocess(self):
        logging.info("Filling small Holes")
        self.pre_processed_peaks = HolesCleaner(
            self.graph, self.pre_processed_peaks, self.info.read_length,
            self.touched_nodes).run()
        self._reporter.add("hole_cleaned", self.pre_processed_peaks)
        self.filtered_peaks = self.pre_processed_peaks

    def _get_max_paths(self):
        logging.info("Getting maxpaths")
        if self.q_values_max_path:
            score_track = self.q_values
        else:
            if self.raw_pileup is None:
                # the reference reads it back from the Reporter's file (callpeaks.py:170-173)
                file_name = self._reporter._base_name + "direct_pileup"
                logging.info("Reading raw direct pileup from file %s" % file_name)
                self.raw_pileup = SparseValues.from_sparse_files(file_name)
            score_track = self.raw_pileup
        finder = SparseMaxPaths(self.filtered_peaks, self.graph, score_track, self.variant_maps)
        finder.set_q_values(self.q_values)
        max_paths, sub_graphs = finder.run()
        chromosome = None if self._reporter._base_name is None \
            else self._reporter._base_name.replace("_", "")
        # scores are floats already; peaks of non-zero length carry the chromosome
        # (callpeaks.py:191-200) -- applied when a Peak object is made
        max_paths.set_chromosome(chromosome)
        self._reporter.add("all_max_paths", max_paths)
        order = finder.order     # stable sort by score, descending (callpeaks.py:204-205)
        logging.info("N unfiltered peaks: %s", len(max_paths))
        length = finder.peak_set.length
        keep = order[length[order] >= self.info.fragment_length] if len(order) else order
        self.max_paths = max_paths.subset(keep)
        logging.info("N filtered peaks: %s", len(self.max_paths))
        self._reporter.add("sub_graphs", sub_graphs.subset(keep))      # lazy (callpeaks.py:210)
        self._reporter.add("max_paths", self.max_paths)

    def callpeaks(self):
        logging.info("Calling peaks")
        self._threshold()
        self._postprocess()
        self._get_max_paths()

"""``MultipleGraphsCallpeaks`` with the reference's constructor and methods
(graph_peak_caller/multiplegraphscallpeaks.py:14-168): one graph per
chromosome, p-values per chromosome, ONE q mapping over all chromosomes,
peaks per chromosome.

The reference runs the chromosomes one after the other (or as separate OS
processes joined through ``*_pvalues_*.npy`` files).  Here the chromosomes are
distributed over the ranks of ``torch.distributed`` (one process per GPU,
longest-processing-time bin packing by graph size) and the join is one
all-gather of the (p, base pairs) tables (multigraph.gather_p_tables).  In a
single process it degenerates to the reference's sequential loop.

Inputs per chromosome: a ``Graph`` (or a file written by ``Graph.to_file``, or
a vg graph in JSON form), reads as ``Intervals`` / ``UniqueIntervals`` /
``ReadBatch`` / iterable of obg-style intervals or a file name
(``*.intervalcollection`` text, vg alignment JSON, or ``ReadBatch.to_file``'s
``.npz``; the text formats go through the native host parsers of SURVEY.md
§8f-1), a linear map (``LinearMap`` or ``.npz`` file).  obg's binary ``.nobg``
graph files need offsetbasedgraph itself and are not read here."""
import logging

import numpy as np

from . import kernels, multigraph
from .callpeaks import CallPeaks
from .graph import Graph
from .intervals import Intervals, UniqueIntervals
from .reads import ReadBatch
from .sparsediffs import SparseValues
from .sparsepvalues import PToQValuesMapper


class MultipleGraphsCallpeaks:
    def __init__(self, graph_names, graph_file_names, samples, controls, linear_maps, config,
                 reporter, sequence_retrievers=None, stop_after_p_values=False,
                 linear_path_file_names=None, variant_maps_path=None):
        if variant_maps_path is not None:
            raise NotImplementedError("variant maps are outside this path")
        self._config = config
        self._reporter = reporter
        self.names = list(graph_names)
        self.graph_file_names = list(graph_file_names)
        # samples / controls / linear maps are None when only the second half runs
        # (run_callpeaks_whole_genome_from_p_values, callpeaks_interface.py:340-371)
        self.linear_maps = None if linear_maps is None else list(linear_maps)
        self.sequence_retrievers = sequence_retrievers
        self.samples = None if samples is None else list(samples)
        self.controls = None if controls is None else list(controls)
        self.stop_after_p_values = stop_after_p_values
        self.linear_path_file_names = linear_path_file_names
        self._graphs = {}
        self._callers = {}
        self._mine = None
        if self.stop_after_p_values:
            logging.info("Will only run until p-values have been computed.")

    # -- helpers ---------------------------------------------------------------
    def _graph(self, i):
        if i not in self._graphs:
            g = self.graph_file_names[i]
            if isinstance(g, str):
                from .conversion import load_graph
                self._graphs[i] = load_graph(g)
            else:
                self._graphs[i] = Graph.from_obg(g)
        return self._graphs[i]

    def _costs(self):
        """Packing weight per chromosome, identical on every rank: graph size in
        base pairs -- or, when every graph is a file, the file sizes, so that a
        rank only ever parses the graphs it works on."""
        import os
        names = self.graph_file_names
        if all(isinstance(g, str) and os.path.isfile(g) for g in names):
            kinds = {g.rsplit(".", 1)[-1] for g in names}
            if len(kinds) == 1:                      # sizes of one format compare
                return [os.path.getsize(g) for g in names]
        return [self._graph(i).size_bp for i in range(len(names))]

    def my_chromosomes(self):
        """Indexes of the chromosomes this rank works on."""
        if self._mine is None:
            rank, world = multigraph.world()
            if world == 1:
                self._mine = list(range(len(self.names)))
            else:
                self._mine = sorted(multigraph.assign_chromosomes(self._costs(), world)[rank])
        return self._mine

    @classmethod
    def count_number_of_unique_reads(cls, sample_reads):
        """reference multiplegraphscallpeaks.py:41-63, on the device."""
        from .device import device_reads
        n_unique = 0
        for reads in sample_reads:
            batch = reads if isinstance(reads, ReadBatch) else ReadBatch.from_intervals(
                getattr(reads, "intervals", reads))
            _, n = kernels.unique_reads(device_reads(batch))
            n_unique += n
        logging.info("In total %d unique reads" % n_unique)
        return n_unique

    def get_intervals(self, sample, control, graph):
        """File names as in the reference (multiplegraphscallpeaks.py:78-104):
        ``*.intervalcollection`` is an interval text file, any other name a vg
        alignment JSON file read against ``graph``; additionally ``*.npz`` is a
        batch written by ``ReadBatch.to_file``."""
        def load(x):
            if isinstance(x, (Intervals, UniqueIntervals)):
                return x
            if isinstance(x, str):
                if x.endswith(".npz"):
                    x = ReadBatch.from_file(x)
                elif x.endswith(".intervalcollection"):
                    x = ReadBatch.from_intervalcollection_file(x)
                else:
                    x = ReadBatch.from_vg_json_file(x, graph)
            if self._config.keep_duplicates:
                logging.warning("Keeping duplicates. Should only be used for testing.")
                return Intervals(x)
            return UniqueIntervals(x)
        return load(sample), load(control)

    # -- reference surface -------------------------------------------------------
    def run(self):
        self.run_to_p_values()
        if self.stop_after_p_values:
            logging.info("Stopping, as planned, after p-values")
            return
        self.create_joined_q_value_mapping()
        self.run_from_p_values()

    def run_to_p_values(self):
        mine = self.my_chromosomes()

        def load(i):
            graph = self._graph(i)
            return (graph,) + tuple(self.get_intervals(self.samples[i], self.controls[i], graph))

        ahead = load(mine[0]) if mine else None
        for k, i in enumerate(mine):
            name = self.names[i]
            logging.info("Running to p values, %s" % name)
            graph, sample, control = ahead
            ahead = None
            if k + 1 < len(mine):
                # the reads of the next chromosome cross the host link (copy stream) while
                # this one is being computed
                ahead = load(mine[k + 1])
                for reads in ahead[1:]:
                    reads.prefetch()
            config = self._config.copy()
            config.linear_map_name = self.linear_maps[i]
            caller = CallPeaks(graph, config, self._reporter.get_sub_reporter(name))
            caller.run_to_p_values(sample, control)
            self._callers[i] = caller
            logging.info("In total %d duplicates were removed from sample" % sample.n_duplicates)

    def create_joined_q_value_mapping(self):
        if not self._callers and self.samples is None:
            # second half on its own: the p-values of every chromosome are in the files the
            # first halves wrote (reference multiplegraphscallpeaks.py:122-124)
            mapper = PToQValuesMapper.from_files(self._reporter._base_name)
            self._q_value_mapping = mapper.get_p_to_q_values()
            return
        torch = kernels.torch_cuda()
        tps, tcs, total = [], [], 0
        for i in self.my_chromosomes():
            pv = self._callers[i].p_values_pileup
            pos, p = pv.device_arrays()
            tp, tc = kernels.ptable_local(pos, p, int(pv.track_size))
            tps.append(tp)
            tcs.append(tc)
            total += int(pv.track_size)
        if tps:
            tp, tc = torch.cat(tps), torch.cat(tcs)
        else:
            from .device import empty
            tp, tc = empty(0, torch.float64), empty(0, torch.int64)
        tp, tc, total = multigraph.gather_p_tables(tp, tc, total)
        mapper = PToQValuesMapper._from_device_table(tp, tc, total)
        self._q_value_mapping = mapper.get_p_to_q_values()

    def _caller_from_files(self, i):
        """State after p-values read back from the Reporter's files (reference
        multiplegraphscallpeaks.py:147-156); the direct pileup is read by
        CallPeaksFromQvalues when it is needed (callpeaks.py:170-173)."""
        name = self.names[i]
        caller = CallPeaks(self._graph(i), self._config, self._reporter.get_sub_reporter(name))
        prefix = self._reporter._base_name + (name + "_" if name != "" else "")
        caller.p_values_pileup = SparseValues.from_sparse_files(prefix + "pvalues")
        caller.touched_nodes = np.load(prefix + "touched_nodes.npy")
        return caller

    def run_from_p_values(self, only_chromosome=None):
        for i in self.my_chromosomes():
            name = self.names[i]
            if only_chromosome is not None and only_chromosome != name:
                logging.info("Skipping %s" % str(name))
                continue
            if i not in self._callers:
                self._callers[i] = self._caller_from_files(i)
            caller = self._callers[i]
            caller.p_to_q_values_mapping = self._q_value_mapping
            caller.get_q_values()
            caller.call_peaks_from_q_values()
            self._write_sequences(i, caller)

    def _sequence_graph(self, i):
        """The sequence graph of chromosome i, or None.  ``sequence_retrievers`` may be
        indexable by chromosome (list, or an object with ``get(i)``: what the entry points
        of callpeaks_interface.py pass) or, as in the reference, an iterator consumed in
        chromosome order (single process only)."""
        source = self.sequence_retrievers
        if source is None:
            return None
        if hasattr(source, "get"):
            return source.get(i)
        if isinstance(source, (list, tuple)):
            return source[i]
        return next(source)

    def _write_sequences(self, i, caller):
        """``<base><name>_sequences.fasta`` (reference multiplegraphscallpeaks.py:157-166);
        a missing sequence file is a warning, not an error."""
        if self.sequence_retrievers is None or self._reporter._base_name is None:
            return
        try:
            sequence_graph = self._sequence_graph(i)
        except FileNotFoundError:
            logging.warning("Could not find sequence graphs. Will not store max paths.")
            return
        except (ValueError, KeyError, OSError) as e:
            # e.g. offsetbasedgraph's own .sequences file: not this package's flat format
            logging.warning("Could not read the sequence graph (%s). Will not store max paths." % e)
            return
        if sequence_graph is None:
            return
        from .peakfasta import PeakFasta
        name = self.names[i]
        prefix = self._reporter._base_name + (name + "_" if name != "" else "")
        PeakFasta(sequence_graph).write_max_path_sequences(prefix + "sequences.fasta",
                                                           caller.max_path_peaks)

    def peaks(self):
        """{chromosome name: list of Peak} for this rank's chromosomes."""
        return {self.names[i]: self._callers[i].max_path_peaks for i in self.my_chromosomes()}

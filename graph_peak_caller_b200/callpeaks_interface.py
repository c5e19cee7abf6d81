"""The ``args``-driven entry points of the reference
(graph_peak_caller/callpeaks_interface.py) by the same names and with the same
file-name conventions, over the device path:

``run_callpeaks_interface`` (:80-92), ``run_callpeaks2`` (:136-251),
``run_callpeaks_whole_genome`` (:253-338),
``run_callpeaks_whole_genome_from_p_values`` (:340-371) and their helpers
``estimate_read_length`` (:17-24), ``get_confiugration`` (sic, :27-38),
``get_intervals`` (:41-46), ``get_callpeaks`` (:49-54), ``parse_input_file``
(:57-77), ``find_or_create_linear_map`` (:107-116), ``_get_file_names``
(:119-133), plus ``create_ob_graph`` / ``create_linear_map`` of
preprocess_interface.py:51-57 / util.py:6-9, which write the two files every run
looks for.

``args`` is any object with the attributes the reference's argparse namespaces
carry (``getattr`` with a default of None for the optional ones).  Differences,
all outside the data-parallel path:

* graph files: the flat graph of ``graph.py`` written by ``create_ob_graph`` /
  ``Graph.to_file`` under whatever name the reference would use (``chr1.nobg``
  works: it is an npz container), or a vg graph in JSON form (``*.json``).
  offsetbasedgraph's own binary ``.nobg`` needs that package and is refused
  with a message that says so;
* sequence graphs (``<graph file>.sequences``, written by ``create_ob_graph`` from
  the vg JSON) are this package's flat ``SequenceGraph``; with them the runs write
  ``sequences.fasta`` as the reference does;
* ``fragment_length`` missing in ``run_callpeaks2``: estimated as in the reference
  (``shiftestimation.py`` of this package, native host sweeps).
"""
import logging
import os
import sys

import numpy as np

from .callpeaks import CallPeaks, Configuration
from .control.linearmap import LinearMap
from .conversion import json_file_to_obg_numpy_graph, load_graph, \
    vg_json_file_to_interval_collection
from .intervals import UniqueIntervals
from .multiplegraphscallpeaks import MultipleGraphsCallpeaks
from .peakcollection import IntervalCollection
from .reads import ReadBatch
from .peakfasta import PeakFasta
from .reporter import Reporter
from .sequencegraph import SequenceGraph


def _arg(args, name, default=None):
    return getattr(args, name, default)


class _SequenceGraphFiles:
    """Sequence graphs by chromosome index, read when asked for (the reference hands
    MultipleGraphsCallpeaks a generator over the same files, :156-159,277-279)."""

    def __init__(self, file_names):
        self.file_names = list(file_names)

    def get(self, i):
        return SequenceGraph.from_file(self.file_names[i])


def create_ob_graph(args):
    """preprocess_interface.py:51-57: vg graph JSON -> graph file (default
    name: the JSON file's name up to its first dot + ``.nobg``)."""
    graph = json_file_to_obg_numpy_graph(args.vg_json_file_name, 0)
    out_name = _arg(args, "out_file_name")
    if out_name is None:
        out_name = args.vg_json_file_name.split(".")[0] + ".nobg"
    with open(out_name, "wb") as fh:           # a file object: numpy keeps the name as it is
        graph.to_file(fh)
    logging.info("Wrote graph to %s" % out_name)
    # the node sequences next to it, as the reference does (preprocess_interface.py:64-68)
    SequenceGraph.from_vg_json_file(args.vg_json_file_name, graph).to_file(out_name + ".sequences")
    logging.info("Wrote sequence graph to %s" % (out_name + ".sequences"))
    return out_name


def create_linear_map(ob_graph, out_file_name="linear_map.npz"):
    """util.py:6-9"""
    LinearMap.from_graph(ob_graph).to_file(out_file_name)
    logging.info("Created linear map, wrote to file %s" % out_file_name)


def create_linear_map_interface(args):
    """preprocess_interface.py:72-80: ``args.graph`` (a graph or its file name) ->
    ``<graph file up to the first dot>_linear_map.npz`` unless a name is given."""
    out_name = args.out_file_base_name if _arg(args, "out_file_base_name") is not None \
        else args.graph_file_name.split(".")[0] + "_linear_map.npz"
    create_linear_map(load_graph(args.graph), out_name)
    logging.info("Wrote linear map to file %s" % out_name)
    return out_name


def count_unique_reads(chromosomes, graph_file_names, reads_file_names):
    """preprocess_interface.py:41-48: unique reads (by start position) summed over the
    alignment files of all chromosomes; de-duplication runs on the device."""
    reads = (parse_input_file(f, load_graph(g)).intervals
             for f, g in zip(reads_file_names, graph_file_names))
    unique_reads = MultipleGraphsCallpeaks.count_number_of_unique_reads(reads)
    print(unique_reads)
    return unique_reads


def split_vg_json_reads_into_chromosomes(args):
    """preprocess_interface.py:83-145 with its ``args`` (``chromosomes``,
    ``vg_json_reads_file_name``, ``range_files_base_name``); the work is
    ``conversion.split_vg_json_reads_into_chromosomes``."""
    from .conversion import split_vg_json_reads_into_chromosomes as split
    return split(args.chromosomes, args.vg_json_reads_file_name, args.range_files_base_name)


def count_unique_reads_interface(args):
    """preprocess_interface.py:11-18"""
    chromosomes = args.chromosomes.split(",")
    graph_file_names = [args.graphs_location + chrom + ".nobg" for chrom in chromosomes]
    reads_file_names = [args.reads_base_name + chrom + ".json" for chrom in chromosomes]
    return count_unique_reads(chromosomes, graph_file_names, reads_file_names)


def find_or_create_linear_map(graph, linear_map_name):
    if os.path.isfile(linear_map_name):
        logging.info("Found linear map %s. Will be used in peak calling." % linear_map_name)
    else:
        logging.info("No linear map provided, and none found. Will create now.")
        create_linear_map(graph, linear_map_name)
        logging.info("Done creating linear map")


def parse_input_file(input, graph):
    """``*.json`` holds vg alignments, anything else is an interval collection
    (text); a missing file ends the run as in the reference (:73-77)."""
    logging.info("Parsing input file %s" % input)
    try:
        if isinstance(input, (IntervalCollection, ReadBatch)):
            return input
        if input.endswith(".json"):
            return vg_json_file_to_interval_collection(input, graph)
        return IntervalCollection(ReadBatch.from_intervalcollection_file(input))
    except FileNotFoundError as e:
        logging.debug(e)
        logging.critical("Input file %s not found. Aborting. " % input)
        sys.exit(1)


def estimate_read_length(file_name, graph_name):
    """Median length of the reads in ``file_name`` (:17-24)."""
    graph = load_graph(graph_name)
    reads = parse_input_file(file_name, graph).intervals
    return int(np.median(reads.lengths(graph)))


def estimate_fragment_length(graphs, samples, args=None):
    """The reference's estimate when ``-f`` is missing (:201-212): MACS2's peak model
    over the read starts on the most-covered path of every graph
    (``shiftestimation.MultiGraphShiftEstimator``), fold-enrichment bounds from
    ``min_fold_enrichment`` / ``max_fold_enrichment`` (defaults 5 and 50)."""
    from .shiftestimation import MultiGraphShiftEstimator
    min_m, max_m = _arg(args, "min_fold_enrichment"), _arg(args, "max_fold_enrichment")
    min_m = 5 if min_m is None else int(min_m)
    max_m = 50 if max_m is None else int(max_m)
    return int(MultiGraphShiftEstimator.from_files(graphs, samples, min_m, max_m).get_estimates())


def shift_estimation(args):
    """preprocess_interface.py:21-38: the estimate for the chromosomes of a data
    directory (``<ob_graphs_location>/<chrom>.nobg``, ``<sample_reads_base_name><chrom>.json``)."""
    chromosomes = args.chromosomes.split(",")
    graphs = [args.ob_graphs_location + "/" + chrom + ".nobg" for chrom in chromosomes]
    samples = [args.sample_reads_base_name + chrom + ".json" for chrom in chromosomes]
    d = estimate_fragment_length(graphs, samples, args)
    logging.info("Found shift: %d" % d)
    return d


def get_confiugration(args):
    config = Configuration()
    config.has_control = _arg(args, "control") is not None
    if _arg(args, "linear_map") is None:
        config.linear_map_name = args.graph_file_name.split(".")[0] + "_linear_map.npz"
    else:
        config.linear_map_name = args.linear_map
    config.fragment_length = int(args.fragment_length)
    config.read_length = int(args.read_length)
    if _arg(args, "q_value_threshold") is not None:
        # the reference assigns a misspelt attribute here (:38), so its threshold stays
        # at the default; set both so the flag means what it says
        config.q_value_threshold = float(args.q_value_threshold)
        config.q_values_threshold = float(args.q_value_threshold)
    return config


def get_intervals(args):
    iclass = UniqueIntervals
    samples = iclass(parse_input_file(args.sample, args.graph))
    control_name = args.sample if _arg(args, "control") is None else args.control
    controls = iclass(parse_input_file(control_name, args.graph))
    return samples, controls


def get_callpeaks(args):
    config = get_confiugration(args)
    find_or_create_linear_map(args.graph, config.linear_map_name)
    out_name = args.out_name if _arg(args, "out_name") is not None else ""
    return CallPeaks(args.graph, config, Reporter(out_name))


def run_callpeaks_interface(args):
    """One graph: ``args.graph`` (a graph object), ``graph_file_name``,
    ``sample``, ``control``, ``fragment_length``, ``read_length``, ``out_name``,
    ``linear_map``, ``q_value_threshold``.  Returns the caller."""
    caller = get_callpeaks(args)
    inputs, controls = get_intervals(args)
    caller.run(inputs, controls)
    outname = args.out_name if _arg(args, "out_name") is not None else ""
    if _arg(args, "sequence_graph") is not None:
        sequence_graph = args.sequence_graph
        if isinstance(sequence_graph, str):         # a file name: <graph file>.sequences
            sequence_graph = SequenceGraph.from_file(sequence_graph, graph=args.graph)
        PeakFasta(sequence_graph).write_max_path_sequences(outname + "sequences.fasta",
                                                           caller.max_path_peaks)
    else:
        logging.info("Not saving max path sequences, since a sequence graph was not found.")
    return caller


def create_alignment_fasta(args):
    """:94-103: the sequence of every unique alignment of a vg JSON file, one FASTA
    record each (all named ``>1`` as in the reference)."""
    intervals = UniqueIntervals(parse_input_file(args.vg_json_file_name, args.graph))
    sequence_graph = SequenceGraph.from_file(
        args.data_dir + "/" + args.chrom + ".nobg.sequences", graph=args.graph)
    with open(args.out_file_name, "w") as out:
        for interval in intervals:
            out.write(">1\n" + sequence_graph.get_interval_sequence(interval) + "\n")


def _get_file_names(pattern):
    from glob import glob
    if pattern is None:
        return None
    assert isinstance(pattern, list)
    pattern = sorted(pattern)
    if len(pattern) == 1 and "*" in pattern[0]:
        return sorted(glob(pattern[0]))
    return pattern


def _graph_base(graph_file_name):
    """``chr1.nobg`` -> ``chr1`` as the reference (:176); other extensions likewise."""
    if ".nobg" in graph_file_name:
        return graph_file_name.split(".nobg")[0]
    return graph_file_name.rsplit(".", 1)[0]


def _linear_maps_for(graph_file_names, linear_map_names):
    for graph_file_name, linear_map_name in zip(graph_file_names, linear_map_names):
        if not os.path.isfile(linear_map_name):
            logging.warning("Did not find linear map for graph %s. Will create." % graph_file_name)
            create_linear_map(load_graph(graph_file_name), linear_map_name)
        else:
            logging.info("Found linear map %s that will be used." % linear_map_name)
    return list(linear_map_names)


def prepare_callpeaks2(args):
    """Everything ``run_callpeaks2`` does before ``caller.run()``: file lists,
    experiment names, linear maps, configuration.  Returns the
    ``MultipleGraphsCallpeaks`` ready to run."""
    graphs = _get_file_names(args.graph)
    samples = _get_file_names(args.sample)
    controls = _get_file_names(_arg(args, "control"))
    if controls is None:
        controls = samples
    if len(samples) != len(graphs):
        logging.critical("Number of sample files must be equal to number of graph files. "
                         "Found %d graphs and %d sample files." % (len(graphs), len(samples)))
        sys.exit(1)
    logging.info("Sample files: %s" % samples)
    logging.info("Using graphs: %s " % graphs)
    data_dir = os.path.dirname(graphs[0])
    if data_dir == "":
        data_dir = "./"
    if len(graphs) > 1:
        # the reference keeps the separator that follows data_dir (names like "/chr1", :163);
        # it is dropped here so that the names can prefix the Reporter's files
        names = [fn.split(data_dir, 1)[-1].lstrip(os.sep).rsplit(".", 1)[0] for fn in graphs]
    else:
        names = [""]
    linear_map_file_names = _linear_maps_for(
        graphs, [_graph_base(fn) + "_linear_map.npz" for fn in graphs])

    config = Configuration()
    if _arg(args, "keep_duplicates") == "True":
        config.keep_duplicates = True
        logging.info("Keeping duplicates")
    if _arg(args, "q_threshold") is not None:
        config.q_values_threshold = float(args.q_threshold)
        logging.info("Running with q value threshold %.3f" % config.q_values_threshold)
    if _arg(args, "fragment_length") is None:
        logging.info("Fragment length was not specified. Will now predict fragment length.")
        config.fragment_length = estimate_fragment_length(graphs, samples, args)
        logging.info("Estimated fragment length to be %d" % config.fragment_length)
        assert config.fragment_length < 1000, \
            "Suspiciously high fragment length. Probably a bad estimate." \
            " Change -m/-M to try to get more paired peaks."
    else:
        config.fragment_length = int(args.fragment_length)
    if _arg(args, "read_length") is None:
        config.read_length = estimate_read_length(samples[0], graphs[0])
        logging.info("Estimated read length to %s" % config.read_length)
    else:
        config.read_length = int(args.read_length)
    if config.fragment_length < config.read_length:
        logging.critical("Fragment length is smaller than read length. Cannot call peaks.")
        sys.exit(1)
    if _arg(args, "genome_size") is not None:
        genome_size = int(args.genome_size)
        # the reference multiplies by int(args.fragment_length) (:222), which fails after
        # an estimate; the configured length is the same number whenever that one works
        config.global_min = int(args.unique_reads) * config.fragment_length / genome_size
        logging.info("Computed min background signal to be %.3f" % config.global_min)
    else:
        logging.info("Not using min background.")
        config.global_min = None
    out_name = args.out_name if _arg(args, "out_name") is not None else ""
    config.has_control = _arg(args, "control") is not None
    return MultipleGraphsCallpeaks(
        names, graphs, samples, controls, linear_map_file_names, config, Reporter(out_name),
        sequence_retrievers=_SequenceGraphFiles(fn + ".sequences" for fn in graphs),
        stop_after_p_values=_arg(args, "stop_after_p_values") == "True")


def run_callpeaks2(args):
    """``graph``, ``sample``, ``control``: lists of file names (or one pattern
    with ``*``); ``fragment_length``, ``read_length``, ``out_name``,
    ``q_threshold``, ``keep_duplicates``, ``genome_size`` + ``unique_reads``,
    ``stop_after_p_values``."""
    caller = prepare_callpeaks2(args)
    caller.run()
    return caller


def prepare_callpeaks_whole_genome(args):
    config = Configuration()
    if _arg(args, "keep_duplicates") == "True":
        config.keep_duplicates = True
        logging.info("Keeping duplicates")
    chromosomes = args.chromosomes.split(",")
    graph_file_names = [args.data_dir + "/" + chrom + ".nobg" for chrom in chromosomes]
    logging.info("Will use graphs: %s" % graph_file_names)
    linear_map_file_names = _linear_maps_for(
        graph_file_names, [args.data_dir + "/" + chrom + "_linear_map.npz" for chrom in chromosomes])

    def per_chromosome(name):
        if name.endswith(".intervalcollection"):
            return [name.replace("chrom", chrom) for chrom in chromosomes]
        base = name.replace(".json", "_")
        return [base + chrom + ".json" for chrom in chromosomes]

    sample_file_names = per_chromosome(args.sample)
    logging.info("Will use input alignments from %s" % sample_file_names)
    if _arg(args, "control") is not None:
        control_file_names = per_chromosome(args.control)
    else:
        logging.info("Using input alignments as control")
        control_file_names = sample_file_names.copy()
    config.fragment_length = int(args.fragment_length)
    config.read_length = int(args.read_length)
    if config.fragment_length < config.read_length:
        logging.critical("Fragment length is smaller than read length. Cannot call peaks.")
        sys.exit(1)
    genome_size = int(args.genome_size)
    config.global_min = int(args.unique_reads) * int(args.fragment_length) / genome_size
    logging.info("Computed min background signal to be %.3f" % config.global_min)
    out_name = args.out_name if _arg(args, "out_name") is not None else ""
    config.has_control = _arg(args, "control") is not None
    return MultipleGraphsCallpeaks(
        chromosomes, graph_file_names, sample_file_names, control_file_names,
        linear_map_file_names, config, Reporter(out_name),
        sequence_retrievers=_SequenceGraphFiles(
            args.data_dir + "/" + chrom + ".nobg.sequences" for chrom in chromosomes),
        stop_after_p_values=_arg(args, "stop_after_p_values") == "True")


def run_callpeaks_whole_genome(args):
    """``chromosomes`` ("1,2,X"), ``data_dir`` (``<chrom>.nobg`` and
    ``<chrom>_linear_map.npz`` live there), ``sample`` / ``control`` (``x.json``
    -> ``x_<chrom>.json``; ``*.intervalcollection`` names have ``chrom``
    replaced), ``fragment_length``, ``read_length``, ``genome_size``,
    ``unique_reads``, ``out_name``, ``stop_after_p_values``."""
    logging.info("Running run_callpeaks_whole_genome")
    caller = prepare_callpeaks_whole_genome(args)
    caller.run()
    return caller


def run_callpeaks_whole_genome_from_p_values(args):
    """Second half of a whole-genome run for one chromosome: joins the
    ``*pvalues*`` files every chromosome's first half wrote under ``out_name``
    and calls the peaks of ``args.chromosome`` (:340-371)."""
    logging.info("Running whole genome from p-values.")
    chromosome = args.chromosome
    if _arg(args, "variant_maps_path") is not None:
        raise NotImplementedError("variant maps are outside this path")
    graph_file_names = [args.data_dir + chromosome + ".nobg"]      # sic: no separator (:344)
    out_name = args.out_name if _arg(args, "out_name") is not None else ""
    config = Configuration()
    if _arg(args, "q_threshold") is not None:
        config.q_values_threshold = float(args.q_threshold)
        logging.info("Running with q value threshold %.3f" % config.q_values_threshold)
    config.fragment_length = int(args.fragment_length)
    config.read_length = int(args.read_length)
    caller = MultipleGraphsCallpeaks(
        [chromosome], graph_file_names, None, None, None, config, Reporter(out_name),
        sequence_retrievers=_SequenceGraphFiles(
            [args.data_dir + "/" + chromosome + ".nobg.sequences"]))
    caller.create_joined_q_value_mapping()
    caller.run_from_p_values(only_chromosome=chromosome)
    return caller

"""The three args-driven callers that sit behind the summits / sequence row
 in the reference's analysis/analysis_interface.py:
``peaks_to_fasta`` (:570-578), ``get_summits`` (:388-406) and
``concatenate_sequence_files`` (:252-310).  ``args`` is any object carrying the
reference's argument names (command_line_interface.py:204-214,271-284 and the
``get_summits`` entry); the rest of that module (motif enrichment, haplotypes,
plots) is analysis around the peak caller and stays out."""
import logging

from .conversion import load_graph
from .peakcollection import Peak, PeakCollection
from .peakfasta import PeakFasta
from .sequencegraph import SequenceGraph
from .sparsediffs import SparseDiffs


def _arg(args, name, default=None):
    return getattr(args, name, default)


def _sequence_graph(source):
    return SequenceGraph.from_file(source) if isinstance(source, str) else source


def peaks_to_fasta(args):
    """Sequences of the peaks of an ``.intervalcollection`` file, in file order."""
    retriever = _sequence_graph(args.sequence_graph)
    intervals = PeakCollection.from_file(args.intervals_file_name, text_file=True)
    PeakFasta(retriever).save_intervals(args.out_file_name, intervals)


def get_summits(args):
    """Cuts every peak of a FASTA file to ``window_size`` bp either side of its q-value
    summit and writes ``<name up to the first dot>_summits.fasta``.  The q-values stay
    run-length encoded: ``PeakCollection.cut_around_summit`` runs ``gpc_peak_summits``
    over the runs where the reference builds a dense array of the whole graph
    (:390-393).  As in the reference the two files are read with
    ``SparseDiffs.from_sparse_files`` (:390), i.e. their values are taken as
    increments and summed along the track before use; for the track a run itself
    holds, call ``PeakCollection.cut_around_summit(caller.q_values_pileup, n)``."""
    graph = load_graph(args.graph) if isinstance(args.graph, str) else args.graph
    qvalues = SparseDiffs.from_sparse_files(args.q_values_base_name)
    peaks = PeakCollection.from_fasta_file(args.peaks_fasta_file, graph)
    if _arg(args, "window_size") is not None:
        window = int(args.window_size)
    else:
        window = 60
        logging.warning("Using default window size %d when cutting peaks around summits." % window)
    peaks.cut_around_summit(qvalues, n_base_pairs_around=window)
    out_file_name = args.peaks_fasta_file.split(".")[0] + "_summits.fasta"
    sequence_graph = _arg(args, "sequence_graph")
    if sequence_graph is None:                  # the reference's wrapper finds it by the graph's name
        sequence_graph = args.graph + ".sequences" if isinstance(args.graph, str) else None
    peaks.to_fasta_file(out_file_name, _sequence_graph(sequence_graph))
    logging.info("Wrote summits to " + out_file_name)
    return peaks


def concatenate_sequence_files(args):
    """``<chrom>_sequences.fasta`` (or ``_sequences_summits.fasta``, or a pattern with
    ``[chrom]``) of every chromosome into one FASTA file and one ``.intervalcollection``
    file, peaks by descending score (stable: equal scores keep chromosome order) and
    renumbered."""
    chromosomes = args.chromosomes.split(",")
    out_file_name = args.out_file_name
    out_file_name_json = ".".join(out_file_name.split(".")[:-1]) + ".intervalcollection"
    ending = "_sequences_summits.fasta" if _arg(args, "is_summits") == "True" else "_sequences.fasta"
    pattern = _arg(args, "use_input_file_pattern")
    if pattern is not None:
        assert "[chrom]" in pattern, "Use [chrom] to specify where chromosome should be replaced."
    peaks = []
    for chromosome in chromosomes:
        name = pattern.replace("[chrom]", chromosome) if pattern is not None else chromosome + ending
        with open(name) as fasta_file:
            for line in fasta_file:
                if line.startswith(">"):
                    peak = Peak.from_file_line("{" + line.rstrip().split(" {", 1)[1])
                    peak.sequence = None
                    peaks.append(peak)
                else:
                    peaks[-1].sequence = line.rstrip()
    peaks.sort(key=lambda p: -p.score)
    with open(out_file_name, "w") as out_fasta, open(out_file_name_json, "w") as out_json:
        for number, peak in enumerate(peaks):
            out_fasta.write(">peak%d %s\n%s\n" % (number, peak.to_file_line(), peak.sequence))
            out_json.write("%s\n" % peak.to_file_line())
    logging.info("Wrote all peaks in sorted order to %s" % out_file_name)
    logging.info("Wrote all peaks in sorted order to %s" % out_file_name_json)
    return peaks

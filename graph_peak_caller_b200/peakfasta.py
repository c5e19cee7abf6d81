"""``PeakFasta`` (reference graph_peak_caller/peakfasta.py:4-31): peaks as FASTA
records, header ``>peak<i> <the peak's JSON line>``, one line of sequence."""
import logging


class PeakFasta:
    def __init__(self, sequence_retriever):
        self._sequence_retriever = sequence_retriever

    def _write(self, file_name, peaks):
        lookup = self._sequence_retriever.get_interval_sequence
        with open(file_name, "w") as out:
            for number, peak in enumerate(peaks):
                out.write(">peak%d %s\n%s\n" % (number, peak.to_file_line(), lookup(peak)))

    def write_max_path_sequences(self, file_name, max_paths):
        self._write(file_name, max_paths)
        logging.info("Wrote max path sequences to fasta file: %s" % file_name)

    def save_intervals(self, out_fasta_file_name, interval_collection):
        self._write(out_fasta_file_name, interval_collection.intervals)

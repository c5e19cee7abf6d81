"""Read-set wrappers of the reference (graph_peak_caller/intervals.py): they
are what CallPeaks.run receives, and after the reads have been consumed they
expose ``n_reads`` (used for the sample / control ratio, callpeaks.py:65-66)
and ``n_duplicates``.

Here they wrap a flat :class:`ReadBatch` (any iterable of obg-style intervals
is flattened once on the host).  ``UniqueIntervals`` does not filter on the
host: it asks the device for the index list of first-per-start-position reads
(kernel gpc_unique_reads) when the pipeline consumes it.
"""
import os

from .reads import ReadBatch


def _as_batch(intervals):
    if isinstance(intervals, ReadBatch):
        return intervals
    if isinstance(intervals, (str, os.PathLike)):
        # by file name, as the reference's parse_input_file (callpeaks_interface.py:58-70):
        # *.json holds vg alignments, anything else is an .intervalcollection text file
        name = os.fspath(intervals)
        if name.endswith(".json"):
            return ReadBatch.from_vg_json_file(name)
        return ReadBatch.from_intervalcollection_file(name)
    inner = getattr(intervals, "intervals", intervals)     # obg IntervalCollection
    if isinstance(inner, ReadBatch):
        return inner
    return ReadBatch.from_intervals(inner)


class Intervals:
    unique = False

    def __init__(self, intervals):
        self._intervals = intervals
        self._batch = None
        self.n_reads = 0
        self.n_duplicates = 0

    @property
    def batch(self):
        if self._batch is None:
            self._batch = _as_batch(self._intervals)
        return self._batch

    def prefetch(self):
        """Start the host->device copy of the reads on a side stream (useful
        when the batch sits in page-locked memory): CallPeaks uploads the
        control reads this way while the sample pileup is being computed."""
        from .device import prefetch_reads
        prefetch_reads(self.batch)

    def _device_reads(self):
        from .device import device_reads
        return device_reads(self.batch)

    def consume(self):
        """Called by the pipeline: returns (DeviceReads, read index tensor or
        None, number of reads) and sets the counters."""
        dr = self._device_reads()
        self.n_reads = dr.n
        return dr, None, dr.n

    def __iter__(self):
        self.n_reads = 0
        for interval in self.batch:
            self.n_reads += 1
            yield interval


class UniqueIntervals(Intervals):
    unique = True

    def consume(self):
        from . import kernels
        dr = self._device_reads()
        index, n = kernels.unique_reads(dr)
        self.n_reads = n
        self.n_duplicates = dr.n - n
        return dr, index, n

    def __iter__(self):
        # host-side view with the same semantics (first read per start position)
        seen = set()
        self.n_reads = self.n_duplicates = 0
        for interval in self.batch:
            key = interval.hash(ignore_end_pos=True)
            if key in seen:
                self.n_duplicates += 1
                continue
            seen.add(key)
            self.n_reads += 1
            yield interval

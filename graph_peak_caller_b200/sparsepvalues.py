"""p-value / q-value classes of the reference
(graph_peak_caller/sparsepvalues.py) over the kernels gpc_pvalues,
gpc_ptable_local, gpc_qtable_build and gpc_qvalues_apply."""
import logging
from glob import glob

import numpy as np

from . import kernels
from .device import to_device, torch_cuda
from .sparsediffs import SparseDiffs, SparseValues


def _runs_on_device(track):
    """(pos, values, scale) of a SparseDiffs / SparseValues on the device."""
    if isinstance(track, SparseDiffs):
        dr = track.device_runs()
        if dr is not None:
            return dr
        track = track.get_sparse_values()
    pos, val = track.device_arrays()
    return pos, val, 1.0


class PValuesFinder:
    def __init__(self, sample_pileup, control_pileup):
        self.sample = sample_pileup
        self.control = control_pileup

    def get_p_values_pileup(self):
        """reference sparsepvalues.py:16-31"""
        torch = torch_cuda()
        spos, sval, sscale = _runs_on_device(self.sample)
        cpos, cval, cscale = _runs_on_device(self.control)
        if sval.dtype != torch.int32:
            # host-built sample track: poisson.logsf floors k (sparsepvalues.py:21); the
            # epsilon only absorbs the residue of a float cumsum of integer-valued steps
            sval = torch.floor(sval.to(torch.float64) + 1e-9).to(torch.int32)
        cval = cval.to(torch.float64)
        pos, p = kernels.pvalues(spos, sval, sscale, cpos, cval, cscale)
        return SparseValues(pos, p)


class PToQValuesMapper:
    """Distinct p-values (descending) with their q-values.  ``p_values`` /
    ``cum_counts`` follow the reference attributes when built on the host;
    device-built mappers keep (p, q) tables in HBM."""

    def __init__(self, p_values, cum_counts):
        self.p_values = np.asanyarray(p_values)
        self.cum_counts = np.asanyarray(cum_counts)
        self._table = None            # (p desc, q) CUDA tensors

    @classmethod
    def _from_device_table(cls, table_p, table_count, total_bp):
        obj = cls.__new__(cls)
        obj.p_values = obj.cum_counts = None
        obj._table = kernels.qtable_build(table_p, table_count, total_bp)
        return obj

    @classmethod
    def _from_subcounts(cls, p_values, counts):
        """reference sparsepvalues.py:50-59 (host arrays in, device table out)"""
        p = np.asarray(p_values, dtype=np.float64).ravel()
        c = np.asarray(counts).ravel().astype(np.int64)
        keep = p != 0
        return cls._from_device_table(to_device(p[keep]), to_device(c[keep]), int(c.sum()))

    @classmethod
    def from_p_values_pileup(cls, p_values):
        logging.info("Creating mapping from p value dense pileup")
        pos, p = p_values.device_arrays()
        tp, tc = kernels.ptable_local(pos, p, int(p_values.track_size))
        # N = sum of the run lengths = track_size - indices[0] (sparsepvalues.py:70-74); tracks
        # made by the kernels start at 0, a host-built one may not
        first = 0 if p_values._h_idx is None else int(np.asarray(p_values._h_idx)[0])
        return cls._from_device_table(tp, tc, int(p_values.track_size) - first)

    @classmethod
    def from_files(cls, base_file_name):
        """reference sparsepvalues.py:76-93: join the tables of every
        chromosome whose p-values were written under this base name."""
        files = glob(base_file_name + "*pvalues_indexes.npy")
        torch = torch_cuda()
        tps, tcs, total = [], [], 0
        for filename in files:
            chrom = SparseValues.from_sparse_files(filename.replace("_indexes.npy", ""))
            pos, p = chrom.device_arrays()
            tp, tc = kernels.ptable_local(pos, p, int(chrom.track_size))
            tps.append(tp)
            tcs.append(tc)
            total += int(chrom.track_size)
        return cls._from_device_table(torch.cat(tps), torch.cat(tcs), total)

    def device_table(self):
        if self._table is None and self.p_values is not None:
            # built the reference's way from (distinct p descending, cumulative counts)
            counts = np.ediff1d(np.asarray(self.cum_counts, dtype=np.int64),
                                to_begin=self.cum_counts[0])
            p = np.asarray(self.p_values, dtype=np.float64)
            keep = p != 0
            self._table = kernels.qtable_build(to_device(p[keep]), to_device(counts[keep]),
                                               int(self.cum_counts[-1]))
        return self._table

    def get_p_to_q_values(self):
        """dict p -> q with d[0] = 0 (reference sparsepvalues.py:95-103).  The
        dict is what the reference API hands around; it is filled from the
        device table the first time somebody looks inside, and lookups on the
        device use the table itself (attribute ``device_table``)."""
        return PToQDict(self.device_table())


class PToQDict(dict):
    """A real ``dict`` (the reference asserts ``isinstance(..., dict)``) that
    copies the (p, q) table from the device on first use."""

    def __init__(self, device_table):
        super().__init__()
        self.device_table = device_table
        self._pending = device_table is not None

    def _fill(self):
        if self._pending:
            self._pending = False
            p, q = self.device_table
            super().update(zip(p.cpu().numpy().tolist(), q.cpu().numpy().tolist()))
            super().__setitem__(0, 0)

    def __getitem__(self, key):
        self._fill()
        return super().__getitem__(key)

    def get(self, key, default=None):
        self._fill()
        return super().get(key, default)

    def __contains__(self, key):
        self._fill()
        return super().__contains__(key)

    def __iter__(self):
        self._fill()
        return super().__iter__()

    def __len__(self):
        self._fill()
        return super().__len__()

    def keys(self):
        self._fill()
        return super().keys()

    def values(self):
        self._fill()
        return super().values()

    def items(self):
        self._fill()
        return super().items()

    def __setitem__(self, key, value):
        self._fill()
        self.device_table = None          # modified by the caller: the dict is the truth now
        super().__setitem__(key, value)

    def __eq__(self, other):
        self._fill()
        return super().__eq__(other)

    def __repr__(self):
        self._fill()
        return super().__repr__()


class QValuesFinder:
    def __init__(self, p_values_pileup, p_to_q_values):
        assert isinstance(p_to_q_values, dict)
        self.p_values = p_values_pileup
        self.p_to_q_values = p_to_q_values

    def _table(self):
        table = getattr(self.p_to_q_values, "device_table", None)
        if table is None:      # a dict built elsewhere: upload it (descending p, 0 left out)
            items = sorted(((float(k), float(v)) for k, v in self.p_to_q_values.items()
                            if k != 0), reverse=True)
            table = (to_device(np.array([k for k, _ in items], dtype=np.float64)),
                     to_device(np.array([v for _, v in items], dtype=np.float64)))
        return table

    def get_q_values(self):
        pos, p = self.p_values.device_arrays()
        tp, tq = self._table()
        return SparseValues(pos, kernels.qvalues_apply(p, tp, tq))

    def get_q_array_from_p_array(self, p_values):
        assert isinstance(p_values, np.ndarray)
        tp, tq = self._table()
        return kernels.qvalues_apply(to_device(p_values.astype(np.float64)), tp, tq) \
            .cpu().numpy()

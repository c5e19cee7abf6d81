"""Track containers with the reference's surface
(graph_peak_caller/sparsediffs.py): ``SparseValues`` (run-length step function)
and ``SparseDiffs`` (event list).

Both can live in HBM (torch CUDA tensors: uint32 positions stored as int32,
int32 / float64 / uint8 values) or on the host (numpy), and move lazily: the
callpeaks path creates them on the device and only materialises numpy arrays
when someone reads ``.indices`` / ``.values`` or writes the ``.npy`` files of
the Reporter contract.  The numpy-side helpers below (sanitize, dense
conversion, file I/O, ``apply_binary_func`` with an arbitrary Python callable)
exist for API compatibility; the peak-calling pipeline itself never calls
them -- it goes through the kernels (pipeline.py).
"""
import logging

import numpy as np


def _is_tensor(x):
    return type(x).__module__.startswith("torch")


def _to_host_index(t):
    return t.cpu().numpy().view(np.uint32).astype(np.int64)


def _device():
    """True when a CUDA device is visible: hand-made tracks then go through the generic
    track kernels (gpc_diffs_to_runs / gpc_diffs_merge / gpc_diffs_combine, csrc/tracks.cu).
    On a machine without a device (the CPU test suite, file conversion tools) the same
    methods are the reference's numpy statements; the callpeaks path itself has no such
    form and raises without a device."""
    from .device import cuda_available
    return cuda_available()


# numpy callables of apply_binary_func that exist as a kernel operation
_DEVICE_FUNCS = {np.maximum: "maximum", np.minimum: "minimum", np.add: "add",
                 np.subtract: "subtract", np.multiply: "multiply"}


class SparseValues:
    def __init__(self, indices, values, sanitize=False):
        self._h_idx = self._h_val = self._d_idx = self._d_val = None
        if _is_tensor(indices):
            self._d_idx, self._d_val = indices, values
        else:
            self._h_idx = np.asanyarray(indices)
            self._h_val = np.asanyarray(values)
            if sanitize:
                self._sanitize()
        self.track_size = None

    # -- lazy host / device views ---------------------------------------------
    @property
    def indices(self):
        if self._h_idx is None:
            self._h_idx = _to_host_index(self._d_idx)
        return self._h_idx

    @indices.setter
    def indices(self, value):
        self.values  # materialise before the device copy is dropped
        self._h_idx = np.asanyarray(value)
        self._d_idx = self._d_val = None

    @property
    def values(self):
        if self._h_val is None:
            v = self._d_val.cpu().numpy()
            if v.dtype == np.int32:
                v = v.astype(np.float64)      # the reference's pileups are float64
            elif v.dtype == np.uint8:
                v = v.astype(bool)
            self._h_val = v
        return self._h_val

    @values.setter
    def values(self, value):
        self.indices
        self._h_val = np.asanyarray(value)
        self._d_idx = self._d_val = None

    def on_device(self):
        return self._d_idx is not None

    def device_arrays(self, value_dtype=None):
        """(positions int32, values) as CUDA tensors, uploading if needed."""
        from .device import to_device
        if self._d_idx is None:
            self._d_idx = to_device(np.asarray(self._h_idx).astype(np.int64).astype(np.uint32)
                                    .view(np.int32))
            v = np.asarray(self._h_val)
            if v.dtype == bool:
                v = v.astype(np.uint8)
            self._d_val = to_device(v)
        val = self._d_val
        if value_dtype is not None and val.dtype != value_dtype:
            val = val.to(value_dtype)
        return self._d_idx, val

    def __len__(self):
        return int(self._d_idx.numel() if self._h_idx is None else self._h_idx.size)

    # -- reference surface -----------------------------------------------------
    def _sanitize(self):
        idx, val = self.indices, self.values
        keep = np.r_[idx[1:] != idx[:-1], True]
        idx, val = idx[keep], val[keep]
        keep = np.r_[True, val[1:] != val[:-1]]
        self._h_idx, self._h_val = idx[keep], val[keep]
        self._d_idx = self._d_val = None

    @classmethod
    def from_sparse_files(cls, file_base_name):
        indices = np.load(file_base_name + "_indexes.npy")
        values = np.load(file_base_name + "_values.npy")
        obj = cls(indices[:-1], values)
        obj.track_size = indices[-1]
        return obj

    def to_sparse_files(self, file_base_name):
        np.save(file_base_name + "_indexes.npy", np.r_[self.indices, self.track_size])
        np.save(file_base_name + "_values.npy", self.values)
        logging.info("Wrote to %s_indexes/values.npy" % file_base_name)

    def __repr__(self):
        return "SV(%s, %s)" % (self.indices, self.values)

    def __eq__(self, other):
        if self.indices.shape != other.indices.shape or \
                not np.all(self.indices == other.indices):
            return False
        if not np.allclose(self.values, other.values):
            return False
        return self.track_size == other.track_size

    def to_bed_graph(self, filename):
        logging.warning("Not writing to %s", filename)

    to_bed_file = to_bed_graph

    def threshold_copy(self, cutoff):
        """values >= cutoff as a canonical boolean track (kernel gpc_threshold
        when the track is device resident)."""
        if self.on_device() and self._d_val.dtype.is_floating_point:
            from . import kernels
            pos, val = kernels.threshold(self._d_idx, self._d_val, cutoff)
            new = SparseValues(pos, val)
        else:
            new = SparseValues(self.indices, self.values >= cutoff, sanitize=True)
        new.track_size = self.track_size
        return new

    def to_dense_pileup(self, size):
        values = self.values
        is_bool = values.dtype == bool
        if is_bool:
            values = values.astype("int")
        diffs = np.ediff1d(values, to_begin=values[0])
        pileup = np.zeros(size + 1, dtype=values.dtype)
        pileup[self.indices[:diffs.size]] = diffs
        pileup = np.cumsum(pileup[:-1])
        return pileup.astype(bool) if is_bool else pileup

    @classmethod
    def from_dense_pileup(cls, pileup):
        change = np.flatnonzero(pileup[1:] != pileup[:-1]) + 1
        indices = np.r_[0, change]
        obj = cls(indices, pileup[indices])
        obj.track_size = pileup.size
        return obj


class SparseDiffs:
    """Event-list track.  When produced by the kernels it is held as canonical
    runs on the device plus a lazy scale factor (``*=`` / ``/=``); ``_indices``
    and ``_diffs`` then expose the *cleaned* event list (what the reference's
    ``clean()`` leaves), since the raw order of equal indices in the reference
    is unspecified anyway (unstable argsort, sparsediffs.py:174,187)."""

    def __init__(self, indices, diffs, sanitize=False):
        self._runs = None           # SparseValues on device (unscaled values)
        self._scale = 1.0
        self._h_indices = np.asanyarray(indices)
        self._h_diffs = np.asanyarray(diffs)
        if sanitize:
            self._sanitize()

    @classmethod
    def from_device_runs(cls, pos, val):
        obj = cls.__new__(cls)
        obj._runs = SparseValues(pos, val)
        obj._scale = 1.0
        obj._h_indices = obj._h_diffs = None
        return obj

    def _materialise(self):
        if self._h_indices is None:
            idx = self._runs.indices
            val = self._runs.values.astype(np.float64) * self._scale \
                if self._scale != 1.0 else self._runs.values.astype(np.float64)
            self._h_indices = idx
            self._h_diffs = np.ediff1d(val, to_begin=val[0])

    @property
    def _indices(self):
        self._materialise()
        return self._h_indices

    @_indices.setter
    def _indices(self, value):
        self._materialise()
        self._h_indices = np.asanyarray(value)
        self._runs = None

    @property
    def _diffs(self):
        self._materialise()
        return self._h_diffs

    @_diffs.setter
    def _diffs(self, value):
        self._materialise()
        self._h_diffs = np.asanyarray(value)
        self._runs = None

    def device_runs(self):
        """(positions, values, scale) of the device-resident canonical runs,
        or None when the track only exists on the host."""
        if self._runs is None:
            return None
        pos, val = self._runs.device_arrays()
        return pos, val, self._scale

    def to_dense_pileup(self, size):
        return self.get_sparse_values().to_dense_pileup(size)

    def to_sparse_files(self, file_base_name):
        np.save(file_base_name + "_indexes.npy", self._indices)
        np.save(file_base_name + "_values.npy", self._diffs)
        logging.info("Wrote to %s and %s" % (file_base_name + "_indexes.npy",
                                             file_base_name + "_values.npy"))

    @classmethod
    def from_sparse_files(cls, file_base_name):
        indices = np.load(file_base_name + "_indexes.npy")
        values = np.load(file_base_name + "_values.npy")
        obj = cls(indices[:-1], values)
        obj.track_size = indices[-1] + 10
        return obj.get_sparse_values()

    def to_bed_graph(self, filename):
        logging.warning("Not writing to %s", filename)

    def __repr__(self):
        return "SD(%s, %s)" % (self._indices, self._diffs)

    def clean(self):
        sv = self.get_sparse_values()
        self._indices = sv.indices
        self._diffs = np.ediff1d(sv.values, to_begin=sv.values[0])

    def clip_min(self, min_value):
        idx = self._indices
        if _device() and idx.size and (idx.size == 1 or bool(np.all(idx[1:] > idx[:-1]))):
            # maximum with the constant track that starts where this one does
            from . import kernels
            pos, val = kernels.diffs_combine(idx, self._diffs, idx[:1], [float(min_value)],
                                             "maximum", drop_leading_zero=True)
            self._runs = SparseValues(pos, val)
            self._scale = 1.0
            self._h_indices = self._h_diffs = None
            return
        values = np.maximum(np.cumsum(self._diffs), min_value)
        diffs = np.empty_like(values)
        diffs[0] = values[0]
        diffs[1:] = np.diff(values)
        self._diffs = diffs
        self._sanitize()

    def get_sparse_values(self):
        if self._runs is not None and self._h_indices is None and self._scale == 1.0:
            return self._runs
        if _device():
            from . import kernels
            return SparseValues(*kernels.diffs_to_runs(self._indices, self._diffs))
        order = np.argsort(self._indices, kind="mergesort")
        return SparseValues(self._indices[order], np.cumsum(self._diffs[order]),
                            sanitize=True)

    def _merged(self, other):
        all_idx = np.r_[self._indices, other._indices]
        order = np.argsort(all_idx, kind="mergesort")
        merged = all_idx[order]
        last = np.ediff1d(merged, to_end=1) != 0
        mine = order < self._indices.size
        d = np.zeros((2, all_idx.size))
        d[0][mine] = self._diffs
        d[1][~mine] = other._diffs
        return merged[last], np.cumsum(d, axis=1), last

    def maximum(self, other):
        if _device():
            from . import kernels
            return SparseDiffs.from_device_runs(*kernels.diffs_combine(
                self._indices, self._diffs, other._indices, other._diffs, "maximum",
                drop_leading_zero=True))
        idx, vals, last = self._merged(other)
        top = np.max(vals, axis=0)[last]
        return SparseDiffs(idx, np.ediff1d(top, to_begin=top[0]), True)

    def apply_binary_func(self, func, other, return_values=False):
        if _device():
            from . import kernels
            if return_values and func in _DEVICE_FUNCS:
                return SparseValues(*kernels.diffs_combine(
                    self._indices, self._diffs, other._indices, other._diffs,
                    _DEVICE_FUNCS[func]))
            # an arbitrary Python callable: the merge and both running sums are kernels,
            # the callable gets the two value arrays as the reference hands them over
            pos, a, b = kernels.diffs_merge(self._indices, self._diffs, other._indices,
                                            other._diffs)
            idx = _to_host_index(pos)
            ret = func(a.cpu().numpy(), b.cpu().numpy())
        else:
            idx, vals, last = self._merged(other)
            ret = func(vals[0], vals[1])[last]
        if return_values:
            return SparseValues(idx, ret, sanitize=True)
        return SparseDiffs(idx, np.ediff1d(ret, to_begin=ret[0]))

    def _sanitize(self):
        keep = self._diffs != 0
        self._h_indices = self._indices[keep]
        self._h_diffs = self._diffs[keep]
        self._runs = None

    def __eq__(self, other):
        return np.all(self._indices == other._indices) and \
            np.all(self._diffs == other._diffs)

    @classmethod
    def from_starts_and_ends(cls, starts_ends):
        indices = starts_ends.ravel(order="C")
        order = np.argsort(indices)
        diffs = np.ones(indices.size)
        diffs[order >= starts_ends.shape[1]] *= -1
        return cls(indices[order], diffs)

    @classmethod
    def from_pileup(cls, pileup, node_indexes):
        indices = np.r_[pileup.starts, pileup.ends, node_indexes]
        diffs = np.r_[np.ones(len(pileup.starts)), -1 * np.ones(len(pileup.ends)),
                      pileup.node_starts]
        order = np.argsort(indices)
        return cls(indices[order], diffs[order])

    @classmethod
    def from_dense_pileup(cls, pileup):
        diffs = np.ediff1d(pileup, to_begin=1)
        changes = np.flatnonzero(diffs)
        diffs = diffs[changes]
        diffs[0] = pileup[0]
        return cls(changes, diffs)

    # In-place scaling (reference sparsediffs.py:200-206).  A kernel-made track keeps its
    # runs on the device with a lazy scale; once somebody has looked at ``_indices`` /
    # ``_diffs`` a host copy exists as well, and BOTH must follow (device_runs() is what
    # the p-value stage reads first).
    def __itruediv__(self, scalar):
        if self._runs is not None:
            self._scale /= scalar
        if self._h_indices is not None:
            self._h_diffs = self._h_diffs / scalar
        return self

    def __imul__(self, scalar):
        if self._runs is not None:
            self._scale *= scalar
        if self._h_indices is not None:
            self._h_diffs = self._h_diffs * scalar
        return self

"""Threshold consumers -- reference graph_peak_caller/postprocess/."""
from .holecleaner import HolesCleaner
from .maxpaths import SparseMaxPaths

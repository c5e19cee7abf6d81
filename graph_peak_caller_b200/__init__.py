"""B200-native callpeaks core behind graph_peak_caller's Python API."""

"""TEST INFRASTRUCTURE ONLY."""


def _na(*a, **k):
    raise NotImplementedError("pyvg is not available; vg JSON ingest is out of scope")


vg_json_file_to_interval_collection = _na
vg_json_file_to_intervals = _na
json_file_to_obg_numpy_graph = _na

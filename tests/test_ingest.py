"""CPU: the native IntervalCollection parser (gpc_parse_intervals_host, SURVEY.md
§8f-1) against the json module on files written by Interval.to_file_line /
Peak.to_file_line."""
import json
import time

import numpy as np
import pytest

from graph_peak_caller_b200.peakcollection import IntervalCollection, Peak
from graph_peak_caller_b200.reads import ReadBatch
from graph_peak_caller_b200.synthetic import make_graph, make_reads


def same_batch(a, b):
    for f in ("start_node", "start_off", "end_node", "end_off", "path_ptr", "path"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f


def test_reads_file_equals_json_parse(tmp_path):
    g = make_graph(20000, 0.05, seed=3)
    reads = make_reads(g, 3000, 25, 90, seed=4)
    name = str(tmp_path / "reads.intervalcollection")
    IntervalCollection(list(reads)).to_file(name, text_file=True)
    got = ReadBatch.from_intervalcollection_file(name)
    same_batch(got, reads)
    via_json = ReadBatch.from_intervals(IntervalCollection.from_file(name, text_file=True))
    same_batch(got, via_json)
    for threads in (1, 3):
        same_batch(ReadBatch.from_intervalcollection_file(name, n_threads=threads), reads)


def test_peak_lines_blank_lines_key_order(tmp_path):
    g = make_graph(2000, 0.05, seed=5)
    peaks = [Peak.from_flat(3, 2, [1, 2, 4], g, 4.5, (True, False), 17),
             Peak.from_flat(0, 7, [9], g, 1.25, (False, False), 7)]
    name = str(tmp_path / "peaks.intervalcollection")
    with open(name, "w") as fh:
        fh.write(peaks[0].to_file_line() + "\n\n")
        # another key order, a negative strand read, odd spacing, a string holding brackets
        fh.write('  {"region_paths":[-7,-6] ,"note":"a [tricky] \\"string\\" }","end":4,'
                 '"nested":{"x":[1,2,{"y":"]"}]},"start" : 2}\r\n')
        fh.write(peaks[1].to_file_line())                      # no newline at the end
    got = ReadBatch.from_intervalcollection_file(name)
    assert got.start_node.tolist() == [1, -7, 9] and got.end_node.tolist() == [4, -6, 9]
    assert got.start_off.tolist() == [3, 2, 0] and got.end_off.tolist() == [2, 4, 7]
    assert got.path_ptr.tolist() == [0, 3, 5, 6] and got.path.tolist() == [1, 2, 4, -7, -6, 9]


def test_empty_file(tmp_path):
    name = str(tmp_path / "none.intervalcollection")
    open(name, "w").close()
    got = ReadBatch.from_intervalcollection_file(name)
    assert len(got) == 0 and got.path_ptr.tolist() == [0]


@pytest.mark.parametrize("line", ['{"start": 1, "end": 2}',
                                  '{"start": 1, "end": 2, "region_paths": []}',
                                  '{"start": 1.5, "end": 2, "region_paths": [1]}',
                                  '["start", 1]',
                                  '{"start": 1, "end": 2, "region_paths": [1, 2'])
def test_malformed_lines_are_rejected(tmp_path, line):
    name = str(tmp_path / "bad.intervalcollection")
    with open(name, "w") as fh:
        fh.write('{"start": 0, "end": 5, "region_paths": [3]}\n' + line + "\n")
    with pytest.raises(Exception, match="malformed line"):
        ReadBatch.from_intervalcollection_file(name)


def test_native_parser_is_much_faster_than_json(tmp_path):
    g = make_graph(200000, 0.02, seed=6)
    reads = make_reads(g, 60000, 36, 200, seed=7)
    name = str(tmp_path / "many.intervalcollection")
    with open(name, "w") as fh:
        for i in range(len(reads)):
            lo, hi = reads.path_ptr[i], reads.path_ptr[i + 1]
            fh.write(json.dumps({"start": int(reads.start_off[i]), "end": int(reads.end_off[i]),
                                 "region_paths": reads.path[lo:hi].tolist(), "direction": 1}) + "\n")
    t0 = time.perf_counter()
    got = ReadBatch.from_intervalcollection_file(name)
    t_native = time.perf_counter() - t0
    t0 = time.perf_counter()
    with open(name) as fh:
        parsed = [json.loads(l) for l in fh]
    t_json = time.perf_counter() - t0
    same_batch(got, reads)
    assert len(parsed) == len(reads)
    assert t_native * 3 < t_json, (t_native, t_json)      # json.loads alone, no Interval objects


def test_intervals_accept_a_file_name(tmp_path):
    from graph_peak_caller_b200.intervals import Intervals, UniqueIntervals
    g = make_graph(5000, 0.05, seed=8)
    reads = make_reads(g, 200, 20, 80, seed=9)
    name = str(tmp_path / "reads.intervalcollection")
    IntervalCollection(list(reads)).to_file(name, text_file=True)
    same_batch(Intervals(name).batch, reads)
    assert sum(1 for _ in UniqueIntervals(name)) == sum(1 for _ in UniqueIntervals(reads))

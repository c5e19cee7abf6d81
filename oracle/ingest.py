"""ORACLE (test infrastructure, CPU, json module) -- vg JSON ingest.

Restates what the reference gets from the third-party package ``pyvg`` (listed
in the reference's setup.py:16 without a version pin; NOT part of
/root/reference and not installed here) on its way into callpeaks:

* ``pyvg.conversion.vg_json_file_to_interval_collection`` -- called at
  callpeaks_interface.py:24,64, multiplegraphscallpeaks.py:97-98,
  preprocess_interface.py:44;
* ``pyvg.conversion.json_file_to_obg_numpy_graph`` -- preprocess_interface.py:53.

Parity status: UNPINNED against pyvg itself (absent).  Anchors that do exist in
the reference tree and are checked by tests/test_ingest_vg.py:

* tests/examples.py:5-31 of the reference builds ``Position(node_id, offset,
  is_reverse)``, ``Edit(to_length, from_length, sequence)``,
  ``Mapping(position, edits)`` from JSON and expects ``Interval(10, 15, [0])``
  for one mapping at offset 10 whose edits have from_length 4 and 1: start =
  first offset, end = last offset + sum(from_length);
* tests/test_vg.py:27-29: ``is_reverse`` is taken from the position;
* tests/test_command_line_interface.py:141-149: the fixture
  tests/node_range_test_data/vg_alignments.json splits into 3 alignments for
  each of the five node ranges next to it;
* preprocess_interface.py:109: node ids may be quoted strings (protobuf JSON);
* tests/vg_test_graph.json with the graph the reference's CLI test builds from
  it (test_command_line_interface.py:18-27: blocks 7, 4, 7, 4; edges 1->2, 1->3,
  2->4, 3->4).

Only tests/ may import this module.
"""
import json
import re

import numpy as np


def alignment_to_interval(obj):
    """One parsed Alignment -> (start, end, region_paths) or None.

    pyvg: Alignment.from_json -> Path.from_json -> Mapping.from_json /
    Position.from_json / Edit.from_json, then Path.to_obg_with_reversals:
    nodes = [-node_id if is_reverse else node_id], start = first mapping's
    offset (default 0), end = last mapping's offset + its edits' from_length
    (default 0 each).  No path / no mappings -> no interval."""
    mappings = obj.get("path", {}).get("mapping", [])
    if not mappings:
        return None
    nodes, offsets = [], []
    for m in mappings:
        pos = m["position"]
        node = int(pos["node_id"])
        nodes.append(-node if pos.get("is_reverse", False) else node)
        offsets.append(int(pos.get("offset", 0)))
    end = offsets[-1] + sum(int(e.get("from_length", 0)) for e in mappings[-1].get("edit", []))
    return offsets[0], end, nodes


def vg_json_lines_to_reads(lines):
    """Flat read arrays (the ReadBatch fields) of an iterable of JSON lines."""
    sn, so, en, eo, ptr, path = [], [], [], [], [0], []
    skipped = 0
    for line in lines:
        if not line.strip():
            continue
        iv = alignment_to_interval(json.loads(line))
        if iv is None:
            skipped += 1
            continue
        start, end, nodes = iv
        sn.append(nodes[0])
        so.append(start)
        en.append(nodes[-1])
        eo.append(end)
        path.extend(nodes)
        ptr.append(len(path))
    return dict(start_node=np.array(sn, dtype=np.int32), start_off=np.array(so, dtype=np.int32),
                end_node=np.array(en, dtype=np.int32), end_off=np.array(eo, dtype=np.int32),
                path_ptr=np.array(ptr, dtype=np.int64), path=np.array(path, dtype=np.int32),
                n_without_path=skipped)


def vg_graph_lines_to_lists(lines):
    """(node ids, node lengths, edge from, edge to) of a vg graph in JSON lines,
    as pyvg.conversion.json_file_to_obg_numpy_graph reads them: node length =
    len(sequence); an edge leaves -from when from_start and enters -to when
    to_end."""
    ids, lens, src, dst = [], [], [], []
    for line in lines:
        if not line.strip():
            continue
        obj = json.loads(line)
        for node in obj.get("node", []):
            ids.append(int(node["id"]))
            lens.append(len(node.get("sequence", "")))
        for edge in obj.get("edge", []):
            a, b = int(edge["from"]), int(edge["to"])
            src.append(-a if edge.get("from_start", False) else a)
            dst.append(-b if edge.get("to_end", False) else b)
    return ids, lens, src, dst


# ---- writers of synthetic files (test inputs) -----------------------------------

def read_to_alignment(node_len, min_node, start_off, end_off, nodes, quoted=False, name=None):
    """A vg Alignment dict for a read: one mapping per node, the edits cover
    the part of the node the read lies on (for a reverse-strand node the
    offset counts on that strand, as in vg and in obg alike)."""
    q = (lambda v: str(v)) if quoted else (lambda v: v)
    mappings = []
    for k, node in enumerate(nodes):
        size = int(node_len[abs(node) - min_node])
        lo = start_off if k == 0 else 0
        hi = end_off if k == len(nodes) - 1 else size
        pos = {"node_id": q(abs(node))}
        if lo:
            pos["offset"] = q(lo)
        if node < 0:
            pos["is_reverse"] = True
        mappings.append({"position": pos, "rank": q(k + 1),
                         "edit": [{"from_length": hi - lo, "to_length": hi - lo}]})
    out = {"sequence": "ACGT", "path": {"mapping": mappings}, "mapping_quality": 60,
           "refpos": [{"name": "1", "offset": q(12345), "is_reverse": True}],
           "identity": 1.0}
    if name is not None:
        out["name"] = name
    return out


def graph_to_vg_json(node_len, min_node, src, dst, chunks=1, with_paths=True):
    """JSON lines of a vg graph with the given 0-based edges, split into
    ``chunks`` lines the way `vg view -Vj` of a chunked file looks."""
    n = len(node_len)
    lines = []
    bounds = np.linspace(0, n, chunks + 1).astype(int)
    src, dst = np.asarray(src), np.asarray(dst)
    for c in range(chunks):
        lo, hi = bounds[c], bounds[c + 1]
        obj = {"node": [{"id": int(i + min_node), "sequence": "A" * int(node_len[i])}
                        for i in range(lo, hi)]}
        mine = (src >= lo) & (src < hi)
        obj["edge"] = [{"from": int(a + min_node), "to": int(b + min_node)}
                       for a, b in zip(src[mine], dst[mine])]
        if with_paths:
            obj["path"] = [{"name": "ref", "mapping": [
                {"position": {"node_id": int(i + min_node)}, "rank": int(i + 1),
                 "edit": [{"from_length": 1, "to_length": 1}]} for i in range(lo, min(hi, lo + 3))]}]
        lines.append(json.dumps(obj))
    return lines


_NODE_ID = re.compile(rb"node_id\":\s?\"?([0-9]+)\"?")


def split_by_chromosome(data, limits):
    """split_vg_json_reads_into_chromosomes (preprocess_interface.py:83-145) on the bytes
    of an alignment file: [bytes per range] + [bytes of "unmapped"].  ``limits``:
    [(first node, last node)] per chromosome, in order."""
    outs = [[] for _ in range(len(limits) + 1)]
    for line in _split_newlines(data):
        found = _NODE_ID.search(line)
        where = len(limits)
        if found is not None:
            node = int(found.group(1))
            for i, (lo, hi) in enumerate(limits):
                if lo <= node <= hi:
                    where = i
                    break
        outs[where].append(line)
    return [b"".join(o) for o in outs]


def _split_newlines(data):
    """Lines end at \\n only (a file object in binary mode)."""
    parts = data.split(b"\n")
    lines = [p + b"\n" for p in parts[:-1]]
    if parts[-1]:
        lines.append(parts[-1])
    return lines

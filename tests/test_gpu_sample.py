"""GPU parity: duplicate filter and sample / direct pileups through the C ABI
against the oracle (bit-exact integer tracks)."""
import numpy as np
import pytest

from golden_util import GoldenCase, case_names
from graph_peak_caller_b200.synthetic import make_graph, make_reads
from oracle import pipeline, sample as osample

pytestmark = pytest.mark.gpu


def gpu_sample(graph, reads, extension, unique, mode=None, headers=True, pinned=False):
    """mode None: gpc_sample_events + gpc_events_to_runs (the caller holds the event list);
    1 / 2: gpc_sample_pileup through the difference array / the event list.  headers False:
    the walk reads the six SoA arrays instead of the 16-byte read records."""
    from graph_peak_caller_b200 import kernels
    from graph_peak_caller_b200.device import DeviceGraph, DeviceReads
    if pinned:           # page-locked batch: goes up in a compact wire form when it can
        reads = reads.pin_memory(delta=(pinned != "plain"))
        assert reads._compact is not None or len(reads) == 0
        assert pinned != "plain" or reads._compact_delta is None
    dg, dr = DeviceGraph(graph), DeviceReads(reads)
    if not headers:
        dr.c.hdr = None
    index, n = None, len(reads)
    if unique:
        index, n = kernels.unique_reads(dr)
    if mode is None:
        events, touched = kernels.sample_events(dg, dr, index, n, extension)
        (fi, fv), (di, dv) = kernels.events_to_runs(events, graph.size_bp)
        n_events = events.numel()
    else:
        (fi, fv), (di, dv), touched, n_events = kernels.sample_pileup(dg, dr, index, n, extension,
                                                                      mode=mode)
    cpu = lambda t: t.cpu().numpy()
    return dict(index=None if index is None else cpu(index), n=n,
                sample=(cpu(fi).view(np.uint32).astype(np.int64), cpu(fv).astype(np.float64)),
                direct=(cpu(di).view(np.uint32).astype(np.int64), cpu(dv).astype(np.float64)),
                touched=np.flatnonzero(cpu(touched)) + graph.min_node,
                n_events=n_events)


def check(graph, reads, extension, unique):
    r = reads
    if unique:
        keep = osample.unique_read_mask(reads)
        r = pipeline.unique_reads(reads)
    ref = osample.fragment_pileup(graph, r, extension)
    first = None
    # every route to the two pileups: difference array, event list, the two-call form
    # (with and without the per-read records)
    for mode, headers, pinned in ((1, True, False), (2, True, False), (None, True, False),
                                  (None, False, False), (1, False, False), (1, True, True),
                                  (2, True, True), (1, True, "plain")):
        got = gpu_sample(graph, reads, extension, unique, mode, headers, pinned)
        what = "mode %s headers %s pinned %s " % (mode, headers, pinned)
        if unique:
            assert np.array_equal(got["index"], np.flatnonzero(keep)), what
        for key in ("sample", "direct"):
            assert np.array_equal(got[key][0], ref[key][0]), what + key
            assert np.array_equal(got[key][1], ref[key][1]), what + key
        assert np.array_equal(got["touched"], ref["touched"]), what
        assert first is None or got["n_events"] == first, what + "steps deposited"
        first = got["n_events"]
    return got


@pytest.mark.parametrize("name", case_names())
def test_golden_cases(name):
    case = GoldenCase(name)
    m = case.meta
    got = check(case.graph, case.sample, m["fragment_length"] - m["read_length"], m["unique"])
    # and straight against what the reference produced (unscaled tracks only)
    if m["n_sample"] <= m["n_control"]:
        gi, gv = case.track("sample")
        assert np.array_equal(got["sample"][0], gi) and np.array_equal(got["sample"][1], gv)
    di, dv = case.track("direct")
    assert np.array_equal(got["direct"][0], di) and np.array_equal(got["direct"][1], dv)
    assert np.array_equal(got["touched"], case.touched)


@pytest.mark.parametrize("trial", range(12))
def test_random_graphs(trial):
    rng = np.random.default_rng(77 + trial)
    g = make_graph(int(rng.integers(500, 60000)),
                   site_rate=float(rng.choice([0.01, 0.03, 0.1, 0.25])), seed=trial,
                   min_node=int(rng.choice([1, 1000])),
                   mnp_frac=float(rng.choice([0, 0.3])),
                   indel_mean=float(rng.choice([2, 5, 30])))
    R = int(rng.integers(5, 40))
    ext = int(rng.integers(1, 300))
    reads = make_reads(g, int(rng.integers(1, 6000)), R, R + ext, seed=trial + 500,
                       peak_frac=0.5, peak_spacing=3000, dup_frac=0.1)
    check(g, reads, ext, unique=bool(trial % 2))


def test_invalid_interval_raises():
    from graph_peak_caller_b200.custom_exceptions import InvalidPileupInterval
    from graph_peak_caller_b200.reads import ReadBatch
    g = make_graph(1000, 0.02, seed=1)
    bad = ReadBatch([g.n_nodes + 5], [0], [g.n_nodes + 5], [3], [0, 1], [g.n_nodes + 5])
    with pytest.raises(InvalidPileupInterval):
        gpu_sample(g, bad, 10, False)


def test_extension_must_be_positive():
    g = make_graph(1000, 0.02, seed=1)
    reads = make_reads(g, 10, 10, 20, seed=3)
    with pytest.raises(Exception):
        gpu_sample(g, reads, 0, False)


def test_medium_size_properties():
    """1 Mbp / 100k reads: too slow for the Python oracle, so check the
    size-independent invariants: canonical form, total coverage == sum over
    reads of covered bases is at least reads*min(len), final value 0."""
    g = make_graph(1_000_000, 0.02, seed=11)
    reads = make_reads(g, 100_000, 36, 200, seed=12)
    got = gpu_sample(g, reads, 164, True)
    for key in ("sample", "direct"):
        idx, val = got[key]
        assert idx[0] == 0 and np.all(np.diff(idx) > 0)
        assert np.all(val[1:] != val[:-1])
        assert val.min() >= 0 and val[-1] == 0
        assert idx[-1] <= g.size_bp
    # direct pileup mass == total read length of unique reads
    idx, val = got["direct"]
    mass = int(np.sum(np.diff(np.r_[idx, g.size_bp]) * val))
    r = pipeline.unique_reads(reads)
    nl = g.node_len[np.abs(r.path) - g.min_node]
    per_read = np.add.reduceat(nl, r.path_ptr[:-1])
    first = np.abs(r.path[r.path_ptr[:-1]]) - g.min_node
    last = np.abs(r.path[r.path_ptr[1:] - 1]) - g.min_node
    fwd = r.start_node > 0
    # forward: start offset in first node, end offset in last; reverse: mirrored
    length = per_read - r.start_off - (g.node_len[np.where(fwd, last, first)] - r.end_off) \
        if False else None
    s_node = np.abs(r.start_node) - g.min_node
    e_node = np.abs(r.end_node) - g.min_node
    length = per_read - r.start_off - (g.node_len[e_node] - r.end_off)
    assert mass == int(length.sum())


def test_page_locked_batch_with_32_bit_path_offsets():
    """A batch that sits in page-locked memory is copied asynchronously; when it
    carries 32-bit path offsets they travel as they are (gpc_reads_t.path_ptr32)."""
    import torch
    from graph_peak_caller_b200 import pipeline as gp
    from graph_peak_caller_b200.device import DeviceGraph, DeviceReads, device_reads, prefetch_reads
    from graph_peak_caller_b200.reads import ReadBatch
    g = make_graph(30000, 0.05, seed=77)
    r = make_reads(g, 4000, 20, 90, seed=78)
    dg = DeviceGraph(g)
    want = gp.sample_pileup(dg, DeviceReads(r), None, len(r), 70)
    pinned = [torch.from_numpy(a).pin_memory() for a in (
        r.start_node, r.start_off, r.end_node, r.end_off, r.path_ptr.astype(np.int32), r.path)]
    for use_prefetch in (False, True):
        rb = ReadBatch.__new__(ReadBatch)
        (rb.start_node, rb.start_off, rb.end_node, rb.end_off, rb.path_ptr, rb.path) = \
            [t.numpy() for t in pinned]
        rb._device = None
        rb._pinned = pinned
        if use_prefetch:
            prefetch_reads(rb)
        dr = device_reads(rb)
        assert dr.path_ptr.element_size() == 4
        got = gp.sample_pileup(dg, dr, None, len(r), 70)
        for a, b in zip(want[:2], got[:2]):
            assert bool((a[0] == b[0]).all()) and bool((a[1] == b[1]).all())
        assert bool((want[2] == got[2]).all())


@pytest.mark.parametrize("n_reads", [1, 300, 70_000, 1_300_000])
def test_delta_coded_wire_form_decodes_to_the_same_reads(n_reads):
    """The three uploads of one page-locked batch -- six arrays, compact, compact with
    delta-coded paths -- leave the same arrays in HBM (start node / offset, per-read
    records, path)."""
    import torch
    from graph_peak_caller_b200.device import DeviceReads
    g = make_graph(200_000, 0.05, seed=91)
    r = make_reads(g, n_reads, 30, 120, seed=92)
    delta, plain = r.pin_memory(), r.pin_memory(delta=False)
    assert delta.wire_form() == "compact+delta" and plain.wire_form() == "compact"
    assert delta.upload_bytes() < plain.upload_bytes() < r.nbytes()
    a, b, c = DeviceReads(delta), DeviceReads(plain), DeviceReads(r)
    torch.cuda.synchronize()
    for x, y in ((a, b), (a, c)):
        assert bool((x.start_node == y.start_node).all()) and bool((x.start_off == y.start_off).all())
        assert bool((x.path == y.path).all()) and bool((x.hdr == y.hdr).all())
    assert np.array_equal(a.path.cpu().numpy(), r.path)


def _fan_graph(rng, n_layers, width):
    """Layers of `width` short parallel nodes with random forward edges between
    neighbouring layers and some skipping a layer: more than four nodes pending
    at once, and nodes reached again while others are pending."""
    from graph_peak_caller_b200.graph import Graph
    node_len, layers, src, dst = [8], [[0]], [], []
    for layer in range(n_layers):
        w = int(rng.integers(2, width + 1)) if layer % 2 == 0 else 1
        ids = list(range(len(node_len), len(node_len) + w))
        node_len += [int(x) for x in rng.integers(1, 4, size=w)]
        for v in ids:                                   # every node gets a predecessor
            src.append(int(rng.choice(layers[-1]))); dst.append(v)
        for u in layers[-1]:                            # ... and a successor
            src.append(u); dst.append(int(rng.choice(ids)))
        for u in layers[-1]:
            for v in ids:
                if rng.random() < 0.5:
                    src.append(u); dst.append(v)
        if len(layers) > 1:
            for u in layers[-2]:
                if rng.random() < 0.3:
                    src.append(u); dst.append(int(rng.choice(ids)))
        layers.append(ids)
    pairs = sorted(set(zip(src, dst)))
    return Graph.from_edges(node_len, [a for a, _ in pairs], [b for _, b in pairs], 1)


@pytest.mark.parametrize("trial", range(6))
def test_wide_fan_out_frontier(trial):
    """Extension frontiers of up to ten nodes (the register part holds four, the
    rest spills), both strands, against the oracle's node sweep."""
    from graph_peak_caller_b200.reads import ReadBatch
    rng = np.random.default_rng(900 + trial)
    g = _fan_graph(rng, n_layers=14, width=10)
    n = g.n_nodes
    fan = np.flatnonzero(np.diff(g.succ_ptr) > 4)
    assert fan.size > 0
    # one-node reads on every node, both strands: every fan-out is crossed by extensions
    nodes = np.arange(n)
    length = g.node_len[nodes]
    start_off = np.zeros(n, dtype=np.int64)
    fwd = ReadBatch(nodes + 1, start_off, nodes + 1, length, np.arange(n + 1), nodes + 1)
    rev = ReadBatch(-(nodes + 1), start_off, -(nodes + 1), length, np.arange(n + 1), -(nodes + 1))
    for reads in (fwd, rev):
        for extension in (3, 11, 40):
            check(g, reads, extension, False)


def test_frontier_outgrows_its_local_spill_list(monkeypatch):
    """Layers of 120 parallel nodes: more than 4 + 32 nodes pending at once.  The spill list
    then moves into a chunk of the launch's pool (no GPC_ERR_FRONTIER_OVERFLOW any more);
    without a pool the call still fails loudly."""
    from graph_peak_caller_b200 import _lib, kernels
    from graph_peak_caller_b200.device import DeviceGraph, DeviceReads
    from graph_peak_caller_b200.reads import ReadBatch
    rng = np.random.default_rng(4242)
    g = _fan_graph(rng, n_layers=6, width=120)
    assert int(np.diff(g.succ_ptr).max()) > 40
    n = g.n_nodes
    nodes = np.arange(n)
    length = g.node_len[nodes]
    zero = np.zeros(n, dtype=np.int64)
    fwd = ReadBatch(nodes + 1, zero, nodes + 1, length, np.arange(n + 1), nodes + 1)
    rev = ReadBatch(-(nodes + 1), zero, -(nodes + 1), length, np.arange(n + 1), -(nodes + 1))
    for reads in (fwd, rev):
        check(g, reads, 9, False)
    monkeypatch.setenv("GPC_SPILL_CHUNKS", "0")
    with pytest.raises(_lib.GpcError, match="frontier"):
        kernels.sample_pileup(DeviceGraph(g), DeviceReads(fwd), None, n, 9, mode=1)

"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's fragment-length
estimation, in plain Python loops over lists:

* ``peak_model``: PeakModel.build (shiftestimation/shiftestimation.py:80-178) with
  ``strand_peaks`` (:262-336), ``pair_centers`` (:231-260), ``add_to_line`` (:180-218),
  ``smooth_flat`` (:339-391 with window="flat");
* ``most_covered_path``: HaploTyper.build + get_maximum_interval_through_graph
  (haplotyping.py:18-58);
* ``start_positions``: LinearFilter.find_start_positions (linear_filter.py:19-46).

Pinned against the live reference in tests/test_shiftestimation.py (PeakModel,
HaploTyper, LinearFilter run unmodified) and by tests/golden/shift/shift_model.npz made from
it (oracle/make_golden_shift.py), plus the reference's own known answers
(tests/test_linearfilter.py:33-42, tests/test_haplotyping.py:28-66,
tests/test_multigraph_shiftestimation.py:8-33).  One step is only HALF PINNED: the
reference projects a read start with offsetbasedgraph's
IndexedInterval.get_offset_at_position, whose source is not in the reference tree.
test_linearfilter.py pins forward starts (distance-to-node + offset); its single
reverse start sits mid-node, so the rule used here and in oracle/refshim for a reverse
start (-node, offset) -- distance-to-node + node size - offset -- agrees with it without
being told apart from distance-to-node + offset."""
import numpy as np


def strand_peaks(tags, peaksize, expansion, min_tags, max_tags):
    """[(summit, number of tags)] of every closed window of sorted ``tags``."""
    peaks = []
    if len(tags) < 2:
        return peaks
    open_at = 0
    for i in range(1, len(tags)):
        if tags[i] - tags[open_at] + 1 > peaksize:
            if min_tags <= i - open_at <= max_tags:
                window = tags[open_at:i]
                first = window[0]
                length = window[-1] + 1 - first + expansion
                line = [0] * length
                for p in set(window):                      # numpy: a repeated index counts once
                    line[p - first] += 1
                for p in set(window):
                    line[p - first + 2 * (expansion // 2)] -= 1
                cover, covers = 0, []
                for x in line:
                    cover += x
                    covers.append(cover)
                top = [x for x, c in enumerate(covers) if c == max(covers)]
                peaks.append((top[len(top) // 2] + first - expansion // 2, i - open_at))
            open_at = i
    return peaks


def pair_centers(plus, minus, peaksize):
    centers = []
    ip = im = im_back = 0
    overlapping = False
    while ip < len(plus) and im < len(minus):
        (pp, pn), (mp, mn) = plus[ip], minus[im]
        if pp - peaksize > mp:
            im += 1
        elif pp + peaksize < mp:
            ip += 1
            im = im_back
            overlapping = False
        else:
            if not overlapping:
                overlapping, im_back = True, im
            if 0.5 < pn / mn < 2 and pp < mp:
                centers.append((pp + mp) // 2)
            im += 1
    return centers


def add_to_line(centers, tags, peaksize, expansion, start, end):
    reach = peaksize + expansion // 2
    ic = it = it_back = 0
    overlapping = False
    while ic < len(centers) and it < len(tags):
        c, t = centers[ic], tags[it]
        if c - reach > t:
            it += 1
        elif c + reach < t:
            ic += 1
            it = it_back
            overlapping = False
        else:
            if not overlapping:
                overlapping, it_back = True, it
            start[max(t - c + peaksize, 0)] += 1
            end[min(t + expansion - c + peaksize, len(end) - 1)] -= 1
            it += 1


def smooth_flat(x, window_len=11):
    padded = np.r_[x[window_len - 1:0:-1], x, x[-1:-window_len:-1]]
    w = np.ones(window_len, "d")
    return np.convolve(w / w.sum(), padded, mode="valid")[window_len // 2:-(window_len // 2)]


def peak_model(positions, genome_size, lmfold=5, umfold=50, bw=300, expansion=10):
    """positions: {"+": {chrom: [...]}, "-": {chrom: [...]}}.  Returns a dict with d,
    the two model lines, the smoothed correlation, the alternative ds and the paired
    centres per chromosome; d is None with fewer than 100 pairs."""
    chroms = sorted(set(positions["+"]) | set(positions["-"]))
    plus = {c: sorted(int(p) for p in positions["+"][c]) for c in chroms}
    minus = {c: sorted(int(p) for p in positions["-"][c]) for c in chroms}
    total = sum(len(v) for v in plus.values()) + sum(len(v) for v in minus.values())
    peaksize = 2 * bw
    min_tags = int(round(float(total) * lmfold * peaksize / genome_size / 2))
    max_tags = int(round(float(total) * umfold * peaksize / genome_size / 2))
    window = 1 + 2 * peaksize + expansion
    lines = [[0] * window for _ in range(4)]
    centers = {}
    for c in chroms:
        pp = strand_peaks(plus[c], peaksize, expansion, min_tags, max_tags)
        mp = strand_peaks(minus[c], peaksize, expansion, min_tags, max_tags)
        if not pp or not mp:
            continue
        centers[c] = pair_centers(pp, mp, peaksize)
        add_to_line(centers[c], plus[c], peaksize, expansion, lines[0], lines[1])
        add_to_line(centers[c], minus[c], peaksize, expansion, lines[2], lines[3])
    out = {"centers": centers, "min_tags": min_tags, "max_tags": max_tags, "d": None,
           "n_pairs": sum(len(v) for v in centers.values())}
    if out["n_pairs"] < 100:
        return out
    plus_line = np.cumsum(np.array(lines[0]) + np.array(lines[1])).astype(np.int32)
    minus_line = np.cumsum(np.array(lines[2]) + np.array(lines[3])).astype(np.int32)
    minus_data = (minus_line - minus_line.mean()) / (minus_line.std() * len(minus_line))
    plus_data = (plus_line - plus_line.mean()) / (plus_line.std() * len(plus_line))
    ycorr = np.correlate(minus_data, plus_data, mode="full")[window - peaksize:window + peaksize]
    xcorr = np.linspace(len(ycorr) // 2 * -1, len(ycorr) // 2, num=len(ycorr))
    ycorr = smooth_flat(ycorr)
    best, alternatives = None, []
    for i in range(1, len(ycorr) - 1):
        if ycorr[i] > ycorr[i - 1] and ycorr[i] > ycorr[i + 1] and xcorr[i] > 0:
            alternatives.append(int(xcorr[i]))
            if best is None or ycorr[i] > ycorr[best]:
                best = i
    # the reference takes the first position of the whole curve that equals the best
    # positive local maximum (:170)
    best = int(np.where(ycorr == ycorr[best])[0][0])
    out.update(d=float(xcorr[best]), plus_line=plus_line, minus_line=minus_line, ycorr=ycorr,
               alternative_d=alternatives)
    return out


def most_covered_path(node_len, min_node, succ_ptr, succ, pred_ptr, read_paths):
    """Node ids of the path; read_paths: iterable of lists of signed node ids."""
    counts = {}
    for path in read_paths:
        for rp in path:
            counts[rp] = counts.get(rp, 0) + 1
    n = len(node_len)
    # nodes without any edge do not count (reference tests/test_haplotyping.py:10-25,50-66)
    sources = [v for v in range(n) if pred_ptr[v + 1] == pred_ptr[v]
               and (succ_ptr[v + 1] > succ_ptr[v] or n == 1)]
    assert len(sources) == 1, "Only works when graph has one start node"
    v, path = sources[0], []
    while True:
        path.append(v + min_node)
        nxt, best = None, -1
        for w in succ[succ_ptr[v]:succ_ptr[v + 1]]:
            if counts.get(int(w) + min_node, 0) > best:
                nxt, best = int(w), counts.get(int(w) + min_node, 0)
        if nxt is None:
            return path
        v = nxt


def start_positions(node_len, min_node, path, starts):
    """starts: iterable of (signed node id, offset).  {"+": [...], "-": [...]} in read
    order."""
    distance, walked = {}, 0
    for node in path:
        distance[node] = walked
        walked += int(node_len[node - min_node])
    out = {"+": [], "-": []}
    for node, offset in starts:
        if abs(node) not in distance:
            continue
        if node > 0:
            out["+"].append(distance[node] + int(offset))
        else:
            out["-"].append(distance[-node] + int(node_len[-node - min_node]) - int(offset))
    return out

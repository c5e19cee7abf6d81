"""Seeded synthetic variation graphs and reads as FLAT ARRAYS.

Pure numpy and free of imports from this package on purpose: ``bench.py --impl
reference`` loads this file by path, so the reference arm gets the same inputs
as the GPU arm without importing (or loading anything of) the product.
``synthetic.py`` wraps the results in the package's ``Graph`` / ``ReadBatch``.

The generator of one chromosome graph:

* a backbone of ``backbone_bp`` bases cut at variant sites (mean spacing
  ``1/site_rate``); each site is a SNP (two 1-bp alternative nodes), an
  optional node (insertion / deletion: one alternative node that a direct edge
  skips) or, when ``mnp_frac > 0``, two alternative nodes of different length
  (non-integer scales in the control projection, reference
  control/linearmap.py:43-49);
* ``nest_levels`` refinement passes replace alternative nodes of three or more
  bases by ``pre -> {x | y} -> post`` (a bubble inside the branch of a bubble;
  two passes give nesting depth 3 as BASELINE config #5 asks);
* node ids are contiguous from ``min_node`` in topological order, as vg emits
  them, and nodes are laid out in id order (``node_off``);
* reads are random walks of ``read_length`` bases; 30 % cluster around peak
  centres, strands are 50/50, reverse-strand reads use negated node ids and
  reversed paths (obg ``Interval.get_reverse`` convention), and a fraction of
  exact duplicates is injected so the duplicate filter has work to do.

Nothing here is on the measured path.
"""
import numpy as np


class FlatGraph:
    """The six flat arrays of a graph (same attribute names as the package's
    ``Graph``; the oracle and the shim bridge only read these)."""

    def __init__(self, node_len, src, dst, min_node=1):
        node_len = np.asarray(node_len, dtype=np.int64)
        src = np.asarray(src, dtype=np.int64)
        dst = np.asarray(dst, dtype=np.int64)
        n = node_len.size
        self.n_nodes = n
        self.min_node = int(min_node)
        self.node_len = node_len
        self.node_off = np.r_[0, np.cumsum(node_len)].astype(np.int64)
        order = np.lexsort((dst, src))
        self.succ = dst[order].astype(np.int32)
        self.succ_ptr = np.r_[0, np.cumsum(np.bincount(src, minlength=n))].astype(np.int64)
        order = np.lexsort((src, dst))
        self.pred = src[order].astype(np.int32)
        self.pred_ptr = np.r_[0, np.cumsum(np.bincount(dst, minlength=n))].astype(np.int64)

    @property
    def n_edges(self):
        return int(self.succ.size)

    @property
    def size_bp(self):
        return int(self.node_off[-1])

    @property
    def node_indexes(self):
        return self.node_off


class FlatReads:
    """Reads SoA (same attribute names as the package's ``ReadBatch``)."""

    def __init__(self, start_node, start_off, end_node, end_off, path_ptr, path):
        self.start_node, self.start_off = start_node, start_off
        self.end_node, self.end_off = end_node, end_off
        self.path_ptr, self.path = path_ptr, path

    def __len__(self):
        return int(self.start_node.size)

    def arrays(self):
        return (self.start_node, self.start_off, self.end_node, self.end_off,
                self.path_ptr, self.path)


def graph_arrays(backbone_bp, site_rate=0.02, seed=0, snp_frac=0.85, ins_frac=0.10,
                 del_frac=0.05, mnp_frac=0.0, indel_mean=5.0, indel_max=None,
                 cuts="choice", nest_levels=0, nest_frac=0.5, tri_frac=0.0):
    """Returns (node_len int64[n], src int64[m], dst int64[m]) with ids ascending
    along every edge.  ``cuts``: "choice" draws the cut points without replacement
    (configs #1, #2: kept so their graphs stay what round 1 measured); "gaps" draws
    geometric gaps between sites (memory ~ number of sites, for the 10^8 bp
    chromosomes of configs #3, #4)."""
    rng = np.random.default_rng(seed)
    backbone_bp = int(backbone_bp)
    if cuts == "choice":
        n_sites = max(0, int(rng.binomial(max(backbone_bp - 2, 0), site_rate)))
        # cut points strictly inside the backbone, distinct -> segment lengths >= 1
        if n_sites:
            cut = np.sort(rng.choice(np.arange(1, backbone_bp, dtype=np.int64),
                                     size=min(n_sites, backbone_bp - 1), replace=False))
        else:
            cut = np.zeros(0, dtype=np.int64)
    else:
        expect = int(backbone_bp * site_rate * 1.05) + 64
        cut = np.cumsum(rng.geometric(site_rate, size=expect).astype(np.int64))
        while cut.size and cut[-1] < backbone_bp:      # (never with the 5 % margin at these sizes)
            more = np.cumsum(rng.geometric(site_rate, size=expect).astype(np.int64)) + cut[-1]
            cut = np.r_[cut, more]
        cut = cut[cut < backbone_bp]
    n_sites = cut.size
    seg_len = np.diff(np.r_[0, cut, backbone_bp]).astype(np.int64)  # n_sites+1

    fr = [snp_frac, ins_frac + del_frac, mnp_frac] + ([tri_frac] if tri_frac > 0 else [])
    fr = np.array(fr, dtype=np.float64)
    fr = fr / fr.sum()
    kind = rng.choice(fr.size, size=n_sites, p=fr)    # 0 SNP, 1 optional, 2 MNP, 3 three alleles
    n_alt = np.where(kind == 1, 1, np.where(kind == 3, 3, 2)).astype(np.int64)

    def geom(size):
        g = rng.geometric(1.0 / indel_mean, size=size).astype(np.int64)
        if indel_max is not None:
            g = np.minimum(g, indel_max - 1)
        return 1 + g

    alt_a = np.ones(n_sites, dtype=np.int64)
    alt_b = np.ones(n_sites, dtype=np.int64)
    opt = kind == 1
    alt_a[opt] = geom(int(opt.sum()))
    mnp = kind == 2
    alt_a[mnp] = geom(int(mnp.sum()))
    alt_b[mnp] = geom(int(mnp.sum()))

    # node layout: B0, alts(site0), B1, alts(site1), ..., B_last
    per_site_nodes = 1 + n_alt                       # backbone node + alts
    first_idx = np.r_[0, np.cumsum(per_site_nodes)]  # index of B_i
    n_nodes = int(first_idx[-1] + 1)
    node_len = np.empty(n_nodes, dtype=np.int64)
    b_idx = np.r_[first_idx[:-1], first_idx[-1]].astype(np.int64)   # n_sites+1
    node_len[b_idx] = seg_len
    a_idx = b_idx[:-1] + 1
    node_len[a_idx] = alt_a
    two = n_alt >= 2
    node_len[a_idx[two] + 1] = alt_b[two]
    three = n_alt == 3
    node_len[a_idx[three] + 2] = 1
    is_alt = np.ones(n_nodes, dtype=bool)
    is_alt[b_idx] = False

    # edges
    src = [b_idx[:-1], a_idx]                        # B_i -> a ; a -> B_{i+1}
    dst = [a_idx, b_idx[1:]]
    src += [b_idx[:-1][two], a_idx[two] + 1]          # B_i -> b ; b -> B_{i+1}
    dst += [a_idx[two] + 1, b_idx[1:][two]]
    src += [b_idx[:-1][three], a_idx[three] + 2]      # B_i -> c ; c -> B_{i+1}
    dst += [a_idx[three] + 2, b_idx[1:][three]]
    src += [b_idx[:-1][opt]]                          # B_i -> B_{i+1} (skip edge)
    dst += [b_idx[1:][opt]]
    src = np.concatenate(src)
    dst = np.concatenate(dst)
    for _ in range(int(nest_levels)):
        node_len, src, dst, is_alt = _nest(node_len, src, dst, is_alt, nest_frac, rng)
    return node_len, src, dst


def _nest(node_len, src, dst, is_alt, frac, rng):
    """One refinement pass: a fraction of the alternative nodes of >= 3 bases become
    pre -> {x | y} -> post.  The four nodes take the place of the old one (and
    everything behind shifts), in-edges go to ``pre`` and out-edges leave ``post``,
    so ids still ascend along every edge.  Only x and y stay candidates for the
    next pass (a bubble inside x is one level deeper)."""
    n = node_len.size
    cand = np.flatnonzero(is_alt & (node_len >= 3))
    pick = cand[rng.random(cand.size) < frac]
    split = np.zeros(n, dtype=bool)
    split[pick] = True
    shift = 3 * (np.cumsum(split) - split)            # nodes inserted before each old node
    new_idx = np.arange(n, dtype=np.int64) + shift    # index of the old node / of its `pre`
    L = node_len[pick]
    pre = 1 + (rng.random(pick.size) * (L - 2)).astype(np.int64)          # 1 .. L-2
    post = 1 + (rng.random(pick.size) * (L - pre - 1)).astype(np.int64)   # 1 .. L-pre-1
    mid = L - pre - post                                                   # >= 1
    other = 1 + (rng.random(pick.size) * np.maximum(mid, 1) * 1.5).astype(np.int64)
    out_len = np.empty(n + 3 * pick.size, dtype=np.int64)
    out_alt = np.zeros(out_len.size, dtype=bool)
    keep = ~split
    out_len[new_idx[keep]] = node_len[keep]
    out_alt[new_idx[keep]] = False                    # only the new inner branches nest further
    p = new_idx[pick]
    out_len[p], out_len[p + 1], out_len[p + 2], out_len[p + 3] = pre, mid, other, post
    out_alt[p + 1] = True
    out_alt[p + 2] = True
    s2 = new_idx[src] + np.where(split[src], 3, 0)    # out-edges leave `post`
    d2 = new_idx[dst]                                 # in-edges enter `pre`
    src = np.concatenate([s2, p, p, p + 1, p + 2])
    dst = np.concatenate([d2, p + 1, p + 2, p + 3, p + 3])
    return out_len, src, dst, out_alt


def make_flat_graph(backbone_bp, site_rate=0.02, seed=0, min_node=1, **kw):
    node_len, src, dst = graph_arrays(backbone_bp, site_rate, seed, **kw)
    return FlatGraph(node_len, src, dst, min_node)


def _walk(graph, start_idx, start_off, length, rng):
    """Random forward walks.  Returns (path_ptr, path_idx0, end_off)."""
    n_reads = start_idx.size
    node_len = graph.node_len
    succ_ptr, succ = graph.succ_ptr, graph.succ
    cur = start_idx.astype(np.int64).copy()
    rem = length - (node_len[cur] - start_off)        # bases still to place
    steps = [cur.copy()]
    owners = [np.arange(n_reads, dtype=np.int64)]
    active = np.flatnonzero(rem > 0)
    end_off = np.where(rem <= 0, node_len[cur] + rem, node_len[cur])
    while active.size:
        c = cur[active]
        deg = succ_ptr[c + 1] - succ_ptr[c]
        alive = deg > 0
        active = active[alive]
        if not active.size:
            break
        c = c[alive]
        deg = deg[alive]
        pick = (rng.random(active.size) * deg).astype(np.int64)
        nxt = succ[succ_ptr[c] + pick].astype(np.int64)
        cur[active] = nxt
        steps.append(nxt)
        owners.append(active.copy())
        rem[active] -= node_len[nxt]
        r = rem[active]
        end_off[active] = np.where(r <= 0, node_len[nxt] + r, node_len[nxt])
        active = active[r > 0]
    owner = np.concatenate(owners)
    nodes = np.concatenate(steps)
    order = np.argsort(owner, kind="stable")
    nodes = nodes[order]
    counts = np.bincount(owner, minlength=n_reads)
    path_ptr = np.r_[0, np.cumsum(counts)].astype(np.int64)
    return path_ptr, nodes, end_off.astype(np.int64)


def read_arrays(graph, n_reads, read_length=36, fragment_length=200, seed=0,
                peak_frac=0.3, peak_spacing=50000, dup_frac=0.05):
    """Returns the six arrays of ``n_reads`` reads (duplicates included):
    start_node, start_off, end_node, end_off (int32), path_ptr (int64), path (int32)."""
    rng = np.random.default_rng(seed)
    n_reads = int(n_reads)
    G = int(graph.node_off[-1])
    n_dup = int(n_reads * dup_frac)
    n_base = n_reads - n_dup
    strand_neg = rng.random(n_base) < 0.5
    in_peak = rng.random(n_base) < peak_frac
    n_centres = max(1, G // peak_spacing)
    centres = (np.arange(n_centres) + 0.5) * (G / n_centres)
    c = centres[rng.integers(0, n_centres, size=n_base)]
    jitter = rng.normal(0.0, fragment_length / 4.0, size=n_base)
    pos_peak = np.where(strand_neg,
                        c + fragment_length / 2.0 - read_length + jitter,
                        c - fragment_length / 2.0 + jitter)
    pos_uni = rng.random(n_base) * G
    pos = np.where(in_peak, pos_peak, pos_uni)
    pos = np.clip(pos, 0, G - 1).astype(np.int64)
    idx = np.searchsorted(graph.node_off, pos, side="right") - 1
    idx = np.clip(idx, 0, graph.n_nodes - 1)
    off = pos - graph.node_off[idx]
    path_ptr, nodes, end_off = _walk(graph, idx, off, read_length, rng)

    if n_dup:
        src = rng.integers(0, n_base, size=n_dup)
        where = rng.integers(0, n_base + 1, size=n_dup)
        order = np.argsort(np.r_[np.arange(n_base) * 2, where * 2 - 1],
                           kind="stable")
        pick = np.r_[np.arange(n_base), src][order]
    else:
        pick = np.arange(n_base)

    lens = np.diff(path_ptr)
    first = nodes[path_ptr[:-1]]
    last = nodes[path_ptr[1:] - 1]
    mn = graph.min_node
    nl = graph.node_len
    # forward representation
    s_node = first + mn
    s_off = off
    e_node = last + mn
    e_off = end_off
    # reverse-strand reads: negate + reverse (obg Interval.get_reverse)
    rs_node = -(last + mn)
    rs_off = nl[last] - end_off
    re_node = -(first + mn)
    re_off = nl[first] - off
    start_node = np.where(strand_neg, rs_node, s_node)
    start_off = np.where(strand_neg, rs_off, s_off)
    end_node = np.where(strand_neg, re_node, e_node)
    end_offs = np.where(strand_neg, re_off, e_off)

    # per-read path, reversed and negated for the reverse strand
    read_of = np.repeat(np.arange(n_base), lens)
    k = np.arange(nodes.size) - path_ptr[read_of]
    neg = strand_neg[read_of]
    src_pos = np.where(neg, path_ptr[read_of] + lens[read_of] - 1 - k,
                       path_ptr[read_of] + k)
    path = (nodes[src_pos] + mn) * np.where(neg, -1, 1)

    # apply duplicate injection / ordering
    new_lens = lens[pick]
    new_ptr = np.r_[0, np.cumsum(new_lens)].astype(np.int64)
    gather = (np.repeat(path_ptr[:-1][pick] - new_ptr[:-1], new_lens)
              + np.arange(new_ptr[-1]))
    return (start_node[pick].astype(np.int32), start_off[pick].astype(np.int32),
            end_node[pick].astype(np.int32), end_offs[pick].astype(np.int32),
            new_ptr, path[gather].astype(np.int32))


def make_flat_reads(graph, n_reads, read_length=36, fragment_length=200, seed=0, **kw):
    return FlatReads(*read_arrays(graph, n_reads, read_length, fragment_length, seed, **kw))


# ---- the named workloads of BASELINE.json `configs` ---------
#
# A workload = a list of chromosomes (name, backbone bp) + graph parameters + the
# TOTAL read counts, split over the chromosomes in proportion to their length.

HG19_BP = [249250621, 243199373, 198022430, 191154276, 180915260, 171115067, 159138663,
           146364022, 141213431, 135534747, 135006516, 133851895, 115169878, 107349540,
           102531392, 90354753, 81195210, 78077248, 59128983, 63025520, 48129895, 51304566,
           155270560, 59373566]                                   # chr1..22, X, Y
_ARA = np.array([30.4, 19.7, 23.5, 18.6, 27.0])                   # TAIR10 Mbp, scaled to 135 Mbp


def _ara_bp():
    bp = np.floor(_ARA / _ARA.sum() * 135_000_000).astype(np.int64)
    bp[0] += 135_000_000 - bp.sum()
    return [int(b) for b in bp]


CONFIGS = {
    # BASELINE.json configs[0]: MHC-shaped stand-in for the missing MHC blobs
    "mhc_shaped": dict(index=1, chromosomes=[("mhc", 5_000_000)], site_rate=0.02,
                       graph=dict(cuts="choice"),
                       n_sample=200_000, n_control=200_000, read_length=36,
                       fragment_length=135, has_control=True, global_min=False),
    # configs[1]: the benchmarked workload
    "chr22_scale": dict(index=2, chromosomes=[("chr22", 51_000_000)], site_rate=0.02,
                        graph=dict(cuts="choice"),
                        n_sample=5_000_000, n_control=5_000_000, read_length=36,
                        fragment_length=200, has_control=True, global_min=False),
    # configs[2]: five chromosomes, the sample is its own control (extension [10000],
    # ratio 1), genome-wide background as global_min (arabidopsis run.sh:6-11: F=230, R=50)
    "arabidopsis_shaped": dict(index=3, chromosomes=[("%d" % (i + 1), b)
                                                     for i, b in enumerate(_ara_bp())],
                               site_rate=0.01, graph=dict(cuts="gaps"),
                               n_sample=20_000_000, n_control=None, read_length=50,
                               fragment_length=230, has_control=False, global_min=True),
    # configs[3]: 24 chromosome graphs with hg19 lengths, sharded over the GPUs
    "human_wg": dict(index=4, chromosomes=[(n, b) for n, b in zip(
                         [str(i) for i in range(1, 23)] + ["X", "Y"], HG19_BP)],
                     site_rate=0.02, graph=dict(cuts="gaps"),
                     n_sample=40_000_000, n_control=40_000_000, read_length=36,
                     fragment_length=200, has_control=True, global_min=True),
    # configs[4]: dense variation, long indels, bubbles nested to depth 3
    # (9.2 Mbp of backbone + the alternative alleles = a 20 Mbp graph)
    "dense_variation": dict(index=5, chromosomes=[("dense", 9_200_000)], site_rate=0.10,
                            graph=dict(cuts="gaps", snp_frac=0.75, ins_frac=0.07, del_frac=0.03,
                                       mnp_frac=0.05, tri_frac=0.10, indel_mean=50.0,
                                       indel_max=500, nest_levels=2, nest_frac=0.5),
                            n_sample=10_000_000, n_control=10_000_000, read_length=36,
                            fragment_length=200, has_control=True, global_min=False),
}


def config_seed(config_index, chrom_index=0):
    return 20260101 + 1000 * config_index + chrom_index


def scaled(cfg, factor):
    """The same workload with every chromosome and both read counts divided by
    ``factor`` (same read density, same graph statistics): the bounded samples of
    the CPU arm and the oracle-sized parity cases."""
    out = dict(cfg)
    out["chromosomes"] = [(n, max(1000, int(b // factor))) for n, b in cfg["chromosomes"]]
    out["n_sample"] = max(100, int(cfg["n_sample"] // factor))
    out["n_control"] = None if cfg["n_control"] is None else max(100, int(cfg["n_control"] // factor))
    out["scale_factor"] = factor
    return out


def split_reads(cfg):
    """Reads per chromosome, in proportion to the backbone length (largest remainder)."""
    bp = np.array([b for _, b in cfg["chromosomes"]], dtype=np.float64)

    def split(total):
        if total is None:
            return [None] * bp.size
        exact = total * bp / bp.sum()
        base = np.floor(exact).astype(np.int64)
        rest = int(total - base.sum())
        order = np.argsort(-(exact - base), kind="stable")
        base[order[:rest]] += 1
        return [int(x) for x in base]
    return split(cfg["n_sample"]), split(cfg["n_control"])


def global_min_of(cfg):
    """`-u`/`-G` of the reference's whole-genome runs: unique reads x fragment length /
    genome size (arabidopsis run.sh:6-11); here from the nominal totals, so that every
    rank and the CPU arm use the same number without a pre-pass."""
    if not cfg.get("global_min"):
        return None
    n = cfg["n_control"] if cfg["n_control"] is not None else cfg["n_sample"]
    total_bp = sum(b for _, b in cfg["chromosomes"])
    return float(n) * cfg["fragment_length"] / float(total_bp)


def chromosome_inputs(cfg, chrom_index, make_graph=make_flat_graph, make_reads=make_flat_reads,
                      seed_index=None):
    """(graph, sample reads, control reads or None) of one chromosome of a workload.
    Seeds: SURVEY.md §8d (20260101 + 1000 config + chromosome; reads +100 / +200).
    ``seed_index`` replaces the chromosome index in the seed (weak scaling: every rank
    gets its own copy of a one-chromosome workload)."""
    name, bp = cfg["chromosomes"][chrom_index]
    seed = config_seed(cfg["index"], chrom_index if seed_index is None else seed_index)
    ns, nc = split_reads(cfg)
    g = make_graph(bp, cfg["site_rate"], seed=seed, **cfg["graph"])
    R, F = cfg["read_length"], cfg["fragment_length"]
    s = make_reads(g, ns[chrom_index], R, F, seed=seed + 100)
    c = None
    if nc[chrom_index] is not None:
        c = make_reads(g, nc[chrom_index], R, F, seed=seed + 200, peak_frac=0.0, dup_frac=0.0)
    return g, s, c

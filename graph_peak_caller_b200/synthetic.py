"""Seeded synthetic variation graphs and reads.

The reference ships no runnable large data set (its MHC blobs are missing), so
every configuration in BASELINE.json is produced by this one generator:

* a backbone of ``backbone_bp`` bases cut at variant sites (mean spacing
  ``1/site_rate``); each site is a SNP (two 1-bp alternative nodes), an
  optional node (insertion / deletion: one alternative node that a direct edge
  skips) or, when ``mnp_frac > 0``, two alternative nodes of different length
  (this is what gives non-integer scales in the control projection,
  reference control/linearmap.py:43-49);
* node ids are contiguous from ``min_node`` in topological order, as vg emits
  them, and nodes are laid out in id order (``node_off``);
* reads are random walks of ``read_length`` bases; 30 % cluster around peak
  centres, strands are 50/50, reverse-strand reads use negated node ids and
  reversed paths (obg ``Interval.get_reverse`` convention), and a fraction of
  exact duplicates is injected so the duplicate filter has work to do.

Everything is plain numpy on the host; nothing here is on the measured path.
"""
import numpy as np

from .graph import Graph
from .reads import ReadBatch


def make_graph(backbone_bp, site_rate=0.02, seed=0, min_node=1,
               snp_frac=0.85, ins_frac=0.10, del_frac=0.05, mnp_frac=0.0,
               indel_mean=5.0, indel_max=None):
    """Returns a :class:`Graph` whose ids ascend topologically."""
    rng = np.random.default_rng(seed)
    backbone_bp = int(backbone_bp)
    n_sites = max(0, int(rng.binomial(max(backbone_bp - 2, 0), site_rate)))
    # cut points strictly inside the backbone, distinct -> segment lengths >= 1
    if n_sites:
        cuts = np.sort(rng.choice(np.arange(1, backbone_bp, dtype=np.int64),
                                  size=min(n_sites, backbone_bp - 1),
                                  replace=False))
    else:
        cuts = np.zeros(0, dtype=np.int64)
    n_sites = cuts.size
    seg_len = np.diff(np.r_[0, cuts, backbone_bp]).astype(np.int64)  # n_sites+1

    fr = np.array([snp_frac, ins_frac + del_frac, mnp_frac], dtype=np.float64)
    fr = fr / fr.sum()
    kind = rng.choice(3, size=n_sites, p=fr)          # 0 SNP, 1 optional, 2 MNP
    n_alt = np.where(kind == 1, 1, 2).astype(np.int64)

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
    two = n_alt == 2
    node_len[a_idx[two] + 1] = alt_b[two]

    # edges
    src = [b_idx[:-1], a_idx]                        # B_i -> a ; a -> B_{i+1}
    dst = [a_idx, b_idx[1:]]
    src += [b_idx[:-1][two], a_idx[two] + 1]          # B_i -> b ; b -> B_{i+1}
    dst += [a_idx[two] + 1, b_idx[1:][two]]
    src += [b_idx[:-1][opt]]                          # B_i -> B_{i+1} (skip edge)
    dst += [b_idx[1:][opt]]
    src = np.concatenate(src)
    dst = np.concatenate(dst)
    return Graph.from_edges(node_len, src, dst, min_node=min_node)


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


def make_reads(graph, n_reads, read_length=36, fragment_length=200, seed=0,
               peak_frac=0.3, peak_spacing=50000, dup_frac=0.05):
    """Returns a :class:`ReadBatch` of ``n_reads`` reads (duplicates included)."""
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
    return ReadBatch(start_node[pick].astype(np.int32),
                     start_off[pick].astype(np.int32),
                     end_node[pick].astype(np.int32),
                     end_offs[pick].astype(np.int32),
                     new_ptr, path[gather].astype(np.int32))


# ---- the named workloads of BASELINE.json -------------------

CONFIGS = {
    # name: (backbone_bp, site_rate, n_sample, n_control, read_len, frag_len, has_control)
    "mhc_shaped": dict(backbone_bp=5_000_000, site_rate=0.02, n_sample=200_000,
                       n_control=200_000, read_length=36, fragment_length=135,
                       has_control=True),
    "chr22_scale": dict(backbone_bp=51_000_000, site_rate=0.02,
                        n_sample=5_000_000, n_control=5_000_000,
                        read_length=36, fragment_length=200, has_control=True),
}


def config_seed(config_index, chrom_index=0):
    return 20260101 + 1000 * config_index + chrom_index

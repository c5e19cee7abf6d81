"""The one-pass form of scipy's Bellman-Ford predecessors that max_paths_kernel uses
(csrc/postprocess.cu: mp_solve) against scipy itself.

The reference calls ``csgraph.shortest_path(subgraph, return_predecessors=True, method="BF")``
(postprocess/graphs.py:346-347) on components whose vertex order is not topological.  Its
sweeps give, for a vertex with several tight in-edges, the predecessor that arrived first:
the smallest (sweep, index) over the tight in-neighbours, where a vertex passes its final
distance on in the sweep it got it in if its own predecessor has a smaller index, else in the
next one.  This file restates mp_solve's push pass in Python and holds it against scipy on
random DAGs with scrambled indexes and many ties."""
import numpy as np
from scipy.sparse import csr_matrix
from scipy.sparse.csgraph import shortest_path

INF = 1 << 60


def one_pass(n, edges, w, first, topo):
    dist, pred, sweep = [INF] * n, [-9999] * n, [0] * n
    dist[first] = 0
    for j in topo:
        if dist[j] >= INF:
            continue
        c = dist[j] + w[j]
        sj = sweep[j] + (1 if pred[j] > j else 0)
        for t in edges.get(j, ()):
            d2 = dist[t]
            if c < d2 or (c == d2 and (sj < sweep[t] or (sj == sweep[t] and j < pred[t]))):
                dist[t], sweep[t], pred[t] = c, sj, j
    return dist, pred


def test_one_pass_predecessors_equal_scipy_bellman_ford():
    rng = np.random.default_rng(1)
    compared = 0
    for _ in range(2000):
        n = int(rng.integers(2, 16))
        rank = rng.permutation(n)                 # topological position of every vertex
        topo = [int(v) for v in np.argsort(rank)]
        w = [-int(x) for x in rng.integers(0, 3, n)]      # weight of every out-edge of a vertex
        edges, rows, cols, data = {}, [], [], []
        for a in range(n):
            for b in range(n):
                if rank[a] < rank[b] and rng.random() < 0.35:
                    edges.setdefault(a, []).append(b)
                    rows.append(a)
                    cols.append(b)
                    data.append(w[a])
        if not rows:
            continue
        m = csr_matrix((np.array(data, dtype=float), (rows, cols)), shape=(n, n))
        first = int(rng.integers(0, n))
        d, p = shortest_path(m, method="BF", return_predecessors=True, indices=[first])
        dist, pred = one_pass(n, edges, w, first, topo)
        assert [INF if np.isinf(x) else int(x) for x in d[0]] == dist
        for t in range(n):
            if t != first and dist[t] < INF:
                assert int(p[0][t]) == pred[t]
                compared += 1
    assert compared > 3000

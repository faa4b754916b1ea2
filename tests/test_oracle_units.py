"""Unit pins for the oracle's relaxation / taumodel / sampling / graphcut restatements."""
import math

import numpy as np
import pytest

from oracle import iq_oracle as O


def topk_mask(v, k):
    order = np.argsort(v, kind="stable")
    m = np.zeros(v.size, dtype=bool)
    m[order[:k]] = True
    return m


@pytest.mark.parametrize("seed", range(20))
def test_relaxation_equals_topk_intersection(seed):
    """the incremental loop of src/relaxation.jl:21-36 equals
    'top-k by (value,index) per aux, AND with top-dbsize of primary, grow k'."""
    r = np.random.default_rng(seed)
    n = int(r.integers(20, 400))
    naux = int(r.integers(1, 3))
    tol = float(r.choice([0.05, 0.1, 0.3, 1.0]))
    levels = int(r.choice([2, 5, 1000]))
    D = r.integers(0, levels, n).astype(float)
    Ds = [r.integers(0, levels, n).astype(float) for _ in range(naux)]
    dis = r.random(n) < 0.1
    D[dis] = np.inf
    for a in Ds:
        a[dis] = np.inf
    got = O.relaxation(D, Ds, tol)
    npat = int((~np.isinf(D)).sum())
    dbsize = npat if np.all(D[~np.isinf(D)] == 0) else math.ceil(tol * npat)
    base = topk_mask(D, dbsize)
    frac = 0.1 * (dbsize / npat)
    while True:
        k = math.ceil(frac * npat)
        m = base.copy()
        for a in Ds:
            m &= topk_mask(a, k)
        if m.any():
            break
        frac = min(frac + 0.1, 1)
    assert np.array_equal(got, np.flatnonzero(m))


def test_taumodel_single_source_is_rank_probability():
    D = np.array([5.0, 1.0, 3.0, 3.0, 9.0])
    ev = np.arange(5)
    p = O.taumodel(ev, D, [])
    ranks = np.array([3, 1, 2, 2, 4])
    P = (5 - ranks + 1) / (5 - ranks + 1).sum()
    assert np.allclose(p, P, rtol=1e-12)
    assert np.array_equal(O.taumodel(np.array([2]), D, []), [1.0])


def test_taumodel_all_equal_is_uniform():
    D = np.zeros(100)
    p = O.taumodel(np.arange(100), D, [np.ones(100)])
    assert np.allclose(p, p[0])


def test_sampling_walk():
    p = np.array([0.1, 0.2, 0.3, 0.4])
    assert O.sample_weighted(0.0, p) == 0
    assert O.sample_weighted(0.05, p) == 0
    assert O.sample_weighted(0.2, p) == 1
    assert O.sample_weighted(0.95, p) == 3
    assert O.sample_weighted(1.0, p) == 3
    assert abs(O.julia_sum(np.full(5000, 0.1)) - 500.0) < 1e-9


@pytest.mark.parametrize("seed", range(6))
def test_graphcut_matches_networkx(seed):
    """Independent check of the Dinic restatement: the keep-mask equals the complement of
    'can reach the sink in the residual graph' computed from networkx's own max-flow."""
    import networkx as nx
    r = np.random.default_rng(seed)
    shape = [(4, 9), (9, 4), (3, 6, 5), (6, 5, 3)][seed % 4]
    dim = [0, 1, 0, 2][seed % 4]
    A, B = r.standard_normal(shape), r.standard_normal(shape)
    M = O.graphcut(A, B, dim)
    # rebuild the graph literally as src/graphcut.jl:18-70 does
    sz = A.shape
    nvox = A.size
    lin = np.arange(nvox).reshape(sz, order="F")
    Af, Bf = A.ravel(order="F"), B.ravel(order="F")
    G = nx.DiGraph()
    eps = np.finfo(float).eps
    for d in range(A.ndim):
        for ind in np.ndindex(*sz):
            if ind[d] >= sz[d] - 1:
                continue
            nxt = tuple(i + 1 if k == d else i for k, i in enumerate(ind))
            w = tuple(i + 2 if k == d else i for k, i in enumerate(ind))
            u, v = lin[ind], lin[nxt]
            Du, Dv = abs(Af[u] - Bf[u]), abs(Af[v] - Bf[v])
            gAu, gBu = abs(Af[v] - Af[u]), abs(Bf[v] - Bf[u])
            ok = w[d] < sz[d]
            gAv = abs(A[w] - A[nxt]) if ok else gAu
            gBv = abs(B[w] - B[nxt]) if ok else gBu
            c = (Du + Dv) / (gAu + gAv + gBu + gBv + eps)
            G.add_edge(int(u), int(v), capacity=c)
            G.add_edge(int(v), int(u), capacity=c)
    for ind in np.ndindex(*sz):
        if ind[dim] == 0:
            G.add_edge("s", int(lin[ind]))  # no capacity attr = infinite
        if ind[dim] == sz[dim] - 1:
            G.add_edge(int(lin[ind]), "t")
    R = nx.algorithms.flow.preflow_push(G, "s", "t")
    # nodes that can reach t in the residual graph
    resid = nx.DiGraph()
    resid.add_nodes_from(R.nodes)
    for u, v, a in R.edges(data=True):
        if a["capacity"] - a["flow"] > 1e-12:
            resid.add_edge(u, v)
    can = nx.ancestors(resid, "t")
    keep = np.array([i not in can for i in range(nvox)]).reshape(sz, order="F")
    assert np.array_equal(M, keep)
    # sanity: source slice kept, sink slice not
    first = tuple(slice(0, 1) if k == dim else slice(None) for k in range(A.ndim))
    last = tuple(slice(sz[k] - 1, sz[k]) if k == dim else slice(None) for k in range(A.ndim))
    assert M[first].all() and not M[last].any()

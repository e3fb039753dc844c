"""GPU parity of the candidate-pair evaluation (csrc/pairs.inl): connected components of the candidate graph,
dense components as 64 x 64 member blocks (k_block_d, k_block_n), sparse ones per candidate (k_pairs_sparse) --
against the oracle and against the per-candidate kernel forced for everything (TRACS_PAIRS=sparse)."""
import numpy as np
import pytest

import tracs_b200
from tracs_b200 import synth

pytestmark = pytest.mark.gpu
COLS = ("rows", "cols", "dist", "ncomp")


def _cmp(res, orc):
    r, c, d, f, nn = orc
    assert res["rows"].tolist() == r.tolist()
    assert res["cols"].tolist() == c.tolist()
    assert res["dist"].tolist() == d.tolist()
    assert res["ncomp"].tolist() == nn.tolist()


def _both_ways(s, dist, monkeypatch, oracle_mod, **kw):
    orc = oracle_mod.pairsnp_ascii(s, dist=dist, n_threads=8)
    res = tracs_b200.pairsnp_matrix(s, dist=dist, **kw)
    st = tracs_b200.last_stats()
    assert st["ms_refine"] > 0, "the filter-and-refine path is the one under test"
    _cmp(res, orc)
    monkeypatch.setenv("TRACS_PAIRS", "sparse")
    ref = tracs_b200.pairsnp_matrix(s, dist=dist, **kw)
    monkeypatch.delenv("TRACS_PAIRS")
    for k in COLS:
        assert res[k].tolist() == ref[k].tolist()
    return res, orc


@pytest.mark.parametrize("n,n_clusters,big", [(300, 60, 0), (1500, 300, 200), (1600, 30, 0), (520, 520, 0), (1400, 200, 150)])
def test_component_blocks_match_oracle(oracle_mod, monkeypatch, n, n_clusters, big):
    """Cliques of ~5 and ~53 members, all singletons (no candidates), and -- `big` -- one clique of 200 / 150 members
    scattered over the sample range (several 64 x 64 blocks per component, off-diagonal blocks included) next to small
    ones; N-rich rows and gap runs so that the N intersections are not trivial."""
    s = synth.generate(n, 200_000, p_var=0.06, n_clusters=n_clusters, mu=4, p_N=0.01, p_amb=0.01, seed=7 + n_clusters, gaps=4)
    if big:
        rng = np.random.default_rng(big)
        who = rng.choice(n, size=big, replace=False)
        var = np.nonzero((s != s[0]).any(axis=0))[0]
        for k in who[1:]:                     # copies of one sample with a few private substitutions each
            s[k] = s[who[0]]
            for t in rng.choice(var, size=3, replace=False):
                s[k, t] = ord("A") if s[k, t] != ord("A") else ord("G")
        s[rng.random(s.shape) < 0.01] = ord("N")
    res, orc = _both_ways(s, 20, monkeypatch, oracle_mod)
    if n_clusters < n:
        assert len(res["rows"]) > n
    if big:
        assert len(res["rows"]) > big * (big - 1) // 2


def test_sparse_components_take_the_per_candidate_kernel(oracle_mod, monkeypatch):
    """A chain (every sample within the threshold of its two neighbours only) is one large, sparse component: it
    must not be evaluated as m x m blocks; mixed with small cliques in the same alignment."""
    n, L = 400, 200_000
    s = synth.generate(n, L, p_var=0.06, n_clusters=40, mu=3, p_N=0.005, seed=99, gaps=2)
    rng = np.random.default_rng(5)
    sites = rng.choice(L, size=100 * 12, replace=False)
    for k in range(1, 100):            # samples 0..99: s_k = s_0 + 12 k substitutions  =>  d(s_a, s_b) = 12 |a - b|
        s[k] = s[0]
        for t in sites[:12 * k]:
            s[k, t] = ord("A") if s[0, t] != ord("A") else ord("C")
    res, orc = _both_ways(s, 20, monkeypatch, oracle_mod)
    chain = [(int(a), int(b)) for a, b in zip(res["rows"], res["cols"]) if b < 100]
    assert (0, 1) in chain and (98, 99) in chain and (0, 2) not in chain and len(chain) == 99


def test_blocks_with_fused_likelihood_and_two_file_ranges(oracle_mod, monkeypatch):
    s = synth.generate(500, 150_000, p_var=0.08, n_clusters=25, mu=4, p_N=0.01, seed=31, gaps=3)
    days = np.random.default_rng(1).integers(0, 120, size=500).astype(np.int32)
    res, orc = _both_ways(s, 25, monkeypatch, oracle_mod, days=days)
    r, c, d = orc[0], orc[1], orc[2]
    dt = np.abs(days[r.astype(int)] * 86400.0 - days[c.astype(int)] * 86400.0) / 31556952.0
    op0, oeK = oracle_mod.trans_dist(d.astype(np.int32), dt, 29.903, 73.0, 0.01)
    assert np.allclose(res["p0_log"], op0, rtol=1e-6, atol=0)
    pos = dt > 0
    assert np.allclose(res["eK"][pos], oeK[pos], rtol=1e-6, atol=0)
    for n1 in (100, 257):
        got = tracs_b200.pairsnp_matrix(s, dist=25, i_end=n1, j_start=n1)
        _cmp(got, oracle_mod.pairsnp_ascii(s, i_end=n1, j_start=n1, dist=25, n_threads=4))
    got = tracs_b200.pairsnp_matrix(s, dist=25, want_ncomp=False)
    assert got["rows"].tolist() == r.tolist() and got["dist"].tolist() == d.tolist() and not got["ncomp"].any()


@pytest.mark.parametrize("p_N", [0.002, 0.3])
def test_dense_and_sparse_N_intersections_agree(oracle_mod, monkeypatch, p_N):
    """|N_i n N_j| of the component blocks: the summary-guided kernel (k_block_n) and the dense AND + POPC contraction
    over whole N-plane rows (k_block_d<1>, picked automatically for N-rich alignments such as BASELINE configs[3]),
    each forced on both kinds of data, against the oracle."""
    s = synth.generate(500, 150_000, p_var=0.08, n_clusters=25, mu=4, p_N=p_N, p_amb=0.02, seed=int(p_N * 1000) + 3, gaps=3)
    dist = 25 if p_N < 0.1 else 12
    orc = oracle_mod.pairsnp_ascii(s, dist=dist, n_threads=8)
    for mode in ("dense", "sparse", None):
        if mode:
            monkeypatch.setenv("TRACS_NBLOCKS", mode)
        else:
            monkeypatch.delenv("TRACS_NBLOCKS")
        res = tracs_b200.pairsnp_matrix(s, dist=dist)
        assert tracs_b200.last_stats()["ms_refine"] > 0
        _cmp(res, orc)
    assert len(orc[0]) > 500


@pytest.mark.parametrize("n,L,p_var,p_N,dist,expect_refine", [(600, 150_001, 0.06, 0.002, 25, True), (1100, 140_000, 0.06, 0.01, 25, True),
                                                              (400, 160_001, 0.06, 0.0, 20, True), (520, 66_030, 0.06, 0.002, 2000, False)])
def test_sparse_nplane_stores_match_oracle(oracle_mod, monkeypatch, n, L, p_var, p_N, dist, expect_refine):
    """Sparse N on the early-extraction ingest: the main launch stores only the 256-site groups of the N bit-plane that hold
    an N (pack_emit_n<true>); the rest of the buffer is stale and must never be read. The buffer is first filled with
    ones by an N-only call of the same shape, so that a consumer that strays off the block summaries shows up.
    Against the oracle and against full stores (TRACS_NPLANE=always); per-candidate kernel forced as well; dist = 2000
    leaves the prefilter out, so the per-edge tail (k_ncomp) reads the sparsely stored plane."""
    s = synth.generate(n, L, p_var=p_var, n_clusters=max(2, n // 12), mu=4, p_N=p_N, p_amb=0.01, seed=n + 11, gaps=3)
    orc = oracle_mod.pairsnp_ascii(s, dist=dist, n_threads=8)
    poison = np.full_like(s, ord("N"))
    poison[:, ::97] = ord("A")
    monkeypatch.setenv("TRACS_INGEST", "early")
    for mode, sparse in (("sparse", 1.0), ("always", 0.0)):
        monkeypatch.setenv("TRACS_NPLANE", "always")
        tracs_b200.pairsnp_matrix(poison, dist=0)        # leaves an all-ones N-plane in the cached buffer
        monkeypatch.setenv("TRACS_NPLANE", mode)
        res = tracs_b200.pairsnp_matrix(s, dist=dist)
        st = tracs_b200.last_stats()
        assert st["sparse_nplane"] == sparse and st["n_early_sites"] > 0
        if expect_refine:
            assert st["ms_refine"] > 0, "filter-and-refine (component blocks) is the path under test"
        _cmp(res, orc)
        monkeypatch.setenv("TRACS_PAIRS", "sparse")
        _cmp(tracs_b200.pairsnp_matrix(s, dist=dist), orc)
        monkeypatch.delenv("TRACS_PAIRS")
    monkeypatch.delenv("TRACS_NPLANE")
    frac_n = float(np.mean((s == ord("N")) | (s == ord("-"))))
    res = tracs_b200.pairsnp_matrix(s, dist=dist)       # the density rule on the first chunk decides
    assert tracs_b200.last_stats()["sparse_nplane"] == (1.0 if 1.0 - (1.0 - frac_n) ** 128 <= 0.30 else tracs_b200.last_stats()["sparse_nplane"])
    _cmp(res, orc)


def test_dense_N_keeps_whole_rows(oracle_mod, monkeypatch):
    """N-rich packed input (BASELINE configs[3] shape): the first chunk shows dense N, every N-plane word is stored and the
    dense contraction (k_block_d<1>) runs."""
    s = synth.generate(700, 80_000, p_var=0.05, n_clusters=30, mu=4, p_N=0.3, p_amb=0.02, seed=5, gaps=2)
    orc = oracle_mod.pairsnp_ascii(s, dist=12, n_threads=8)
    monkeypatch.setenv("TRACS_INGEST", "early")
    res = tracs_b200.pairsnp_matrix(s, dist=12)
    st = tracs_b200.last_stats()
    assert st["sparse_nplane"] == 0.0 and st["n_early_sites"] > 0
    _cmp(res, orc)

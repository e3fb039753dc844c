"""GPU, larger-than-oracle sizes: size-independent properties of the sweep on device-generated
alignments (no oracle run over the whole input): path equivalence (filter-and-refine == full-length
tile sweep), shard union == single sweep, idempotence, ordering, and spot checks of individual
pairs against the per-site definition computed in NumPy from the two rows."""
import ctypes as C

import numpy as np
import pytest

import tracs_b200
from tracs_b200 import _lib

pytestmark = pytest.mark.gpu


class DevAln:
    def __init__(self, n, L, **kw):
        self.n, self.L = n, L
        self.pitch = (L + 127) // 128 * 128
        self.p = C.c_void_p()
        _lib.check(_lib.lib().tracs_dev_alloc(C.byref(self.p), n * self.pitch))
        tracs_b200.synth_device(self.p.value, n, L, self.pitch, **kw)

    def row(self, i):
        out = np.empty(self.L, np.uint8)
        _lib.check(_lib.lib().tracs_memcpy_d2h(out.ctypes.data, C.c_void_p(self.p.value + i * self.pitch), self.L))
        return out

    def free(self):
        _lib.lib().tracs_dev_free(self.p)


def _pair_truth(a, b):
    from oracle import oracle
    ma, mb = oracle.masks_of(a), oracle.masks_of(b)
    d = int(((ma & mb) == 0).sum())
    nn = int(len(a) - ((ma == 15) | (mb == 15)).sum())
    return d, nn


@pytest.mark.parametrize("n,L,clusters,dist", [(12000, 400_000, 150, 20), (3000, 2_000_000, 40, 20)])
def test_large_properties(n, L, clusters, dist, monkeypatch):
    aln = DevAln(n, L, seed=11, p_var=0.03, n_clusters=clusters, mu=5.0, p_N=1e-3, p_amb=0.002, gc=0.5)
    try:
        res = tracs_b200.pairsnp_device(aln.p.value, n, L, aln.pitch, dist=dist)
        st = tracs_b200.last_stats()
        assert st["ms_refine"] > 0, "filter-and-refine path expected at this shape"
        assert st["n_early_sites"] > 0, "early-extraction ingest expected at this shape"
        # the two-pass ingest (k_pack + k_gather) gives the same edge table and the same variable sites
        monkeypatch.setenv("TRACS_INGEST", "split")
        two = tracs_b200.pairsnp_device(aln.p.value, n, L, aln.pitch, dist=dist)
        st2 = tracs_b200.last_stats()
        monkeypatch.delenv("TRACS_INGEST")
        assert st2["n_early_sites"] == 0 and st2["n_variable_sites"] == st["n_variable_sites"]
        for k in ("rows", "cols", "dist", "ncomp"):
            assert np.array_equal(res[k], two[k])
        E = len(res["rows"])
        assert E > 1000
        key = (res["rows"] << np.uint64(32)) | res["cols"]
        assert np.all(key[1:] > key[:-1]), "edges must be strictly (row, col) ordered"
        assert np.all(res["rows"] < res["cols"]) and np.all(res["cols"] < n) and np.all(res["dist"] <= dist)
        assert np.all(res["ncomp"] <= L)
        # idempotence
        res2 = tracs_b200.pairsnp_device(aln.p.value, n, L, aln.pitch, dist=dist)
        for k in ("rows", "cols", "dist", "ncomp"):
            assert np.array_equal(res[k], res2[k])
        # the full-length tile sweep gives the same edge list
        full = tracs_b200.pairsnp_device(aln.p.value, n, L, aln.pitch, dist=dist, full_sweep=True)
        assert tracs_b200.last_stats()["ms_refine"] == 0
        for k in ("rows", "cols", "dist", "ncomp"):
            assert np.array_equal(res[k], full[k])
        # shard union == single
        parts = [tracs_b200.pairsnp_device(aln.p.value, n, L, aln.pitch, dist=dist, shard_rank=r, shard_world=4) for r in range(4)]
        pk = np.concatenate([(p["rows"] << np.uint64(32)) | p["cols"] for p in parts])
        order = np.argsort(pk, kind="stable")
        assert np.array_equal(pk[order], key)
        assert np.array_equal(np.concatenate([p["dist"] for p in parts])[order], res["dist"])
        # spot checks: emitted edges, and random pairs that were not emitted
        rng = np.random.default_rng(0)
        for e in rng.choice(E, size=25, replace=False):
            i, j = int(res["rows"][e]), int(res["cols"][e])
            d, nn = _pair_truth(aln.row(i), aln.row(j))
            assert d == int(res["dist"][e]) and nn == int(res["ncomp"][e]), (i, j)
        emitted = set(key.tolist())
        checked = 0
        while checked < 15:
            i, j = sorted(rng.choice(n, size=2, replace=False).tolist())
            if ((i << 32) | j) in emitted:
                continue
            d, _ = _pair_truth(aln.row(i), aln.row(j))
            assert d > dist, (i, j, d)
            checked += 1
        # monotone in the threshold: dist=5 edges are exactly the d<=5 subset
        small = tracs_b200.pairsnp_device(aln.p.value, n, L, aln.pitch, dist=5)
        m = res["dist"] <= 5
        assert np.array_equal(small["rows"], res["rows"][m]) and np.array_equal(small["cols"], res["cols"][m])
        assert np.array_equal(small["ncomp"], res["ncomp"][m])
    finally:
        aln.free()


def test_c1_shape_dense_and_thresholded():
    # BASELINE.json configs[0] shape (1000 x 2.8 Mb): dense output = every pair, exactly once, ordered
    n, L = 1000, 2_800_000
    aln = DevAln(n, L, seed=1, p_var=0.01, n_clusters=20, mu=5.0, p_N=1e-3, gc=0.33)
    try:
        dense = tracs_b200.pairsnp_device(aln.p.value, n, L, aln.pitch, dist=2147483647)
        assert len(dense["rows"]) == n * (n - 1) // 2
        ii, jj = np.triu_indices(n, 1)
        assert np.array_equal(dense["rows"], ii.astype(np.uint64)) and np.array_equal(dense["cols"], jj.astype(np.uint64))
        thr = tracs_b200.pairsnp_device(aln.p.value, n, L, aln.pitch, dist=20)
        m = dense["dist"] <= 20
        assert np.array_equal(thr["rows"], dense["rows"][m]) and np.array_equal(thr["dist"], dense["dist"][m])
        assert np.array_equal(thr["ncomp"], dense["ncomp"][m])
        rng = np.random.default_rng(1)
        for e in rng.choice(len(ii), size=10, replace=False):
            d, nn = _pair_truth(aln.row(int(ii[e])), aln.row(int(jj[e])))
            assert d == int(dense["dist"][e]) and nn == int(dense["ncomp"][e])
    finally:
        aln.free()


def test_c5_shape_dense_variation():
    # BASELINE.json configs[4] shape scaled down: EVERY site variable, unambiguous bases, no N
    n, L = 4000, 500_000
    aln = DevAln(n, L, seed=5, p_var=1.0, n_clusters=40, mu=5.0, p_N=0.0, p_amb=0.0, gc=0.5, gaps=0)
    try:
        res = tracs_b200.pairsnp_device(aln.p.value, n, L, aln.pitch, dist=20)
        st = tracs_b200.last_stats()
        assert st["n_variable_sites"] > 0.99 * L and st["ms_refine"] > 0
        full = tracs_b200.pairsnp_device(aln.p.value, n, L, aln.pitch, dist=20, full_sweep=True)
        for k in ("rows", "cols", "dist", "ncomp"):
            assert np.array_equal(res[k], full[k])
        assert np.all(res["ncomp"] == L)
        rng = np.random.default_rng(5)
        E = len(res["rows"])
        assert E > 1000
        for e in rng.choice(E, size=10, replace=False):
            d, nn = _pair_truth(aln.row(int(res["rows"][e])), aln.row(int(res["cols"][e])))
            assert d == int(res["dist"][e]) and nn == L
    finally:
        aln.free()

"""Pins the recombination filter's binomial CDF.

Boost's ibetac-based binomial CDF is absent from this image: the reference build here (oracle/_ref), the C oracle and
the CUDA kernel k_filter_recomb all sum the pmf term by term. This test shows that the substitution cannot change a
result: on every window the filter evaluates for the golden alignments and for adversarial, near-threshold SNP
layouts, the keep / drop decision is identical under (a) the pmf sum, (b) the regularised incomplete beta function
evaluated with a continued fraction written from the published algorithm (oracle/binom_check.py) and (c)
scipy.stats.binom -- and it records the smallest margin between 1 - CDF and the threshold 0.05 / d that was met."""
import json
import os

import numpy as np
import pytest
from scipy import stats

from oracle import binom_check as bc
from oracle import oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
G = json.load(open(os.path.join(GOLD, "golden.json")))


def _scipy_cdf(n, p, k):
    return 1.0 if k >= n else float(stats.binom.cdf(k, n, p))


def _read_fasta(path):
    import gzip
    op = gzip.open if path.endswith(".gz") else open
    seqs, cur = [], []
    for ln in op(path, "rt"):
        if ln.startswith(">"):
            if cur:
                seqs.append("".join(cur))
            cur = []
        else:
            cur.append(ln.strip())
    seqs.append("".join(cur))
    return np.array([np.frombuffer(s.encode(), dtype=np.uint8) for s in seqs])


def test_cdf_implementations_agree_on_a_grid():
    worst = 0.0
    for n in (2, 3, 7, 50, 101, 400, 1000, 5000, 10001):
        for p in (1e-6, 1e-4, 0.003, 0.05, 0.3, 0.75, 0.999):
            for k in sorted({0, 1, 2, n // 100, n // 10, n // 2, n - 2, n - 1}):
                if 0 <= k < n:
                    a, b, c = bc.binom_cdf_pmf_sum(n, p, k), bc.binom_cdf_ibeta(n, p, k), _scipy_cdf(n, p, k)
                    worst = max(worst, abs(a - b), abs(b - c))
    assert worst < 1e-10, worst   # lgamma cancellation at n ~ 1e4 limits the pmf sum to ~1e-11 absolute


@pytest.mark.parametrize("case", G["filter"], ids=lambda c: c["fasta"])
def test_filter_decisions_do_not_depend_on_the_cdf(case):
    """Every pair of the golden filter cases (made by oracle/_ref): the Python restatement of the windowing reproduces
    the reference's filt column, and all three CDFs take the same decisions."""
    seqs = _read_fasta(os.path.join(GOLD, case["fasta"]))
    m = oracle.masks_of(seqs)
    L = m.shape[1]
    margins = []
    for r, c, d, f in zip(case["rows"], case["cols"], case["d"], case["filt"]):
        snp = np.nonzero((m[r] & m[c]) == 0)[0].tolist()
        assert len(snp) == d
        got = [bc.filtered_distance(snp, L, cdf) for cdf in (bc.binom_cdf_pmf_sum, bc.binom_cdf_ibeta, _scipy_cdf)]
        assert got[0][0] == f, "windowing restatement differs from the reference build"
        assert got[1][0] == f and got[2][0] == f, "a CDF implementation flips a keep / drop decision"
        margins.append(got[0][1])
    finite = [x for x in margins if np.isfinite(x)]
    assert finite and min(finite) > 1e-9, min(finite)
    print("smallest |p - 0.05/d| / (0.05/d) over %d pairs of %s: %.3g" % (len(finite), case["fasta"], min(finite)))


def test_adversarial_near_threshold_layouts():
    """SNP layouts built to sit near the decision boundary: clumps whose density is tuned so that 1 - CDF straddles
    0.05 / d. 4000 layouts; the three CDFs must agree on every decision; the smallest margin met is recorded."""
    rng = np.random.default_rng(2024)
    L = 200_000
    smallest, tested, flips = float("inf"), 0, 0
    for trial in range(4000):
        n_bg = int(rng.integers(3, 60))
        snp = set(rng.choice(L, size=n_bg, replace=False).tolist())
        start = int(rng.integers(1000, L - 6000))
        width = int(rng.integers(20, 5000))
        k = int(rng.integers(2, 12))
        snp |= set((start + rng.choice(width, size=min(k, width), replace=False)).tolist())
        snp = sorted(snp)
        res = [bc.filtered_distance(snp, L, cdf) for cdf in (bc.binom_cdf_pmf_sum, bc.binom_cdf_ibeta, _scipy_cdf)]
        flips += not (res[0][0] == res[1][0] == res[2][0])
        if np.isfinite(res[0][1]):
            smallest = min(smallest, res[0][1])
            tested += 1
    assert flips == 0 and tested > 3000
    # double-precision CDFs differ by ~1e-13 relative here: a decision could only flip inside that band
    assert smallest > 1e-10, smallest
    print("adversarial layouts: %d pairs, smallest relative margin %.3g" % (tested, smallest))


def test_c_oracle_equals_python_restatement():
    rng = np.random.default_rng(7)
    L = 60_000
    for _ in range(30):
        a = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=L)]
        b = a.copy()
        sites = rng.choice(L, size=int(rng.integers(2, 80)), replace=False)
        s0 = int(rng.integers(0, L - 500))
        sites = np.unique(np.concatenate([sites, s0 + rng.choice(400, size=int(rng.integers(0, 30)), replace=False)]))
        b[sites] = np.where(a[sites] == ord("A"), ord("C"), ord("A"))
        r, c, d, f, nn = oracle.pairsnp_ascii(np.stack([a, b]), filter=True)
        assert int(d[0]) == len(sites)
        assert int(f[0]) == bc.filtered_distance(sorted(sites.tolist()), L, bc.binom_cdf_ibeta)[0]

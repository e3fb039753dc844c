"""Generates the golden fixtures in this directory from the UNMODIFIED reference
(oracle/_ref, built by oracle/build_ref.sh from /root/reference/src) and, for the CSV fixture, the
reference's own tracs/distance.py run on top of it. Needs /root/reference; run once, commit outputs.

    python tests/golden/make_golden.py
"""
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import numpy as np

from oracle import refmod
from tracs_b200 import synth

ref, variant = refmod.load()
IMAX = 2147483647


def aln_cases():
    cases = []
    specs = [dict(n=5, L=40, p_var=0.3, n_clusters=2, mu=2, p_N=0.05, p_amb=0.2, seed=1, three_base=True, dist=10),
             dict(n=12, L=333, p_var=0.1, n_clusters=3, mu=3, p_N=0.02, p_amb=0.05, seed=2, lowercase=0.2, odd_chars=0.02, dist=IMAX),
             dict(n=40, L=2000, p_var=0.05, n_clusters=4, mu=4, p_N=0.01, p_amb=0.02, seed=3, dist=25),
             dict(n=9, L=64, p_var=0.5, n_clusters=2, mu=1, p_N=0.3, p_amb=0.3, seed=4, three_base=True, dist=0),
             dict(n=130, L=1500, p_var=0.05, n_clusters=5, mu=3, p_N=0.01, p_amb=0.0, seed=5, dist=40)]
    for k, sp in enumerate(specs):
        sp = dict(sp)
        dist = sp.pop("dist")
        seqs = synth.generate(**sp)
        name = "aln%d.fasta%s" % (k, ".gz" if k % 2 else "")
        synth.write_fasta(os.path.join(HERE, name), seqs, width=[0, 60, 70, 0, 80][k], descriptions=bool(k % 2))
        cases.append((name, dist))
    return cases


def main():
    out = {"reference_variant": variant, "pairsnp": [], "two_file": [], "filter": []}
    cases = aln_cases()
    for name, dist in cases:
        p = os.path.join(HERE, name)
        r = ref.pairsnp(fasta=[p], n_threads=1, dist=dist, filter=False)
        out["pairsnp"].append({"fasta": name, "dist": dist, "rows": list(r[0]), "cols": list(r[1]), "d": list(r[2]), "names": list(r[3]),
                               "filt": list(r[4]), "ncomp": list(r[5])})
    # query x db mode: aln2 split into two files
    seqs = synth.generate(n=40, L=2000, p_var=0.05, n_clusters=4, mu=4, p_N=0.01, p_amb=0.02, seed=3)
    synth.write_fasta(os.path.join(HERE, "query.fasta"), seqs[:7], names=["q%d" % i for i in range(7)])
    synth.write_fasta(os.path.join(HERE, "db.fasta.gz"), seqs[7:], names=["d%d" % i for i in range(33)])
    r = ref.pairsnp(fasta=[os.path.join(HERE, "query.fasta"), os.path.join(HERE, "db.fasta.gz")], n_threads=2, dist=30, filter=False)
    out["two_file"].append({"fasta": ["query.fasta", "db.fasta.gz"], "dist": 30, "rows": list(r[0]), "cols": list(r[1]), "d": list(r[2]),
                            "names": list(r[3]), "filt": list(r[4]), "ncomp": list(r[5])})
    # recombination filter (stand-in binomial CDF: parity vs real Boost unpinned, see oracle header)
    for name, dist in cases[:3]:
        r = ref.pairsnp(fasta=[os.path.join(HERE, name)], n_threads=1, dist=IMAX, filter=True)
        out["filter"].append({"fasta": name, "dist": IMAX, "rows": list(r[0]), "cols": list(r[1]), "d": list(r[2]), "filt": list(r[4])})
    # transcluster grid
    grid = {"cases": []}
    for lamb, beta in ((29.903, 73.0), (5.3, 6.0)):
        for thr in (0.01, 1e-6):
            N = np.repeat(np.array([0, 1, 2, 3, 5, 8, 13, 20, 35, 50]), 8)
            days = np.tile(np.array([0, 1, 2, 7, 30, 60, 120, 180]), 10)
            dt = days * 86400.0 / 31556952.0
            p0, eK = ref.trans_dist(N.astype(int).tolist(), dt.tolist(), lamb, beta, thr)
            grid["cases"].append({"lamb": lamb, "beta": beta, "thr": thr, "N": N.tolist(), "days": days.tolist(), "p0_log": list(p0), "eK": list(eK)})
    out["trans_dist"] = grid
    out["lprob_k_given_N"] = []
    from scipy.special import gammaln
    for (N, k, delta, lamb, beta) in [(7, 4, 0.16963, 3.0, 52.0), (0, 0, 0.01, 29.903, 73.0), (3, 9, 0.5, 5.3, 6.0), (2, 1, 0.0, 5.3, 6.0)]:
        r = ref.lprob_k_given_N(N, k, delta, lamb, beta, gammaln(range(40)).tolist())
        out["lprob_k_given_N"].append({"args": [N, k, delta, lamb, beta], "out": list(r)})
    rng = np.random.default_rng(0)
    counts = np.round(rng.random((6, 4)) * 10 * (rng.random((6, 4)) > 0.3))
    counts[2] = 0
    alphas = [0.1, 0.5, 0.2, 0.3]
    out["calculate_posteriors"] = [{"counts": counts.tolist(), "alphas": alphas, "keep": keep, "threshold": 0.05,
                                    "out": np.asarray(ref.calculate_posteriors(counts, alphas, keep, 0.05)).tolist()} for keep in (True, False)]
    json.dump(out, open(os.path.join(HERE, "golden.json"), "w"))

    # ---- the reference's own CLI (tracs/distance.py, unchanged) on top of the reference module ----
    sys.modules["TRACS"] = ref
    sys.modules["pyfastx"] = types.ModuleType("pyfastx")  # tracs/utils.py:8 imports it at module top; unused here
    sys.path.insert(0, "/root/reference")
    from tracs.distance import main as distance_main
    seqs = synth.generate(n=30, L=3000, p_var=0.05, n_clusters=3, mu=3, p_N=0.01, p_amb=0.02, seed=9)
    names = synth.write_fasta(os.path.join(HERE, "cli_combined.fasta.gz"), seqs, names=["seq%d" % (i + 1) for i in range(30)])
    rng = np.random.default_rng(3)
    with open(os.path.join(HERE, "cli_dates.csv"), "w") as f:
        f.write("sample,date\n")
        for nm in names:
            f.write("%s,2020-%02d-%02d\n" % (nm, int(rng.integers(1, 6)), int(rng.integers(1, 28))))
    for tag, extra in (("meta", ["--meta", os.path.join(HERE, "cli_dates.csv"), "-K", "100"]), ("nometa", [])):
        sys.argv = ["", "--msa", os.path.join(HERE, "cli_combined.fasta.gz"), "-o", os.path.join(HERE, "cli_%s.csv" % tag),
                    "--snp_threshold", "40", "-t", "2"] + extra
        distance_main()
    # the reference's tracs/cluster.py (pure Python + scipy) on the CSV just written
    import importlib
    import tracs.cluster as ref_cluster
    for tag, thr, dist in (("snp10", 10, "snp"), ("ek5", 5, "expectedK"), ("direct", 0.05, "direct")):
        importlib.reload(ref_cluster)  # fresh function-static name table per run (tracs/cluster.py:11-21)
        sys.argv = ["", "-d", os.path.join(HERE, "cli_meta.csv"), "-o", os.path.join(HERE, "cluster_%s.csv" % tag), "-c", str(thr), "-D", dist]
        ref_cluster.main()
    # crafted KAT from the reference's tests/test_trans_distance.py
    with open(os.path.join(HERE, "kat.fasta"), "w") as f:
        f.write(">seq1\nACGTACGTAC\n>seq2\nACGTACGTAN\n>seq3\nACGTACGTGG\n")
    with open(os.path.join(HERE, "kat_dates.csv"), "w") as f:
        f.write("sample,date\nseq1,2020-01-01\nseq2,2020-01-02\nseq3,2020-01-02\n")
    sys.argv = ["", "--msa", os.path.join(HERE, "kat.fasta"), "--meta", os.path.join(HERE, "kat_dates.csv"), "-o",
                os.path.join(HERE, "kat.csv"), "-K", "10", "--snp_threshold", "5"]
    distance_main()
    print(open(os.path.join(HERE, "kat.csv")).read())


if __name__ == "__main__":
    main()

"""Seeded synthetic alignments on the host (NumPy) for oracle-scale tests and fixtures: generator
G(N, L, p_var, C, mu, p_N, p_amb, gc, seed) of SURVEY 8d. The large bench configs are generated
directly in device memory by tracs_synth_device (same model, hash-based stream)."""
import gzip
import numpy as np

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)
_AMB2 = {(0, 1): "M", (0, 2): "R", (0, 3): "W", (1, 2): "S", (1, 3): "Y", (2, 3): "K"}
_AMB3 = {0: "B", 1: "D", 2: "H", 3: "V"}  # code for "everything but base b"


def generate(n, L, p_var=0.01, n_clusters=20, mu=5.0, p_N=1e-3, p_amb=0.0, gc=0.5, seed=1, gaps=2, lowercase=0.0,
             odd_chars=0.0, three_base=False):
    """Returns uint8[n][L] ASCII."""
    rng = np.random.default_rng(seed)
    pr = np.array([(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2])
    anc = rng.choice(4, size=L, p=pr)
    V = int(round(p_var * L))
    var_sites = rng.choice(L, size=V, replace=False) if V else np.zeros(0, np.int64)
    founders = np.repeat(anc[None, :], n_clusters, axis=0)
    if V:
        flip = rng.random((n_clusters, V)) < 0.3
        alt = (anc[var_sites][None, :] + rng.integers(1, 4, size=(n_clusters, V))) % 4
        founders[:, var_sites] = np.where(flip, alt, anc[var_sites][None, :])
    cl = rng.integers(0, n_clusters, size=n)
    codes = founders[cl].astype(np.int64)
    if V:
        for s in range(n):
            k = rng.poisson(mu)
            if k:
                sites = rng.choice(var_sites, size=min(k, V), replace=False)
                codes[s, sites] = (codes[s, sites] + rng.integers(1, 4, size=sites.size)) % 4
    seqs = BASES[codes]
    if p_amb > 0 and V:
        for s in range(n):
            m = rng.random(V) < p_amb
            for site in var_sites[m]:
                b = int(codes[s, site])
                if three_base and rng.random() < 0.3:
                    drop = int((b + rng.integers(1, 4)) % 4)
                    seqs[s, site] = ord(_AMB3[drop])
                else:
                    o = int((b + rng.integers(1, 4)) % 4)
                    seqs[s, site] = ord(_AMB2[(min(b, o), max(b, o))])
    if p_N > 0:
        seqs[rng.random((n, L)) < p_N] = ord("N")
    glen = max(1, L // 1000)
    for s in range(n):
        for _ in range(gaps):
            if L > glen:
                g0 = int(rng.integers(0, L - glen))
                seqs[s, g0:g0 + glen] = ord("-")
    if odd_chars > 0:
        m = rng.random((n, L)) < odd_chars
        seqs[m] = rng.choice(np.frombuffer(b"X?.*nN-U0", dtype=np.uint8), size=int(m.sum()))
    if lowercase > 0:
        m = (rng.random((n, L)) < lowercase) & (seqs >= 65) & (seqs <= 90)
        seqs[m] += 32
    return np.ascontiguousarray(seqs)


def write_fasta(path, seqs, names=None, width=0, gz=None, descriptions=False):
    n = seqs.shape[0]
    names = names or ["s%d" % i for i in range(n)]
    if gz is None:
        gz = str(path).endswith(".gz")
    op = gzip.open if gz else open
    with op(path, "wb") as f:
        for i in range(n):
            hdr = ">" + names[i] + ((" sample %d len=%d" % (i, seqs.shape[1])) if descriptions else "")
            f.write(hdr.encode() + b"\n")
            row = seqs[i].tobytes()
            if width and width > 0:
                for k in range(0, len(row), width):
                    f.write(row[k:k + width] + b"\n")
            else:
                f.write(row + b"\n")
    return names

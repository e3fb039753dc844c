"""CPU property tests (hypothesis): the library's FASTA loader (tracs_read_fasta) against the oracle's
kseq-semantics reader on randomly assembled files -- headers with comments, ragged line widths, CR/LF,
blank lines, stray spaces/tabs, lower case, IUPAC and junk symbols, optional gzip."""
import gzip
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import tracs_b200

ALPHABET = "ACGTacgtNn-MRWSYKVHDBmrwsykvhdbXx?.*0"


@st.composite
def fasta_files(draw):
    n = draw(st.integers(1, 6))
    L = draw(st.integers(0, 90))
    recs = []
    for i in range(n):
        seq = "".join(draw(st.lists(st.sampled_from(ALPHABET), min_size=L, max_size=L)))
        name = "s%d" % i + draw(st.sampled_from(["", "_x", ".1"]))
        comment = draw(st.sampled_from(["", " a comment", "\tlen=%d" % L, " > not a header"]))
        width = draw(st.sampled_from([0, 1, 7, 60]))
        eol = draw(st.sampled_from(["\n", "\r\n"]))
        lines = [seq] if width == 0 or L == 0 else [seq[k:k + width] for k in range(0, L, width)]
        noise = draw(st.sampled_from(["", " ", "\t", "\n"]))
        body = eol.join(ln + noise for ln in lines)
        recs.append(">" + name + comment + eol + body + eol)
    lead = draw(st.sampled_from(["", "\n", "junk before the first header\n"]))
    return lead + "".join(recs), n, L


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(data=fasta_files(), gz=st.booleans())
def test_loader_matches_oracle_reader(oracle_mod, tmp_path_factory, data, gz):
    text, n, L = data
    d = tmp_path_factory.mktemp("fa")
    p = str(d / ("x.fa.gz" if gz else "x.fa"))
    with (gzip.open(p, "wb") if gz else open(p, "wb")) as f:
        f.write(text.encode())
    # TRACS_FASTA_SINK_CHECK: the reader also feeds a row sink while parsing (what the FASTA entry point streams to the
    # device) and raises if those rows are not exactly the finished matrix, in order
    os.environ["TRACS_FASTA_SINK_CHECK"] = "1"
    a, names = tracs_b200.read_fasta(p)
    exp = oracle_mod.pairsnp([p], dist=2147483647)
    assert names == exp[3]
    assert a.shape == (n, L)
    # the multi-threaded reader (plain files; forced on tiny inputs) must agree byte for byte
    os.environ["TRACS_FASTA_PAR_MIN"] = "0"
    try:
        a4, names4 = tracs_b200.read_fasta(p, n_threads=4)
    finally:
        del os.environ["TRACS_FASTA_PAR_MIN"]
        del os.environ["TRACS_FASTA_SINK_CHECK"]
    assert names4 == names and a4.shape == a.shape and np.array_equal(a4, a)
    if n >= 2 and L > 0:
        got = oracle_mod.pairsnp_ascii(a, dist=2147483647)
        assert got[2].tolist() == exp[2] and got[4].tolist() == exp[5]

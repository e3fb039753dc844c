"""GPU parity of the 4-bit packed input path (csrc/pack4.inl: k_encode, k_pack4) and of the host-streaming
ingest that is built on it, against the CPU oracle and against the ASCII-resident path; plus the native last
step of the site-sharded sweep (tracs_site_shard_finish). Everything goes through the C ABI."""
import ctypes as C

import numpy as np
import pytest

import tracs_b200
from tracs_b200 import _lib, synth

pytestmark = pytest.mark.gpu
IMAX = 2147483647
COLS = ("rows", "cols", "dist", "ncomp")


def _cmp(res, orc):
    r, c, d, f, nn = orc
    assert res["rows"].tolist() == r.tolist()
    assert res["cols"].tolist() == c.tolist()
    assert res["dist"].tolist() == d.tolist()
    assert res["ncomp"].tolist() == nn.tolist()


def _dev(arr):
    """numpy uint8 matrix -> (device pointer holder, pitch)"""
    arr = np.ascontiguousarray(arr, dtype=np.uint8)
    p = C.c_void_p()
    _lib.check(_lib.lib().tracs_dev_alloc(C.byref(p), max(16, arr.size)))
    _lib.check(_lib.lib().tracs_memcpy_h2d(p, arr.ctypes.data, arr.size))
    return p


@pytest.mark.parametrize("n,L,dist", [(2, 1, IMAX), (5, 31, IMAX), (7, 32, 3), (33, 33, IMAX), (128, 1000, 40), (129, 4097, 25),
                                      (300, 20001, 30), (257, 1023, 0), (64, 5000, IMAX), (3, 70000, IMAX)])
def test_packed_host_and_device_match_oracle(oracle_mod, n, L, dist, monkeypatch):
    """The same alignment as ASCII (streamed + encoded on the device, in several chunks), as ASCII kept resident,
    as a packed host matrix and as a packed device matrix: four routes, one edge table, equal to the oracle."""
    s = synth.generate(n, L, p_var=0.05, n_clusters=4, mu=3, p_N=0.02, p_amb=0.05, seed=n * 7 + L, lowercase=0.05,
                       odd_chars=0.01, three_base=True)
    orc = oracle_mod.pairsnp_ascii(s, dist=dist, n_threads=4)
    monkeypatch.setenv("TRACS_STREAM_CHUNK_BYTES", str(max(1, n // 3) * ((L + 31) // 32 * 32)))   # >= 3 chunks
    _cmp(tracs_b200.pairsnp_matrix(s, dist=dist), orc)
    monkeypatch.setenv("TRACS_HOST_INGEST", "ascii")
    _cmp(tracs_b200.pairsnp_matrix(s, dist=dist), orc)
    monkeypatch.delenv("TRACS_HOST_INGEST")
    nib, pitch = tracs_b200.pack_nibbles(s)
    _cmp(tracs_b200.pairsnp_packed_host(nib, L, dist=dist), orc)
    d = _dev(nib)
    try:
        _cmp(tracs_b200.pairsnp_packed(d.value, n, L, pitch, dist=dist), orc)
    finally:
        _lib.lib().tracs_dev_free(d)


def test_encode_every_byte_value():
    """k_encode: all 256 byte values, lengths that are not multiples of 32, padding sites = 1111."""
    rng = np.random.default_rng(9)
    for L in (1, 31, 32, 33, 4095, 4096, 5001):
        s = rng.integers(0, 256, size=(19, L), dtype=np.uint8)
        s[0, :min(L, 256)] = np.arange(min(L, 256), dtype=np.uint8)
        apitch = (L + 31) // 32 * 32
        a = np.full((19, apitch), 0x5A, np.uint8)      # garbage beyond L must not leak into the packed rows
        a[:, :L] = s
        want, pitch = tracs_b200.pack_nibbles(s)
        da, dn = _dev(a), _dev(np.zeros_like(want))
        try:
            _lib.check(_lib.lib().tracs_encode_packed(da, 19, L, apitch, dn, pitch))
            got = np.empty_like(want)
            _lib.check(_lib.lib().tracs_memcpy_d2h(got.ctypes.data, dn, got.size))
        finally:
            _lib.lib().tracs_dev_free(da)
            _lib.lib().tracs_dev_free(dn)
        assert np.array_equal(got[:, :(L + 31) // 32 * 16], want[:, :(L + 31) // 32 * 16]), L


@pytest.mark.parametrize("n,L,p_var", [(600, 20000, 0.03), (300, 70001, 0.02), (1100, 9000, 0.04), (257, 4096, 0.01), (900, 33000, 0.2)])
def test_packed_early_extraction_matches_oracle(oracle_mod, monkeypatch, n, L, p_var):
    """k_pack4<true> (early extraction out of the packed rows) + late gather + k_slice vs the two-pass packed ingest
    (k_pack4<false> + k_gather<packed>) vs the oracle; with sites that only start to vary after the first chunk."""
    s = synth.generate(n, L, p_var=p_var, n_clusters=6, mu=3, p_N=0.02, p_amb=0.05, seed=n + L, lowercase=0.05,
                       odd_chars=0.01, three_base=True)
    rng = np.random.default_rng(n)
    for c in rng.integers(0, L, size=40):
        s[:, c] = ord("A")
        s[rng.integers(256, n), c] = ord("T")
    dist = IMAX if n < 1000 else 60
    orc = oracle_mod.pairsnp_ascii(s, dist=dist, n_threads=4)
    nib, pitch = tracs_b200.pack_nibbles(s)
    monkeypatch.setenv("TRACS_INGEST", "early")
    res = tracs_b200.pairsnp_packed_host(nib, L, dist=dist)
    st = tracs_b200.last_stats()
    assert st["n_early_sites"] > 0 or p_var > 1.0 / 16   # more than L/16 early sites: the ingest keeps the two-pass path
    _cmp(res, orc)
    monkeypatch.setenv("TRACS_INGEST", "split")
    _cmp(tracs_b200.pairsnp_packed_host(nib, L, dist=dist), orc)
    assert tracs_b200.last_stats()["n_variable_sites"] == st["n_variable_sites"]


def test_packed_filter_and_fused_trans(oracle_mod):
    s = synth.generate(60, 30_000, p_var=0.03, n_clusters=4, mu=5, p_N=0.002, seed=52)
    days = np.random.default_rng(3).integers(0, 100, size=60).astype(np.int32)
    nib, pitch = tracs_b200.pack_nibbles(s)
    res = tracs_b200.pairsnp_packed_host(nib, 30_000, dist=200, filter=True, days=days)
    r, c, d, f, nn = oracle_mod.pairsnp_ascii(s, dist=200, filter=True)
    _cmp(res, (r, c, d, f, nn))
    assert res["filt"].tolist() == f.tolist()
    dt = np.abs(days[r.astype(int)] * 86400.0 - days[c.astype(int)] * 86400.0) / 31556952.0
    op0, oeK = oracle_mod.trans_dist(f.astype(np.int32), dt, 29.903, 73.0, 0.01)
    assert np.allclose(res["p0_log"], op0, rtol=1e-6, atol=0)
    pos = dt > 0
    assert np.allclose(res["eK"][pos], oeK[pos], rtol=1e-6, atol=0)


def test_synth_packed_is_the_ascii_alignment_packed():
    import torch
    n, L = 70, 9_999
    kw = dict(seed=5, p_var=0.05, n_clusters=4, mu=3.0, p_N=0.01, p_amb=0.02, gc=0.4)
    apitch = (L + 127) // 128 * 128
    a = torch.empty((n, apitch), dtype=torch.uint8, device="cuda")
    tracs_b200.synth_device(a.data_ptr(), n, L, apitch, **kw)
    want, pitch = tracs_b200.pack_nibbles(a[:, :L].cpu().numpy())
    p = torch.empty((n, pitch), dtype=torch.uint8, device="cuda")
    tracs_b200.synth_device(p.data_ptr(), n, L, pitch, packed=True, **kw)
    assert np.array_equal(p.cpu().numpy(), want)
    # a column slab of the packed alignment
    lo, hi = 3840, 7680
    sp = torch.empty((n, (hi - lo) // 2), dtype=torch.uint8, device="cuda")
    tracs_b200.synth_device(sp.data_ptr(), n, hi - lo, (hi - lo) // 2, packed=True, site_offset=lo, L_total=L, **kw)
    assert np.array_equal(sp.cpu().numpy(), want[:, lo // 2:hi // 2])


def test_packed_device_large_equals_ascii_device():
    """Bench-like shape generated on the device in both formats: identical edge tables and variable-site counts,
    early-extraction path on both."""
    import torch
    n, L = 12000, 400_000
    kw = dict(seed=11, p_var=0.03, n_clusters=150, mu=5.0, p_N=1e-3, p_amb=0.002, gc=0.5)
    apitch = (L + 127) // 128 * 128
    a = torch.empty(n * apitch, dtype=torch.uint8, device="cuda")
    tracs_b200.synth_device(a.data_ptr(), n, L, apitch, **kw)
    ra = tracs_b200.pairsnp_device(a.data_ptr(), n, L, apitch, dist=20)
    sa = tracs_b200.last_stats()
    del a
    pitch = (L + 31) // 32 * 16
    p = torch.empty(n * pitch, dtype=torch.uint8, device="cuda")
    tracs_b200.synth_device(p.data_ptr(), n, L, pitch, packed=True, **kw)
    rp = tracs_b200.pairsnp_packed(p.data_ptr(), n, L, pitch, dist=20)
    sp = tracs_b200.last_stats()
    assert sp["n_early_sites"] > 0 and sa["n_early_sites"] == sp["n_early_sites"]
    assert sp["n_variable_sites"] == sa["n_variable_sites"] and len(rp["rows"]) > 1000
    for k in COLS:
        assert np.array_equal(ra[k], rp[k])


def _site_sharded_packed(torch, s, world, dist, days):
    """R emulated ranks on one GPU, packed slabs, native finish (collectives replaced by torch adds)."""
    from tracs_b200 import sites
    n, L = s.shape
    dev = torch.device("cuda")
    keep, handles, cands = [], [], []
    for r in range(world):
        lo, hi = sites.slab_bounds(L, r, world)
        Ls = hi - lo
        nib, pitch = tracs_b200.pack_nibbles(s[:, lo:hi]) if Ls else (np.full((n, 16), 255, np.uint8), 16)
        buf = torch.from_numpy(nib).to(dev)
        keep.append(buf)
        h, kptr, cnt, st = sites.open_shard(buf.data_ptr(), n, Ls, pitch, dist, r, world, packed=True)
        handles.append(h)
        cands.append(torch.as_tensor(sites._Dev(kptr, cnt * 8), device=dev).view(torch.int64).clone() if cnt
                     else torch.empty(0, dtype=torch.int64, device=dev))
    keys, _ = torch.sort(torch.cat(cands))
    E = keys.numel()
    both = torch.zeros((2, max(E, 1)), dtype=torch.int32, device=dev)
    for h in handles:
        part = torch.zeros_like(both)
        if E:
            _lib.check(_lib.lib().tracs_site_shard_partials(h, C.c_void_p(keys.data_ptr()), E, C.c_void_p(part[0].data_ptr()),
                                                            C.c_void_p(part[1].data_ptr())))
        both += part
        _lib.lib().tracs_site_shard_close(h)
    be = sites.LibBackend(torch, dev, 0, n, 0, 0, packed=True)
    res, st = be.finish(keys, both[0], both[1], L, dist, days, 29.903, 73.0, 0.01)
    return res, E


@pytest.mark.parametrize("world", [1, 2, 5])
def test_site_sharded_packed_native_finish(oracle_mod, world):
    import torch
    s = synth.generate(420, 150_000, p_var=0.06, n_clusters=50, mu=4, p_N=0.002, p_amb=0.01, seed=71)
    days = np.random.default_rng(4).integers(0, 150, size=420).astype(np.int32)
    for dist in (0, 25):
        res, n_cand = _site_sharded_packed(torch, s, world, dist, days)
        r, c, od, f, onn = oracle_mod.pairsnp_ascii(s, dist=dist, n_threads=4)
        _cmp(res, (r, c, od, f, onn))
        dt = np.abs(days[r.astype(int)] * 86400.0 - days[c.astype(int)] * 86400.0) / 31556952.0
        assert res["datediff"].tolist() == dt.tolist()
        op0, oeK = oracle_mod.trans_dist(od.astype(np.int32), dt, 29.903, 73.0, 0.01)
        assert np.allclose(res["p0_log"], op0, rtol=1e-6, atol=0)
        pos = dt > 0
        assert np.allclose(res["eK"][pos], oeK[pos], rtol=1e-6, atol=0)
        assert n_cand < 0.2 * 420 * 419 / 2


def test_sites_driver_world1_packed_matches_single_gpu():
    import torch
    from tracs_b200 import sites
    s = synth.generate(500, 150_000, p_var=0.06, n_clusters=60, mu=4, p_N=0.002, seed=81)
    days = np.random.default_rng(4).integers(0, 150, size=500).astype(np.int32)
    n, L = s.shape
    nib, pitch = tracs_b200.pack_nibbles(s)
    buf = torch.from_numpy(nib).cuda()
    res, st = sites.sweep(torch, None, torch.device("cuda"), 0, 1, buf.data_ptr(), n, L, pitch, L, 20, days=days, packed=True)
    ref = tracs_b200.pairsnp_matrix(s, dist=20, days=days)
    for k in COLS:
        assert res[k].tolist() == ref[k].tolist()
    assert np.array_equal(res["datediff"], ref["datediff"])
    assert np.allclose(res["p0_log"], ref["p0_log"], rtol=1e-12) and np.allclose(res["eK"], ref["eK"], rtol=1e-12)
    assert st["kernel_launches"] > 0 and st["n_edges"] == len(ref["rows"])


def test_filter_offsets_beyond_32_bits(oracle_mod):
    """with the filter on and an unthresholded sweep the summed SNP distances of one band pass
    2^32; the position offsets must be 64-bit. Pairs late in the edge list (whose offsets lie beyond the wrap) are
    checked against the oracle run on just their two rows (the filter of a pair depends on nothing else)."""
    rng = np.random.default_rng(123)
    n, L = 2600, 2400
    s = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(n, L))]
    res = tracs_b200.pairsnp_matrix(s, dist=IMAX, filter=True, want_ncomp=False)
    E = len(res["rows"])
    assert E == n * (n - 1) // 2
    assert int(res["dist"].sum()) > (1 << 32) + (1 << 30)
    assert np.all(res["filt"] <= res["dist"])
    for e in list(rng.integers(0, E, size=12)) + [E - 1, E - 2, E // 2 + 7]:
        i, j = int(res["rows"][e]), int(res["cols"][e])
        r, c, d, f, nn = oracle_mod.pairsnp_ascii(s[[i, j]], dist=IMAX, filter=True)
        assert int(res["dist"][e]) == int(d[0]) and int(res["filt"][e]) == int(f[0]), (e, i, j)


def test_fasta_entry_point_streams_rows_to_the_device(oracle_mod, tmp_path, monkeypatch):
    """tracs_pairsnp(paths): rows go to the device while the reader is still parsing (DeviceRowStreamer). Plain file
    (parallel reader, record count known), gzip (sequential reader: the packed matrix grows by doubling), two-file
    mode, tiny staging slots so that many batches are in flight -- all equal to the oracle and to the
    parse-then-copy path (TRACS_FASTA_STREAM=0)."""
    s = synth.generate(420, 30_000, p_var=0.05, n_clusters=6, mu=3, p_N=0.02, p_amb=0.03, seed=77, lowercase=0.05, odd_chars=0.01)
    monkeypatch.setenv("TRACS_STREAM_CHUNK_BYTES", str(64 * 30_016))
    monkeypatch.setenv("TRACS_FASTA_PAR_MIN", "0")
    files = {"plain": (str(tmp_path / "a.fa"), dict(width=0)), "wrapped": (str(tmp_path / "b.fa"), dict(width=70, descriptions=True)),
             "gz": (str(tmp_path / "c.fa.gz"), dict(width=60))}
    for tag, (p, kw) in files.items():
        synth.write_fasta(p, s, **kw)
        exp = oracle_mod.pairsnp([p], n_threads=4, dist=40)
        got = tracs_b200.pairsnp(fasta=[p], n_threads=8, dist=40, filter=False)
        st = tracs_b200.last_stats()
        assert st["h2d_bytes"] == s.size, tag
        monkeypatch.setenv("TRACS_FASTA_STREAM", "0")
        ref = tracs_b200.pairsnp(fasta=[p], n_threads=8, dist=40, filter=False)
        monkeypatch.delenv("TRACS_FASTA_STREAM")
        for t in range(6):
            assert got[t] == exp[t] and ref[t] == exp[t], (tag, t)
    p1, p2 = str(tmp_path / "q.fa"), str(tmp_path / "db.fa.gz")
    synth.write_fasta(p1, s[:100], names=["q%d" % i for i in range(100)])
    synth.write_fasta(p2, s[100:], names=["d%d" % i for i in range(320)])
    got = tracs_b200.pairsnp(fasta=[p1, p2], n_threads=4, dist=40, filter=False)
    exp = oracle_mod.pairsnp([p1, p2], dist=40)
    for t in range(6):
        assert got[t] == exp[t]
    bad = str(tmp_path / "other_len.fa")
    synth.write_fasta(bad, s[:5, :1000])
    with pytest.raises(RuntimeError, match="variable sequence lengths"):
        tracs_b200.pairsnp(fasta=[p1, bad], n_threads=2, dist=40, filter=False)


def test_sigint_between_bands_unwinds_the_call(tmp_path):
    """Ctrl-C during a long sweep (reference: 'Interrupted by user!' + exit(1), src/pairsnp.hpp:434-441): the library's
    own handler raises a flag, the band loop sees it, the call returns status 4 -> KeyboardInterrupt with that message;
    the previous SIGINT disposition is restored."""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import os, signal, sys, threading, time, torch
sys.path.insert(0, %r)
import tracs_b200
n, L = 60000, 40000
pitch = L // 2
buf = torch.empty(n * pitch, dtype=torch.uint8, device="cuda")
tracs_b200.synth_device(buf.data_ptr(), n, L, pitch, seed=9, p_var=1.0, n_clusters=n, mu=0.0, p_N=0.0, gc=0.5, gaps=0, packed=True)
tracs_b200.pairsnp_packed(buf.data_ptr(), 2000, L, pitch, dist=2047)          # warm-up (contexts, caches)
t0 = time.perf_counter()
tracs_b200.pairsnp_packed(buf.data_ptr(), n, L, pitch, dist=2047, full_sweep=True)   # unselective: banded full-length sweeps
t_full = time.perf_counter() - t0
threading.Timer(min(0.15, t_full / 4), lambda: os.kill(os.getpid(), signal.SIGINT)).start()
t0 = time.perf_counter()
try:
    tracs_b200.pairsnp_packed(buf.data_ptr(), n, L, pitch, dist=2047, full_sweep=True)
    print("NOT-INTERRUPTED")
except KeyboardInterrupt as ex:
    print("INTERRUPTED", str(ex), "%%.3f %%.3f" %% (time.perf_counter() - t0, t_full))
assert signal.getsignal(signal.SIGINT) is signal.default_int_handler
r = tracs_b200.pairsnp_packed(buf.data_ptr(), 3000, L, pitch, dist=2047)     # the library is usable afterwards
print("AFTER", len(r["rows"]))
''' % root
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "INTERRUPTED Interrupted by user!" in out.stdout and "AFTER" in out.stdout, out.stdout + out.stderr
    took, full = map(float, out.stdout.split("INTERRUPTED Interrupted by user!")[1].split()[:2])
    assert took < full, (took, full)

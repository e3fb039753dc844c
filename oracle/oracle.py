"""ctypes front-end of oracle/liboracle.so (tracs_oracle.c) + a slow NumPy cross-check.

TEST INFRASTRUCTURE ONLY (see tracs_oracle.c). Restates /root/reference/src/pairsnp.hpp and
src/transcluster.hpp; function-level citations live in the C file."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class _Edges(C.Structure):
    _fields_ = [("n_edges", C.c_uint64), ("rows", C.POINTER(C.c_uint64)), ("cols", C.POINTER(C.c_uint64)),
                ("dist", C.POINTER(C.c_uint64)), ("filt", C.POINTER(C.c_uint64)), ("ncomp", C.POINTER(C.c_uint64)),
                ("n_names", C.c_uint64), ("names", C.POINTER(C.c_char_p)), ("seq_length", C.c_uint64)]


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "tracs_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-std=gnu99", "-fPIC", "-fopenmp", "-shared", "-o", so, src, "-lz", "-lm"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_pairsnp.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_Edges), C.c_char_p, C.c_size_t]
        L.orc_pairsnp_ascii.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, C.POINTER(_Edges)]
        L.orc_edges_free.argtypes = [C.POINTER(_Edges)]
        L.orc_trans_dist.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_lprob_k_given_N.argtypes = [C.c_uint64, C.c_uint64, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
        L.orc_lprob_k_given_N_2.argtypes = [C.c_uint64, C.c_uint64, C.c_double, C.c_double, C.c_double, C.c_void_p]
        L.orc_base_mask.argtypes = [C.c_int]
        L.orc_base_mask.restype = C.c_uint8
        _LIB = L
    return _LIB


def _take(e, with_names=True):
    n = e.n_edges
    def arr(p):
        return np.ctypeslib.as_array(p, shape=(n,)).copy() if n else np.zeros(0, np.uint64)
    out = [arr(e.rows), arr(e.cols), arr(e.dist)]
    names = [e.names[i].decode() for i in range(e.n_names)] if with_names else []
    out += [names, arr(e.filt), arr(e.ncomp)]
    lib().orc_edges_free(C.byref(e))
    return tuple(out)


def pairsnp(fasta, n_threads=1, dist=2147483647, filter=False, as_lists=True):
    """Same contract as TRACS.pairsnp (src/python_bindings.cpp:12-13): 6-tuple
    (rows, cols, distances, names, filt, n_compared)."""
    paths = (C.c_char_p * len(fasta))(*[os.fsencode(p) for p in fasta])
    e = _Edges()
    err = C.create_string_buffer(512)
    rc = lib().orc_pairsnp(paths, len(fasta), n_threads, dist, int(bool(filter)), C.byref(e), err, 512)
    if rc:
        raise RuntimeError(err.value.decode())
    t = _take(e)
    if as_lists:
        return tuple(x.tolist() if isinstance(x, np.ndarray) else x for x in t)
    return t


def pairsnp_ascii(seqs, i_end=None, j_start=0, n_threads=1, dist=2147483647, filter=False):
    """Pair sweep on an ASCII matrix uint8[n][L] (what load_seqs holds per record). Returns
    numpy (rows, cols, d, filt, ncomp)."""
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
    n, L = seqs.shape
    if i_end is None:
        i_end = n
    e = _Edges()
    lib().orc_pairsnp_ascii(seqs.ctypes.data, n, L, i_end, j_start, n_threads, dist, int(bool(filter)), C.byref(e))
    r = _take(e, with_names=False)
    return r[0], r[1], r[2], r[4], r[5]


def trans_dist(snpdiff, datediff, lamb, beta, threshold_Ek, with_k_exit=False):
    """TRACS.trans_dist (src/python_bindings.cpp:19-21): returns (log p0, E[K]) as numpy arrays."""
    snp = np.ascontiguousarray(snpdiff, dtype=np.int32)
    dt = np.ascontiguousarray(datediff, dtype=np.float64)
    n = snp.size
    p0 = np.empty(n, np.float64)
    eK = np.empty(n, np.float64)
    ke = np.empty(n, np.int32)
    rc = lib().orc_trans_dist(snp.ctypes.data, dt.ctypes.data, n, lamb, beta, threshold_Ek, p0.ctypes.data, eK.ctypes.data, ke.ctypes.data)
    if rc:
        raise RuntimeError("negative SNP distance")
    return (p0, eK, ke) if with_k_exit else (p0, eK)


def lprob_k_given_N(N, k, delta, lamb, beta, lgamma):
    lg = np.ascontiguousarray(lgamma, dtype=np.float64)
    if lg.size < N + k + 2:
        raise IndexError("lgamma table too short")
    out = np.empty(2, np.float64)
    lib().orc_lprob_k_given_N(N, k, delta, lamb, beta, lg.ctypes.data, out.ctypes.data)
    return float(out[0]), float(out[1])


def lprob_k_given_N_2(N, k, delta, lamb, beta):
    out = np.empty(2, np.float64)
    lib().orc_lprob_k_given_N_2(N, k, delta, lamb, beta, out.ctypes.data)
    return float(out[0]), float(out[1])


# ---- independent, obviously-correct NumPy restatement (small inputs only) --------------------
_MASK = np.full(256, 15, np.uint8)
for _c, _m in dict(A=1, C=2, G=4, T=8, M=3, R=5, W=9, S=6, Y=10, K=12, V=7, H=11, D=13, B=14).items():
    _MASK[ord(_c)] = _m
    _MASK[ord(_c.lower())] = _m


def masks_of(seqs):
    """ASCII uint8[n][L] -> 4-bit base masks (src/pairsnp.hpp:107-199)."""
    return _MASK[np.asarray(seqs, dtype=np.uint8)]


def pairsnp_numpy(seqs, i_end=None, j_start=0, dist=2147483647):
    """Per-site definition: d = #{s: m_i[s] & m_j[s] == 0}; nn = L - #{s: m_i==15 or m_j==15}
    (src/pairsnp.hpp:398-403, :417-419); (i, j) lexicographic emission (:395, :450-457)."""
    m = masks_of(seqs)
    n, L = m.shape
    if i_end is None:
        i_end = n
    rows, cols, ds, nns = [], [], [], []
    isn = m == 15
    for i in range(i_end):
        j0 = max(j_start, i + 1)
        if j0 >= n:
            continue
        d = ((m[i][None, :] & m[j0:]) == 0).sum(axis=1)
        nn = L - (isn[i][None, :] | isn[j0:]).sum(axis=1)
        keep = np.nonzero(d <= dist)[0]
        rows += [i] * len(keep)
        cols += (keep + j0).tolist()
        ds += d[keep].tolist()
        nns += nn[keep].tolist()
    return (np.array(rows, np.uint64), np.array(cols, np.uint64), np.array(ds, np.uint64), np.array(nns, np.uint64))

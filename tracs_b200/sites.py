"""Site-sharded multi-GPU sweep (strong scaling of ONE alignment): rank r holds the column slab
[L*r/R, L*(r+1)/R) of every sequence, as ASCII or as 4-bit packed masks.

d(i,j) and |N_i u N_j| are sums over disjoint site ranges (src/pairsnp.hpp:398-403, 417-419), so (csrc/shard.inl):
  open      each rank ingests its slab and prefilters ITS share of the triangle row-blocks
  gather    candidate pair lists are all-gathered                        (NCCL all-gather, O(candidates))
  partials  every rank evaluates its slab's share of d and |N u N| for all candidates
  reduce    the two integer vectors are summed over ranks               (NCCL all-reduce)
  finish    rank 0, native (tracs_site_shard_finish): d <= dist, compared sites = L_total - union, fused
            transmission likelihood, columns to page-locked host memory
No bit-plane crosses NVLink. Results equal the single-GPU sweep of the whole alignment.

The collective plumbing is torch.distributed; the three compute steps go through a small backend object so that
the exchange logic runs under gloo on CPU in tests (tests/test_multi_gloo.py) with the oracle standing in."""
import ctypes as C

import numpy as np

from . import _lib, api


def slab_bounds(L_total, rank, world, align=128):
    """Column range of `rank`: contiguous, multiples of `align` except the last."""
    per = (L_total + world - 1) // world
    per = (per + align - 1) // align * align
    lo = min(L_total, rank * per)
    hi = min(L_total, lo + per)
    return lo, hi


class _Dev:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class LibBackend:
    """The three compute steps on libtracs_b200.so (device pointers in, device pointers out)."""

    def __init__(self, torch, device, slab_ptr, n, L_slab, pitch, packed=False):
        self.torch, self.device = torch, device
        self.slab_ptr, self.n, self.L_slab, self.pitch, self.packed = slab_ptr, n, L_slab, pitch, packed
        self.h = None

    def open(self, dist, rank, world):
        """-> (int64 device tensor of this rank's candidate keys, sorted; stats dict)"""
        o, _ = api.make_opts(dist=dist, shard_rank=rank, shard_world=world, packed=self.packed)
        h, keys, cnt = C.c_void_p(), C.c_void_p(), C.c_size_t(0)
        _lib.check(_lib.lib().tracs_site_shard_open(C.c_void_p(self.slab_ptr), self.n, self.L_slab, self.pitch, C.byref(o), C.byref(h),
                                                    C.byref(keys), C.byref(cnt)))
        self.h = h
        torch = self.torch
        mine = (torch.as_tensor(_Dev(keys.value, cnt.value * 8), device=self.device).view(torch.int64) if cnt.value
                else torch.empty(0, dtype=torch.int64, device=self.device))
        return mine, _lib.last_stats()

    def partials(self, keys, d, u):
        E = int(keys.numel())
        if E:
            _lib.check(_lib.lib().tracs_site_shard_partials(self.h, C.c_void_p(keys.data_ptr()), E, C.c_void_p(d.data_ptr()),
                                                            C.c_void_p(u.data_ptr())))
        return _lib.last_stats()

    def finish(self, keys, d, u, L_total, dist, days, lamb, beta, threshold_Ek):
        o, keep = api.make_opts(dist=dist, days=days, lamb=lamb, beta=beta, threshold_Ek=threshold_Ek)
        e = _lib.Edges()
        _lib.check(_lib.lib().tracs_site_shard_finish(C.c_void_p(keys.data_ptr()), C.c_void_p(d.data_ptr()), C.c_void_p(u.data_ptr()),
                                                      int(keys.numel()), self.n, L_total, C.byref(o), C.byref(e)))
        return _lib.take_edges(e, names=False, copy=False), _lib.last_stats()

    def close(self):
        if self.h is not None:
            _lib.lib().tracs_site_shard_close(self.h)
            self.h = None


def exchange_candidates(torch, dist_mod, device, world, mine):
    """All ranks end up with the union of the per-rank candidate lists, sorted by key (= (row, col) order).
    Row-block shards are disjoint, so the union has no duplicates."""
    if world == 1:
        return mine.clone()
    cnt = torch.tensor([int(mine.numel())], dtype=torch.int64, device=device)
    cnts = torch.zeros(world, dtype=torch.int64, device=device)
    dist_mod.all_gather_into_tensor(cnts, cnt)
    cs = cnts.tolist()                      # the one host round trip of the exchange
    mx = max(cs + [1])
    pad = torch.zeros(mx, dtype=torch.int64, device=device)
    pad[:mine.numel()] = mine
    allk = torch.empty(world * mx, dtype=torch.int64, device=device)
    dist_mod.all_gather_into_tensor(allk, pad)
    keys = torch.cat([allk[r * mx:r * mx + k] for r, k in enumerate(cs)])
    keys, _ = torch.sort(keys)
    return keys


def sweep(torch, dist_mod, device, rank, world, slab_ptr, n, L_slab, pitch, L_total, dist, days=None, lamb=29.903, beta=73.0,
          threshold_Ek=0.01, packed=False, backend=None, profile=False):
    """All ranks call this. Returns (edge table on rank 0 / None elsewhere, per-rank stats dict).
    profile=True adds host wall-clock times of the phases (each closed by a device synchronise: slower, diagnosis only)."""
    import time
    be = backend if backend is not None else LibBackend(torch, device, slab_ptr, n, L_slab, pitch, packed)
    marks = []

    def mark(name):
        if profile:
            if device.type == "cuda":
                torch.cuda.synchronize(device)
            marks.append((name, time.perf_counter()))
    try:
        mark("start")
        mine, st_open = be.open(dist, rank, world)
        mark("open")
        keys = exchange_candidates(torch, dist_mod, device, world, mine)
        mark("exchange")
        E = int(keys.numel())
        both = torch.zeros((2, max(E, 1)), dtype=torch.int32, device=device)
        st_part = be.partials(keys, both[0], both[1])
        mark("partials")
        if world > 1:
            dist_mod.all_reduce(both)
        mark("allreduce")
        stats = dict(st_open)
        stats.update(st_part)   # the library's counters run on from open() through partials()
        stats["n_candidates_all"] = E
        if profile:
            stats["phases_ms"] = {b[0]: 1e3 * (b[1] - a[1]) for a, b in zip(marks, marks[1:])}
        if rank != 0:
            return None, stats
        res, st_fin = be.finish(keys, both[0], both[1], L_total, dist, days, lamb, beta, threshold_Ek)
        mark("finish")
        if profile:
            stats["phases_ms"] = {b[0]: 1e3 * (b[1] - a[1]) for a, b in zip(marks, marks[1:])}
        stats["kernel_launches"] = stats.get("kernel_launches", 0) + st_fin.get("kernel_launches", 0)
        for k in ("ms_trans", "ms_d2h", "d2h_bytes", "n_edges"):
            stats[k] = st_fin.get(k, 0)
        stats["ms_finish"] = st_fin.get("ms_sort", 0.0)
        return res, stats
    finally:
        be.close()


def open_shard(slab_ptr, n, L_slab, pitch, dist, rank, world, packed=False):
    """Bare library call (tests drive several emulated ranks on one GPU with it):
    -> (handle, device pointer of the sorted candidate keys, count, stats)."""
    o, _ = api.make_opts(dist=dist, shard_rank=rank, shard_world=world, packed=packed)
    h, keys, cnt = C.c_void_p(), C.c_void_p(), C.c_size_t(0)
    _lib.check(_lib.lib().tracs_site_shard_open(C.c_void_p(slab_ptr), n, L_slab, pitch, C.byref(o), C.byref(h), C.byref(keys), C.byref(cnt)))
    return h, (keys.value or 0), cnt.value, _lib.last_stats()

"""Site-sharded multi-GPU sweep (strong scaling of ONE alignment): rank r holds the column slab
[L*r/R, L*(r+1)/R) of every sequence, as ASCII or as 4-bit packed masks.

d(i,j) and |N_i u N_j| are sums over disjoint site ranges (src/pairsnp.hpp:398-403, 417-419), so (csrc/shard.inl):
  open      each rank ingests its slab and prefilters ITS share of the triangle row-blocks
  gather    candidate pair lists are all-gathered                        (NCCL all-gather, O(candidates))
  partials  every rank evaluates its slab's share of d and |N u N| for all candidates
  reduce    the two integer vectors are summed over ranks, each rank receiving the slice it will finish
                                                                         (NCCL reduce-scatter)
  finish    every rank, native, on its slice of the candidates (tracs_site_shard_select / _emit): d <= dist, compared
            sites = L_total - union, fused transmission likelihood; each GPU copies its edges into ONE shared,
            page-locked host table over its own PCIe link (SharedEdgeTable); world 1: tracs_site_shard_finish
No bit-plane crosses NVLink. Results equal the single-GPU sweep of the whole alignment.

The collective plumbing is torch.distributed; the three compute steps go through a small backend object so that
the exchange logic runs under gloo on CPU in tests (tests/test_multi_gloo.py) with the oracle standing in."""
import ctypes as C
import itertools
import mmap
import os

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
        self.trans = None

    def prepare(self, days, lamb, beta, threshold_Ek):
        """Optional, before open(): with the sampling days known up front the library computes the likelihood table on a
        side stream while the slab is ingested (the emit step of the same sweep picks it up)."""
        self.trans = dict(days=days, lamb=lamb, beta=beta, threshold_Ek=threshold_Ek) if days is not None else None

    def open(self, dist, rank, world):
        """-> (int64 device tensor of this rank's candidate keys, sorted; stats dict)"""
        o, _keep = api.make_opts(dist=dist, shard_rank=rank, shard_world=world, packed=self.packed, **(self.trans or {}))
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

    def select(self, keys, d, u, L_total, dist, days, lamb, beta, threshold_Ek):
        """First half of the sharded finish on a slice of the candidates -> (selection handle, edge count, stats)."""
        o, keep = api.make_opts(dist=dist, days=days, lamb=lamb, beta=beta, threshold_Ek=threshold_Ek)
        sel, cnt = C.c_void_p(), C.c_size_t(0)
        _lib.check(_lib.lib().tracs_site_shard_select(C.c_void_p(keys.data_ptr()), C.c_void_p(d.data_ptr()), C.c_void_p(u.data_ptr()),
                                                      int(keys.numel()), self.n, L_total, C.byref(o), C.byref(sel), C.byref(cnt)))
        return sel, cnt.value, _lib.last_stats()

    def emit(self, sel, table, at):
        """Second half: likelihood columns + device -> host copies into the shared table at row `at`."""
        e = table.as_edges()
        ht = C.c_int(0)
        _lib.check(_lib.lib().tracs_site_shard_emit(sel, C.byref(e), int(at), C.byref(ht)))
        return bool(ht.value), _lib.last_stats()

    def close(self):
        if self.h is not None:
            _lib.lib().tracs_site_shard_close(self.h)
            self.h = None


class SharedEdgeTable:
    """The edge table of a multi-process run: seven 8-byte columns in ONE shared-memory segment that every rank maps
    and page-locks, so that each GPU copies its slice of the edges straight into place over its own PCIe link and
    rank 0 reads the whole table without a gather through one GPU. Rank 0 creates the segment, the name is broadcast,
    and the file is unlinked as soon as everybody has mapped it."""
    COLS = (("rows", np.uint64), ("cols", np.uint64), ("dist", np.uint64), ("ncomp", np.uint64), ("p0_log", np.float64),
            ("eK", np.float64), ("datediff", np.float64))
    _serial = itertools.count()

    def __init__(self, dist_mod, rank, world, capacity, page_lock=True):
        self.capacity = int(capacity)
        self.nbytes = 8 * len(self.COLS) * self.capacity
        name = [None]
        if rank == 0:
            name[0] = "/dev/shm/tracs_edges_%d_%d" % (os.getpid(), next(self._serial))
            fd = os.open(name[0], os.O_CREAT | os.O_RDWR | os.O_EXCL, 0o600)
            os.ftruncate(fd, self.nbytes)
        if world > 1:
            dist_mod.broadcast_object_list(name, src=0)
        if rank != 0:
            fd = os.open(name[0], os.O_RDWR)
        self.mm = mmap.mmap(fd, self.nbytes)
        os.close(fd)
        self.buf = np.frombuffer(self.mm, dtype=np.uint8)
        self.locked = False
        if page_lock:
            _lib.check(_lib.lib().tracs_host_register(self.buf.ctypes.data, self.nbytes))
            self.locked = True
        if world > 1:
            dist_mod.barrier()
        if rank == 0:
            os.unlink(name[0])
        self.views = {}
        for k, (c, dt) in enumerate(self.COLS):
            self.views[c] = self.buf[k * 8 * self.capacity:(k + 1) * 8 * self.capacity].view(dt)

    def as_edges(self):
        """tracs_edges_t whose column pointers are the table's columns (destination of tracs_site_shard_emit)."""
        e = _lib.Edges()
        for c, dt in self.COLS:
            ct = C.c_uint64 if dt is np.uint64 else C.c_double
            setattr(e, c, C.cast(self.views[c].ctypes.data, C.POINTER(ct)))
        return e

    def close(self):
        if self.locked:
            _lib.lib().tracs_host_unregister(self.buf.ctypes.data)
            self.locked = False
        self.views = {}
        self.buf = None
        try:
            self.mm.close()
        except BufferError:
            pass


_TABLE = None  # this process's shared edge table (kept across sweeps, grown collectively when too small)


def _shared_table(dist_mod, rank, world, need, page_lock):
    global _TABLE
    if _TABLE is None or _TABLE.capacity < need:     # `need` is the same on every rank: a collective decision
        if _TABLE is not None:
            _TABLE.close()
        _TABLE = SharedEdgeTable(dist_mod, rank, world, int(need * 1.25) + 1024, page_lock=page_lock)
    return _TABLE


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
        if world > 1 and hasattr(be, "prepare"):
            be.prepare(days, lamb, beta, threshold_Ek)
        mine, st_open = be.open(dist, rank, world)
        mark("open")
        keys = exchange_candidates(torch, dist_mod, device, world, mine)
        mark("exchange")
        E = int(keys.numel())
        # every rank finishes the slice [rank * chunk, (rank + 1) * chunk) of the candidates, so it only needs the summed
        # vectors there: a reduce-scatter (half the traffic of an all-reduce) where the backend has one
        chunk = (max(E, 1) + world - 1) // world
        both = torch.zeros((2, chunk * world), dtype=torch.int32, device=device)
        st_part = be.partials(keys, both[0], both[1])
        mark("partials")
        scattered = None
        if world > 1:
            if dist_mod.get_backend() == "nccl":
                scattered = torch.empty((2, chunk), dtype=torch.int32, device=device)
                dist_mod.reduce_scatter_tensor(scattered[0], both[0])
                dist_mod.reduce_scatter_tensor(scattered[1], both[1])
            else:
                dist_mod.all_reduce(both)
        mark("allreduce")
        stats = dict(st_open)
        stats.update(st_part)   # the library's counters run on from open() through partials()
        stats["n_candidates_all"] = E
        if world > 1:
            # sharded finish: every rank thresholds a contiguous slice of the (identical) summed vectors, the counts are
            # all-gathered, and each rank copies its edges into the shared host table at its offset
            lo, hi = min(E, rank * chunk), min(E, (rank + 1) * chunk)
            dsum, usum = (scattered[0][:hi - lo], scattered[1][:hi - lo]) if scattered is not None else (both[0][lo:hi], both[1][lo:hi])
            sel, cnt, st_sel = be.select(keys[lo:hi], dsum, usum, L_total, dist, days, lamb, beta, threshold_Ek)
            cnts = torch.zeros(world, dtype=torch.int64, device=device)
            dist_mod.all_gather_into_tensor(cnts, torch.tensor([cnt], dtype=torch.int64, device=device))
            cs = cnts.tolist()
            total = int(sum(cs))
            table = _shared_table(dist_mod, rank, world, total, page_lock=(device.type == "cuda"))
            has_trans, st_emit = be.emit(sel, table, int(sum(cs[:rank])))
            dist_mod.barrier()      # every slice has landed in the table
            mark("finish")
            for k in ("ms_trans", "ms_d2h", "d2h_bytes"):
                stats[k] = st_emit.get(k, 0)
            stats["kernel_launches"] = stats.get("kernel_launches", 0) + st_sel.get("kernel_launches", 0) + st_emit.get("kernel_launches", 0)
            stats["ms_finish"] = st_sel.get("ms_sort", 0.0)
            stats["n_edges"] = total
            if profile:
                stats["phases_ms"] = {b[0]: 1e3 * (b[1] - a[1]) for a, b in zip(marks, marks[1:])}
            if rank != 0:
                return None, stats
            res = {c: table.views[c][:total] for c, _ in SharedEdgeTable.COLS if has_trans or c not in ("p0_log", "eK", "datediff")}
            for c in ("p0_log", "eK", "datediff"):
                res.setdefault(c, None)
            return res, stats     # views of the shared table: valid until the next sweep
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

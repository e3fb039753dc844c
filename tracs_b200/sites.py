"""Site-sharded multi-GPU sweep: rank r holds the column slab [L*r/R, L*(r+1)/R) of every sequence.

d(i,j) and |N_i u N_j| are sums over disjoint site ranges, so (see csrc/shard.inl):
  open      each rank ingests its slab and prefilters ITS share of the triangle row-blocks
  gather    candidate pair lists are all-gathered                      (NCCL, O(candidates))
  partials  every rank evaluates its slab's share of d and |N u N| for all candidates
  reduce    the two integer vectors are summed over ranks             (NCCL all-reduce)
  keep      d <= dist; compared sites = L_total - union; optional transmission likelihood
No bit-plane crosses NVLink. Results equal the single-GPU sweep of the whole alignment."""
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


def open_shard(slab_ptr, n, L_slab, pitch, dist, rank, world):
    o, _ = api.make_opts(dist=dist, shard_rank=rank, shard_world=world)
    h, keys, cnt = C.c_void_p(), C.c_void_p(), C.c_size_t(0)
    _lib.check(_lib.lib().tracs_site_shard_open(C.c_void_p(slab_ptr), n, L_slab, pitch, C.byref(o), C.byref(h), C.byref(keys), C.byref(cnt)))
    return h, (keys.value or 0), cnt.value, _lib.last_stats()


def sweep(torch, dist_mod, device, rank, world, slab_ptr, n, L_slab, pitch, L_total, dist, days=None, lamb=29.903, beta=73.0,
          threshold_Ek=0.01):
    """All ranks call this. Returns (result dict on rank 0 / None elsewhere, per-rank stats dict)."""
    h, kptr, cnt, st_open = open_shard(slab_ptr, n, L_slab, pitch, dist, rank, world)
    try:
        mine = (torch.as_tensor(_Dev(kptr, cnt * 8), device=device).view(torch.int64) if cnt
                else torch.empty(0, dtype=torch.int64, device=device))
        if world > 1:
            c = torch.tensor([cnt], dtype=torch.int64, device=device)
            cs = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
            dist_mod.all_gather(cs, c)
            cs = [int(x.item()) for x in cs]
            mx = max(cs + [1])
            pad = torch.zeros(mx, dtype=torch.int64, device=device)
            pad[:cnt] = mine
            bufs = [torch.empty(mx, dtype=torch.int64, device=device) for _ in range(world)]
            dist_mod.all_gather(bufs, pad)
            keys = torch.cat([b[:k] for b, k in zip(bufs, cs)])
            keys, _ = torch.sort(keys)      # row-block shards are disjoint: a plain sort restores (row, col) order
        else:
            keys = mine.clone()
        E = int(keys.numel())
        d = torch.zeros(max(E, 1), dtype=torch.int32, device=device)
        u = torch.zeros(max(E, 1), dtype=torch.int32, device=device)
        if E:
            _lib.check(_lib.lib().tracs_site_shard_partials(h, C.c_void_p(keys.data_ptr()), E, C.c_void_p(d.data_ptr()), C.c_void_p(u.data_ptr())))
        st_part = _lib.last_stats()
        if world > 1:
            both = torch.stack([d, u])
            dist_mod.all_reduce(both)
            d, u = both[0], both[1]
        stats = dict(st_open)
        stats["ms_refine"] = st_part["ms_refine"]
        stats["n_candidates_all"] = E
        if rank != 0:
            return None, stats
        keep = (d[:E] <= dist) if E else torch.zeros(0, dtype=torch.bool, device=device)
        k = keys[:E][keep].cpu().numpy().astype(np.uint64)
        res = {"rows": k >> np.uint64(32), "cols": k & np.uint64(0xFFFFFFFF),
               "dist": d[:E][keep].cpu().numpy().astype(np.uint64),
               "ncomp": (L_total - u[:E][keep].cpu().numpy().astype(np.int64)).astype(np.uint64),
               "p0_log": None, "eK": None, "datediff": None}
        if days is not None and len(k):
            days = np.asarray(days)
            dt = np.abs(days[res["rows"].astype(np.int64)] * 86400.0 - days[res["cols"].astype(np.int64)] * 86400.0) / 31556952.0
            p0, eK = api.trans_dist(res["dist"].astype(np.int32), dt, lamb, beta, threshold_Ek)
            res["p0_log"], res["eK"], res["datediff"] = np.asarray(p0), np.asarray(eK), dt
        return res, stats
    finally:
        _lib.lib().tracs_site_shard_close(h)

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


_OUT = None  # cached page-locked result block (rank 0)


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
        # rank 0: threshold, compared sites and the transmission table stay on the device; ONE copy of the
        # finished columns into cached page-locked memory
        keep = (d[:E] <= dist) if E else torch.zeros(0, dtype=torch.bool, device=device)
        k = keys[:E][keep]
        rows_t, cols_t = k >> 32, k & 0xFFFFFFFF
        d_t = d[:E][keep].to(torch.int64)
        nn_t = L_total - u[:E][keep].to(torch.int64)
        cols_out = [rows_t, cols_t, d_t, nn_t]
        names = ["rows", "cols", "dist", "ncomp"]
        if days is not None and k.numel():
            # day-resolution dates: the memo key of trans_dist is (d, |day_i - day_j|); evaluate each used key once
            days_t = torch.as_tensor(np.asarray(days, dtype=np.int64), device=device)
            dd = (days_t[rows_t] - days_t[cols_t]).abs()
            DD = int(dd.max().item()) + 1
            key = d_t * DD + dd
            used = torch.nonzero(torch.bincount(key, minlength=(dist + 1) * DD)).flatten().cpu().numpy()
            p0k, eKk = api.trans_dist_np((used // DD).astype(np.int32), (used % DD) * 86400.0 / 31556952.0, lamb, beta, threshold_Ek)
            lut = np.zeros((2, (dist + 1) * DD))
            lut[0, used], lut[1, used] = p0k, eKk
            lut_t = torch.as_tensor(lut, device=device)
            # tensor / tensor is a true IEEE division (tensor / python-scalar multiplies by the reciprocal on CUDA)
            year = torch.full((), 31556952.0, dtype=torch.float64, device=device)
            cols_out += [lut_t[0][key], lut_t[1][key], torch.div(dd.to(torch.float64) * 86400.0, year)]
            names += ["p0_log", "eK", "datediff"]
        res = {"p0_log": None, "eK": None, "datediff": None}
        n_out = int(k.numel())
        global _OUT
        need = 8 * n_out * len(cols_out)
        if _OUT is None or _OUT.numel() < need:
            _OUT = torch.empty(int(need * 1.25) + 64, dtype=torch.uint8, pin_memory=(device.type == "cuda"))
        off = 0
        for nm, t in zip(names, cols_out):
            dst = _OUT[off:off + 8 * n_out].view(t.dtype)
            dst.copy_(t, non_blocking=True)
            res[nm] = dst
            off += 8 * n_out
        if device.type == "cuda":
            torch.cuda.current_stream(device).synchronize()
        for nm in names:   # NumPy views of the pinned block: valid until the next sweep() call
            a = res[nm].numpy()
            res[nm] = a.view(np.uint64) if nm in ("rows", "cols", "dist", "ncomp") else a
        return res, stats
    finally:
        _lib.lib().tracs_site_shard_close(h)

"""CPU, world_size 2 over gloo: the N>1 path's host logic -- per-rank shards (emulated with the CPU
oracle restricted to the rank's row-blocks) gathered to rank 0 and merged must equal the
single-process result."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import tracs_b200
    from tracs_b200 import synth
    from tracs_b200.multi import gather_edges
    from oracle import oracle
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s = synth.generate(700, 900, p_var=0.05, n_clusters=5, mu=3, p_N=0.01, seed=17)
    r, c, d, f, nn = oracle.pairsnp_ascii(s, dist=60, n_threads=2)
    mine = tracs_b200.shard_rowblocks((700 + 127) // 128, world, rank)
    keep = np.isin(r // np.uint64(128), mine.astype(np.uint64))
    res = {"rows": r[keep], "cols": c[keep], "dist": d[keep], "ncomp": nn[keep], "p0_log": None, "eK": None}
    merged = gather_edges(res, rank, world, dist, torch, torch.device("cpu"))
    ok = True
    if rank == 0:
        ok = (merged["rows"].tolist() == r.tolist() and merged["cols"].tolist() == c.tolist()
              and merged["dist"].tolist() == d.tolist() and merged["ncomp"].tolist() == nn.tolist())
        parts = None
    else:
        ok = merged is None
    # one-MSA-per-rank mode: per-rank tables come back unmerged, in rank order
    parts = gather_edges(res, rank, world, dist, torch, torch.device("cpu"), merge=False)
    if rank == 0:
        ok = ok and len(parts) == world and parts[0]["rows"].tolist() == r[keep].tolist() and sum(len(p["rows"]) for p in parts) == len(r)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, bool(ok), int(keep.sum())))


@pytest.mark.timeout(300)
def test_two_rank_gather_equals_single(oracle_mod):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in out), out
    assert all(cnt > 0 for _, _, cnt in out), out


class _OracleBackend:
    """Stands in for LibBackend in sites.sweep on CPU: the three compute steps of a rank restated with NumPy masks
    over the rank's column slab (per-site definition, src/pairsnp.hpp:398-403, 417-419)."""

    def __init__(self, torch, masks_slab, rank, world):
        self.torch, self.m, self.rank, self.world = torch, masks_slab, rank, world

    def open(self, dist, rank, world):
        from tracs_b200.multi import shard_owner
        n = self.m.shape[0]
        keys = []
        for i in range(n):
            if shard_owner(i // 128, world) != rank or i + 1 >= n:
                continue
            d = ((self.m[i][None, :] & self.m[i + 1:]) == 0).sum(axis=1)   # slab-partial distance: a lower bound of d
            for j in (np.nonzero(d <= dist)[0] + i + 1).tolist():
                keys.append((i << 32) | j)
        return self.torch.tensor(sorted(keys), dtype=self.torch.int64), {"kernel_launches": 1}

    def partials(self, keys, d, u):
        k = keys.numpy()
        for e, key in enumerate(k.tolist()):
            i, j = key >> 32, key & 0xFFFFFFFF
            d[e] = int(((self.m[i] & self.m[j]) == 0).sum())
            u[e] = int(((self.m[i] == 15) | (self.m[j] == 15)).sum())
        return {"kernel_launches": 2}

    def finish(self, keys, d, u, L_total, dist, days, lamb, beta, threshold_Ek):
        k, dn, un = keys.numpy().astype(np.uint64), d.numpy(), u.numpy()
        keep = dn[:len(k)] <= dist
        return {"rows": k[keep] >> np.uint64(32), "cols": k[keep] & np.uint64(0xFFFFFFFF), "dist": dn[:len(k)][keep].astype(np.uint64),
                "ncomp": (L_total - un[:len(k)][keep].astype(np.int64)).astype(np.uint64)}, {"kernel_launches": 3, "n_edges": int(keep.sum())}

    def select(self, keys, d, u, L_total, dist, days, lamb, beta, threshold_Ek):
        res, st = self.finish(keys, d, u, L_total, dist, days, lamb, beta, threshold_Ek)
        return res, len(res["rows"]), {"kernel_launches": 3}

    def emit(self, sel, table, at):
        for c in ("rows", "cols", "dist", "ncomp"):
            table.views[c][at:at + len(sel[c])] = sel[c]
        return False, {"kernel_launches": 1}

    def close(self):
        pass


def _sites_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from tracs_b200 import synth, sites
    from oracle import oracle
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, L, thr = 300, 1500, 25
    s = synth.generate(n, L, p_var=0.2, n_clusters=12, mu=3, p_N=0.02, p_amb=0.03, seed=23)
    lo, hi = sites.slab_bounds(L, rank, world)
    be = _OracleBackend(torch, oracle.masks_of(s[:, lo:hi]), rank, world)
    res, st = sites.sweep(torch, dist, torch.device("cpu"), rank, world, 0, n, hi - lo, 0, L, thr, backend=be)
    ok = True
    if rank == 0:
        r, c, d, f, nn = oracle.pairsnp_ascii(s, dist=thr, n_threads=2)
        ok = (res["rows"].tolist() == r.tolist() and res["cols"].tolist() == c.tolist() and res["dist"].tolist() == d.tolist()
              and res["ncomp"].tolist() == nn.tolist() and len(r) > 50 and st["n_candidates_all"] >= len(r))
    else:
        ok = res is None
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, bool(ok), int(st["n_candidates_all"])))


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [2, 3])
def test_two_rank_site_sharded_exchange(oracle_mod, world):
    """tracs_b200/sites.py at world_size 2 and 3 over gloo: slabs of one alignment (uneven at 3), candidate all-gather,
    partial sums summed over ranks, every rank finishes its slice of the candidates into the shared table (the last slice is
    shorter at 3) -- equals the single-process oracle."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() * 7 + world) % 2000
    procs = [ctx.Process(target=_sites_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in out), out
    assert len({c for _, _, c in out}) == 1 and out[0][2] > 0, out

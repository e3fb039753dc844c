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
